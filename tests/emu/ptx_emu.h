// Host versions of the helpers csrc/common.cuh implements with inline PTX: mbarrier,
// TMA bulk copies, system-scope flags, cache-policy loads.  Semantics kept:
//
//   mbarrier   a 64-bit word {phase, expected arrivals, pending arrivals, pending bytes};
//              a phase completes when both pending counts reach zero; try_wait(parity)
//              is true once the phase of that parity has completed.
//   TMA        bulk copies complete immediately (a legal schedule of the asynchronous
//              engine) and credit their bytes to the barrier.
//   flags      plain loads/stores (one thread runs at a time under the emulator).
#pragma once
#include <time.h>

struct bs_mbar_emu {
  unsigned phase : 1;
  unsigned expected : 15;
  unsigned pending : 16;
  int tx;
};
static_assert(sizeof(bs_mbar_emu) == 8, "mbarrier word");
static inline bs_mbar_emu* bs_mb(uint64_t* bar) { return reinterpret_cast<bs_mbar_emu*>(bar); }
static inline void bs_mb_check(bs_mbar_emu* m) {
  if (m->pending == 0 && m->tx == 0) {
    m->phase ^= 1u;
    m->pending = m->expected;
  }
}
static inline void bs_mbar_init(uint64_t* bar, uint32_t count) {
  bs_mbar_emu* m = bs_mb(bar);
  m->phase = 0;
  m->expected = count;
  m->pending = count;
  m->tx = 0;
}
static inline void bs_fence_barrier_init() {}
static inline void bs_fence_proxy_async() {}
static inline void bs_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  bs_mbar_emu* m = bs_mb(bar);
  m->tx += (int)bytes;
  m->pending -= 1;
  bs_mb_check(m);
}
static inline void bs_mbar_arrive(uint64_t* bar) {
  bs_mbar_emu* m = bs_mb(bar);
  m->pending -= 1;
  bs_mb_check(m);
}
static inline bool bs_mbar_try_wait(uint64_t* bar, uint32_t parity) { return bs_mb(bar)->phase != (parity & 1u); }
static inline void bs_mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!bs_mbar_try_wait(bar, parity)) bs_yield();
}
static inline void bs_tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  if (((uintptr_t)smem_dst | (uintptr_t)gsrc | bytes) & 15) abort();  // the hardware's alignment rule
  memcpy(smem_dst, gsrc, bytes);
  bs_mbar_emu* m = bs_mb(bar);
  m->tx -= (int)bytes;
  bs_mb_check(m);
}
static inline void bs_tma_load_1d_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                       unsigned long long) {
  bs_tma_load_1d(smem_dst, gsrc, bytes, bar);
}
static inline void bs_tma_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  if (((uintptr_t)gdst | (uintptr_t)smem_src | bytes) & 15) abort();
  memcpy(gdst, smem_src, bytes);
}
static inline void bs_tma_store_commit() {}
static inline void bs_tma_store_wait_all() {}
static inline void bs_tma_store_wait_read() {}

static inline void bs_st_release_sys(unsigned long long* p, unsigned long long v) { *(volatile unsigned long long*)p = v; }
static inline unsigned long long bs_ld_acquire_sys(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }
static inline unsigned long long bs_ld_relaxed_sys(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }
static inline unsigned long long bs_globaltimer_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (unsigned long long)ts.tv_sec * 1000000000ULL + (unsigned long long)ts.tv_nsec;
}
static inline int4 bs_ldg_stream(const int4* p) { return *p; }
static inline int4 bs_ldg_nocache(const int4* p) { return *p; }
static inline unsigned long long bs_policy_evict_last() { return 0; }
static inline unsigned long long bs_policy_evict_first() { return 0; }
static inline long long bs_ld_hint_s64(const long long* p, unsigned long long) { return *p; }
