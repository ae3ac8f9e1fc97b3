// Shared device helpers for the bigslice_b200 sm_100a kernels.
//
// Everything here is hand-written for Blackwell (compute_100a): seeded
// MurmurHash3_x86_32 that is bit-exact with the framework's host reference
// (bigslice_b200/frame/hashing.py; reference semantics frame/ops_builtin.go:
// 143-164), mbarrier + cp.async.bulk (TMA 1-D bulk copy) wrappers, and the
// system-scope release/acquire primitives used by the NVLink peer exchange.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define BS_EXPORT extern "C" __attribute__((visibility("default")))

// ---- host-side bookkeeping ------------------------------------------------
extern "C" long long bs_launch_counter;  // kernels launched by this library
#define BS_COUNT_LAUNCH() (__sync_fetch_and_add(&bs_launch_counter, 1))

static inline int bs_check(cudaError_t e) { return (int)e; }
#define BS_LAUNCH_CHECK()                         \
  do {                                            \
    BS_COUNT_LAUNCH();                            \
    cudaError_t _e = cudaGetLastError();          \
    if (_e != cudaSuccess) return (int)_e;        \
  } while (0)

static inline int bs_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- murmur3 ---------------------------------------------------------------
__device__ __forceinline__ uint32_t bs_rotl32(uint32_t x, int r) {
  return __funnelshift_l(x, x, r);
}
__device__ __forceinline__ uint32_t bs_mm3_k(uint32_t k) {
  k *= 0xcc9e2d51u;
  k = bs_rotl32(k, 15);
  return k * 0x1b873593u;
}
__device__ __forceinline__ uint32_t bs_mm3_h(uint32_t h, uint32_t k) {
  h ^= bs_mm3_k(k);
  h = bs_rotl32(h, 13);
  return h * 5u + 0xe6546b64u;
}
__device__ __forceinline__ uint32_t bs_fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
// murmur3_x86_32 over the 8 little-endian bytes of v.
__device__ __forceinline__ uint32_t bs_murmur3_u64(uint64_t v, uint32_t seed) {
  uint32_t h = seed;
  h = bs_mm3_h(h, (uint32_t)v);
  h = bs_mm3_h(h, (uint32_t)(v >> 32));
  return bs_fmix32(h ^ 8u);
}
// murmur3_x86_32 over the 4 little-endian bytes of v.
__device__ __forceinline__ uint32_t bs_murmur3_u32(uint32_t v, uint32_t seed) {
  uint32_t h = bs_mm3_h(seed, v);
  return bs_fmix32(h ^ 4u);
}
// Cheap 64->32 mixer for *table placement* (not observable placement): two
// 32-bit multiplies.  Independent of murmur3 so that keys which collide modulo
// the shard count do not cluster in the tables.
__device__ __forceinline__ uint32_t bs_mix_u64(uint64_t v) {
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  uint32_t h = (lo ^ (hi * 0x9e3779b1u)) * 0x85ebca77u;
  return h ^ (h >> 15);
}

// ---- mbarrier / TMA bulk copy ------------------------------------------------
__device__ __forceinline__ uint32_t bs_smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void bs_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bs_smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void bs_fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bs_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bs_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bs_smem_addr(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bs_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bs_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool bs_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bs_smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bs_mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!bs_mbar_try_wait(bar, parity)) {
  }
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier.
// Requires 16-byte aligned src/dst and size % 16 == 0.  (SASS: UBLKCP)
__device__ __forceinline__ void bs_tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes,
                                               uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          bs_smem_addr(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(bs_smem_addr(bar))
      : "memory");
}
// TMA 1-D bulk copy shared -> global (may target a peer GPU's mapped memory).
__device__ __forceinline__ void bs_tma_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(bs_smem_addr(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bs_tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bs_tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void bs_tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- system-scope flags for cross-GPU signalling ----------------------------
__device__ __forceinline__ void bs_st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long bs_ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long bs_ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long bs_globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Streaming 128-bit global load that does not pollute L1.
__device__ __forceinline__ int4 bs_ldg_stream(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// 128-bit global load that bypasses L1 allocation but stays on the coherent path
// (for data this kernel also writes, or that peers wrote before the launch).
__device__ __forceinline__ int4 bs_ldg_nocache(const int4* p) {
  int4 r;
  asm volatile("ld.global.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}

// ---- L2 cache-policy hints ----------------------------------------------------
__device__ __forceinline__ unsigned long long bs_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long bs_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ long long bs_ld_hint_s64(const long long* p, unsigned long long pol) {
  long long v;
  // .cg: served by L2 only (L1 is not coherent with the atomics other SMs issue)
  asm volatile("ld.global.cg.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
// TMA 1-D bulk load with an L2 cache hint (streaming inputs use evict_first so
// they do not displace the L2-resident hash-table level).
__device__ __forceinline__ void bs_tma_load_1d_hint(void* smem_dst, const void* gsrc, uint32_t bytes,
                                                    uint64_t* bar, unsigned long long pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(bs_smem_addr(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(bs_smem_addr(bar)), "l"(pol)
      : "memory");
}

// Value dtypes / combine ops shared by the combine, segmented-reduce kernels.
enum BsVType { BS_I32 = 0, BS_I64 = 1, BS_F32 = 2, BS_F64 = 3 };
enum BsOp { BS_ADD = 0, BS_MIN = 1, BS_MAX = 2 };
#define BS_EMPTY_KEY ((long long)0x8000000000000000LL)
