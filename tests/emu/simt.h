// A small SIMT emulator for running bigslice_b200's CUDA kernels on the CPU in tests.
//
// Every CUDA thread of a CTA is a ucontext fiber; CTAs run one after the other.
// Fibers switch only at synchronisation points (__syncthreads, warp collectives), so
// between two such points a thread runs alone — data races cannot be observed, but
// everything else can: indexing, barriers, ballots/shuffles, shared memory, atomics,
// grid-stride loops, tail handling.  All emulator state is thread_local: every host thread
// is its own "device", so several ranks of a multi-GPU protocol can run concurrently in one
// process (their kernels spin on each other's flags through ordinary memory).  Built with -fsanitize=address the same run is a
// memcheck of the device code.  `kernel<<<grid, block, smem, stream>>>(args)` is
// rewritten (tests/emu/__init__.py) into `bs_emu_launch(grid, block, smem, [=]{ kernel(args); })`.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string.h>
#include <algorithm>
#include <functional>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static thread_local   // one copy per emulated device (= per host thread)
#define __constant__ static
#define __align__(x) __attribute__((aligned(x)))

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaDevAttrMultiProcessorCount = 16, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
// graphs are not emulated (callers fall back to direct launches); pinned memory is ordinary memory
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
typedef void* cudaGraphNode_t;
enum { cudaStreamCaptureModeThreadLocal = 1, cudaHostAllocPortable = 1, cudaHostAllocMapped = 2,
       cudaErrorNotSupported = 801 };
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return cudaErrorNotSupported; }
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphGetNodes(cudaGraph_t, cudaGraphNode_t*, size_t* n) { *n = 0; return cudaSuccess; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) { *e = nullptr; return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
static inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) {
  *p = aligned_alloc(64, (n + 63) & ~(size_t)63);
  return *p ? cudaSuccess : 2;
}
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
// device memory is host memory; IPC handles carry the pointer itself (one process)
enum { cudaErrorPeerAccessAlreadyEnabled = 704, cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };
static inline cudaError_t cudaMalloc(void** p, size_t n) { return cudaHostAlloc(p, n ? n : 64, 0); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  memset(h, 0, sizeof(*h));
  memcpy(h->reserved, &p, sizeof(p));
  return cudaSuccess;
}
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
  memcpy(p, h.reserved, sizeof(*p));
  return cudaSuccess;
}
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice,
                      cudaMemcpyDefault };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  memcpy(d, s, n);
  return cudaSuccess;
}
#define cudaMemcpyToSymbol(sym, src, n) (memcpy((void*)&(sym), (src), (n)), cudaSuccess)
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 2; return cudaSuccess; }  // "2 SMs"

struct alignas(16) int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
struct bs_dim3 { unsigned x, y, z; };
static thread_local bs_dim3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0}, blockDim{1, 1, 1}, gridDim{1, 1, 1};

// ---- fibers -------------------------------------------------------------------
static thread_local ucontext_t bs_sched;
static thread_local std::vector<ucontext_t> bs_fib;
static thread_local std::vector<char> bs_done;
static thread_local int bs_cur, bs_live, bs_nthreads;
static thread_local std::function<void()> bs_kernel;
static inline void bs_yield() { swapcontext(&bs_fib[bs_cur], &bs_sched); }

static thread_local int bs_bar_gen, bs_bar_count;
static inline void bs_bar_check() {
  if (bs_live > 0 && bs_bar_count == bs_live) { bs_bar_count = 0; ++bs_bar_gen; }
}
static inline void __syncthreads() {
  const int gen = bs_bar_gen;
  ++bs_bar_count;
  bs_bar_check();
  while (bs_bar_gen == gen) bs_yield();
}
// barrier + block-wide reduction of a predicate.  Every thread counts its own reduction
// rounds; round r accumulates in slot r % 4, initialised by its first arrival — a slot is
// reused four rounds later, when every thread has long read the old result.
struct bs_red_t { long long round; int any, all, count; };
static thread_local bs_red_t bs_red[4];
static thread_local std::vector<long long> bs_red_round;   // per CUDA thread
static inline int bs_sync_reduce(int p, int what) {
  const long long r = ++bs_red_round[bs_cur];
  bs_red_t& a = bs_red[r & 3];
  if (a.round != r) a = bs_red_t{r, 0, 1, 0};
  a.any |= (p != 0);
  a.all &= (p != 0);
  a.count += (p != 0);
  __syncthreads();
  return what == 0 ? a.any : what == 1 ? a.all : a.count;
}
static inline int __syncthreads_or(int p) { return bs_sync_reduce(p, 0); }
static inline int __syncthreads_and(int p) { return bs_sync_reduce(p, 1); }
static inline int __syncthreads_count(int p) { return bs_sync_reduce(p, 2); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) { bs_yield(); }

// ---- warp collectives: every live lane of the warp deposits, the last arrival publishes ----
struct bs_warp_t {
  unsigned bits, result_bits;
  unsigned long long val[32], result_val[32];
  int count, gen, lanes, width;  // lanes: still running; width: lanes the warp started with
};
static thread_local std::vector<bs_warp_t> bs_warps;
static inline bs_warp_t& bs_warp() { return bs_warps[bs_cur >> 5]; }
static inline int bs_lane() { return bs_cur & 31; }
static inline void bs_warp_arrive(bs_warp_t& w) {
  const int gen = w.gen;
  if (++w.count == w.lanes) {
    w.result_bits = w.bits;
    std::memcpy(w.result_val, w.val, sizeof(w.val));
    w.bits = 0;
    w.count = 0;
    ++w.gen;
  }
  while (w.gen == gen) bs_yield();
}
static inline unsigned __ballot_sync(unsigned, int p) {
  bs_warp_t& w = bs_warp();
  if (p) w.bits |= 1u << bs_lane();
  bs_warp_arrive(w);
  return w.result_bits;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { bs_warp_arrive(bs_warp()); }  // a real warp barrier
static inline unsigned __activemask() { return 0xffffffffu; }
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, !p) == 0; }
template <class T> static inline unsigned long long bs_bits(T v) {
  unsigned long long b = 0;
  std::memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T> static inline T bs_unbits(unsigned long long b) {
  T v;
  std::memcpy(&v, &b, sizeof(T));
  return v;
}
template <class T> static inline T bs_exchange(T v, int src_lane) {
  bs_warp_t& w = bs_warp();
  w.val[bs_lane()] = bs_bits(v);
  bs_warp_arrive(w);
  if (src_lane < 0 || src_lane >= w.width) return v;
  return bs_unbits<T>(w.result_val[src_lane]);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  return bs_exchange(v, (bs_lane() / width) * width + (src % width));
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  const int l = bs_lane(), s = l + (int)d;
  return bs_exchange(v, (s / width == l / width) ? s : l);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  const int l = bs_lane(), s = l - (int)d;
  return bs_exchange(v, (s >= 0 && s / width == l / width) ? s : l);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  return bs_exchange(v, bs_lane() ^ m);
}
template <class T> static inline unsigned __match_any_sync(unsigned, T v) {
  bs_warp_t& w = bs_warp();
  w.val[bs_lane()] = bs_bits(v);
  bs_warp_arrive(w);
  unsigned m = 0;
  for (int l = 0; l < w.width; ++l)
    if (w.result_val[l] == bs_bits(v)) m |= 1u << l;
  return m;
}

// ---- scalar intrinsics ------------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
  s &= 31;
  return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline float __int_as_float(int v) { return bs_unbits<float>((unsigned)v); }
static inline int __float_as_int(float v) { return (int)bs_bits(v); }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline double __longlong_as_double(long long v) { return bs_unbits<double>((unsigned long long)v); }
static inline long long __double_as_longlong(double v) { return (long long)bs_bits(v); }

// atomics: a thread runs alone between synchronisation points
template <class T, class U> static inline T atomicAdd(T* p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicSub(T* p, U v) { T o = *p; *p = (T)(o - (T)v); return o; }
template <class T, class U> static inline T atomicMax(T* p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMin(T* p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicOr(T* p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicAnd(T* p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U> static inline T atomicExch(T* p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U, class V> static inline T atomicCAS(T* p, U cmp, V v) {
  T o = *p;
  if (o == (T)cmp) *p = (T)v;
  return o;
}

// ---- launch -----------------------------------------------------------------------
static void bs_trampoline() {
  bs_kernel();
  bs_done[bs_cur] = 1;
  --bs_live;
  bs_warp_t& w = bs_warp();
  --w.lanes;  // a lane that has exited no longer takes part in collectives
  if (w.lanes > 0 && w.count == w.lanes) {
    w.result_bits = w.bits;
    std::memcpy(w.result_val, w.val, sizeof(w.val));
    w.bits = 0;
    w.count = 0;
    ++w.gen;
  }
  bs_bar_check();
  swapcontext(&bs_fib[bs_cur], &bs_sched);
}

alignas(128) static thread_local unsigned char bs_dyn_smem[256 << 10];
static thread_local unsigned long long bs_rng = 0;
static thread_local long long bs_sched_seed = -1;   // -1: not read yet
static inline unsigned bs_rand() {                  // xorshift64*
  bs_rng ^= bs_rng >> 12;
  bs_rng ^= bs_rng << 25;
  bs_rng ^= bs_rng >> 27;
  return (unsigned)((bs_rng * 2685821657736338717ULL) >> 33);
}  // dynamic shared memory of the running CTA

static void bs_emu_launch(long long grid, long long block, long long smem_bytes, std::function<void()> k) {
  static thread_local std::vector<std::vector<char>> stacks;
  if (smem_bytes > (long long)sizeof(bs_dyn_smem)) abort();
  if ((long long)stacks.size() < block) stacks.resize(block, std::vector<char>(1 << 17));
  if (bs_sched_seed < 0) {
    const char* e = getenv("BS_EMU_SCHED");
    bs_sched_seed = e ? atoll(e) : 0;
    bs_rng = 0x9E3779B97F4A7C15ULL ^ (unsigned long long)bs_sched_seed;
  }
  bs_kernel = k;
  gridDim = bs_dim3{(unsigned)grid, 1, 1};
  blockDim = bs_dim3{(unsigned)block, 1, 1};
  bs_nthreads = (int)block;
  std::vector<long long> cta_order(grid);
  for (long long b = 0; b < grid; ++b) cta_order[b] = b;
  if (bs_sched_seed)
    for (long long b = grid - 1; b > 0; --b) std::swap(cta_order[b], cta_order[bs_rand() % (unsigned)(b + 1)]);
  for (long long bi = 0; bi < grid; ++bi) {
    const long long b = cta_order[bi];
    blockIdx = bs_dim3{(unsigned)b, 0, 0};
    bs_bar_gen = bs_bar_count = 0;
    bs_red_round.assign(bs_nthreads, 0);
    for (auto& a : bs_red) a = bs_red_t{0, 0, 1, 0};
    bs_live = bs_nthreads;
    bs_warps.assign((bs_nthreads + 31) / 32, bs_warp_t());
    for (size_t w = 0; w < bs_warps.size(); ++w) {
      const int left = bs_nthreads - (int)w * 32;
      bs_warps[w].lanes = bs_warps[w].width = left < 32 ? left : 32;
    }
    bs_fib.assign(bs_nthreads, ucontext_t());
    bs_done.assign(bs_nthreads, 0);
    for (int t = 0; t < bs_nthreads; ++t) {
      getcontext(&bs_fib[t]);
      bs_fib[t].uc_stack.ss_sp = stacks[t].data();
      bs_fib[t].uc_stack.ss_size = stacks[t].size();
      bs_fib[t].uc_link = &bs_sched;
      makecontext(&bs_fib[t], bs_trampoline, 0);
    }
    // Scheduling order.  Default: round robin.  BS_EMU_SCHED=<seed> (non-zero): CTAs are
    // visited in a random order and, on every pass, the runnable threads too — a schedule
    // fuzzer: code that only works because threads happen to run in index order between
    // barriers (an unsynchronised read of something a neighbour writes) gives wrong answers.
    static thread_local std::vector<int> order;
    order.resize(bs_nthreads);
    for (int t = 0; t < bs_nthreads; ++t) order[t] = t;
    int ndone = 0;
    while (ndone < bs_nthreads) {
      if (bs_sched_seed)
        for (int t = bs_nthreads - 1; t > 0; --t) std::swap(order[t], order[bs_rand() % (unsigned)(t + 1)]);
      for (int q = 0; q < bs_nthreads; ++q) {
        const int t = order[q];
        if (bs_done[t]) continue;
        bs_cur = t;
        threadIdx = bs_dim3{(unsigned)t, 0, 0};
        swapcontext(&bs_sched, &bs_fib[t]);
        if (bs_done[t]) ++ndone;
      }
    }
  }
}
