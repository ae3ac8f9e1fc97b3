// K3 (local half): hash aggregation of <key,value> rows for Reduce/Fold with a
// builtin combiner — the kernel behind the wordcount benchmark.
//
//   rows --TMA bulk--> smem stages --> per-CTA hot-key table (shared memory)
//                                   \-> HBM open-addressing table (atomic CAS
//                                       on key + atomic combine on value)
//
// * One persistent CTA per SM.  A dedicated producer warp streams row tiles
//   global->shared with cp.async.bulk (TMA, SASS UBLKCP) through a 3-stage
//   mbarrier ring; 16 consumer warps drain the stages.
// * A direct-mapped shared-memory table absorbs the head of the key
//   distribution (Zipf head keys never leave the SM); misses go to the HBM
//   table.  At the end each CTA flushes its shared table into the HBM table.
// * Values combine with native atomics (add/min/max on i32/i64/f32/f64).
//
// This replaces the reference's per-row combiningFrame probe with a reflective
// combine call (exec/combiner.go:149-174) and the per-partition 8-slot tables
// of worker.runCombine (exec/bigmachine.go:1144-1209).
#include "common.cuh"
#include <type_traits>

struct BsCombineStatus {
  unsigned long long failed;      // rows that found no slot (table too small)
  unsigned long long has_special; // a row carried the reserved EMPTY key
  unsigned long long special_val; // its combined value (bit pattern)
  unsigned long long out_count;   // rows written by compact
};

// ---- typed atomic combine ---------------------------------------------------
template <typename V, int OP>
struct Combine;

template <int OP>
struct Combine<long long, OP> {
  __device__ static __forceinline__ long long identity() {
    return OP == BS_ADD ? 0 : (OP == BS_MIN ? 0x7fffffffffffffffLL : (long long)0x8000000000000000LL);
  }
  __device__ static __forceinline__ void apply(long long* p, long long v) {
    if (OP == BS_ADD) atomicAdd((unsigned long long*)p, (unsigned long long)v);
    else if (OP == BS_MIN) atomicMin(p, v);
    else atomicMax(p, v);
  }
};
template <int OP>
struct Combine<int, OP> {
  __device__ static __forceinline__ int identity() {
    return OP == BS_ADD ? 0 : (OP == BS_MIN ? 0x7fffffff : (int)0x80000000);
  }
  __device__ static __forceinline__ void apply(int* p, int v) {
    if (OP == BS_ADD) atomicAdd(p, v);
    else if (OP == BS_MIN) atomicMin(p, v);
    else atomicMax(p, v);
  }
};
template <int OP>
struct Combine<float, OP> {
  __device__ static __forceinline__ float identity() {
    return OP == BS_ADD ? 0.f : (OP == BS_MIN ? __int_as_float(0x7f800000) : __int_as_float(0xff800000));
  }
  __device__ static __forceinline__ void apply(float* p, float v) {
    if (OP == BS_ADD) {
      atomicAdd(p, v);
    } else {
      int* ip = (int*)p;
      int old = *ip;
      while (true) {
        float f = __int_as_float(old);
        float nv = OP == BS_MIN ? fminf(f, v) : fmaxf(f, v);
        if (nv == f) break;
        int prev = atomicCAS(ip, old, __float_as_int(nv));
        if (prev == old) break;
        old = prev;
      }
    }
  }
};
template <int OP>
struct Combine<double, OP> {
  __device__ static __forceinline__ double identity() {
    return OP == BS_ADD ? 0.0
                        : (OP == BS_MIN ? __longlong_as_double(0x7ff0000000000000LL)
                                        : __longlong_as_double((long long)0xfff0000000000000LL));
  }
  __device__ static __forceinline__ void apply(double* p, double v) {
    if (OP == BS_ADD) {
      atomicAdd(p, v);
    } else {
      unsigned long long* ip = (unsigned long long*)p;
      unsigned long long old = *ip;
      while (true) {
        double f = __longlong_as_double((long long)old);
        double nv = OP == BS_MIN ? fmin(f, v) : fmax(f, v);
        if (nv == f) break;
        unsigned long long prev = atomicCAS(ip, old, (unsigned long long)__double_as_longlong(nv));
        if (prev == old) break;
        old = prev;
      }
    }
  }
};

// ---- table init -------------------------------------------------------------
template <typename V, int OP>
__global__ void __launch_bounds__(256) table_init_kernel(long long* __restrict__ tkeys,
                                                         V* __restrict__ tvals, long long cap) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  V id = Combine<V, OP>::identity();
  for (; i < cap; i += stride) {
    tkeys[i] = BS_EMPTY_KEY;
    tvals[i] = id;
  }
}

// ---- HBM table insert -------------------------------------------------------
template <typename V, int OP>
__device__ __forceinline__ void table_insert(long long* __restrict__ tkeys, V* __restrict__ tvals,
                                             unsigned long long mask, long long key, V val,
                                             uint32_t h, BsCombineStatus* st) {
  unsigned long long slot = (unsigned long long)h & mask;
  // Linear probing; bounded so a full table cannot hang the GPU.
  for (int probe = 0; probe < 4096; ++probe) {
    long long cur = __ldcg(&tkeys[slot]);
    if (cur == key) {
      Combine<V, OP>::apply(&tvals[slot], val);
      return;
    }
    if (cur == BS_EMPTY_KEY) {
      long long prev = (long long)atomicCAS((unsigned long long*)&tkeys[slot],
                                            (unsigned long long)BS_EMPTY_KEY, (unsigned long long)key);
      if (prev == BS_EMPTY_KEY || prev == key) {
        Combine<V, OP>::apply(&tvals[slot], val);
        return;
      }
    }
    slot = (slot + 1) & mask;
  }
  atomicAdd(&st->failed, 1ULL);
}

template <typename V, int OP>
__device__ __forceinline__ void special_insert(V val, BsCombineStatus* st) {
  // The reserved EMPTY key itself: combined in a dedicated status slot.
  if (atomicCAS(&st->has_special, 0ULL, 1ULL) == 0ULL) {
    // first writer initialises to identity before combining
  }
  Combine<V, OP>::apply((V*)&st->special_val, val);
}

// ---- the combine kernel -------------------------------------------------------
#define HC_CONSUMERS 512
#define HC_THREADS (HC_CONSUMERS + 32)
#define HC_TILE 2048
#define HC_STAGES 3
#define HC_SLOTS 4096  // shared hot-key table entries (power of two)

template <typename V, int OP>
__device__ __forceinline__ void combine_row(long long key, V val, long long* __restrict__ skeys,
                                            V* __restrict__ svals, long long* __restrict__ tkeys,
                                            V* __restrict__ tvals, unsigned long long mask,
                                            BsCombineStatus* st) {
  if (key == BS_EMPTY_KEY) {
    special_insert<V, OP>(val, st);
    return;
  }
  uint32_t h = bs_mix_u64((uint64_t)key);
  uint32_t s = (h >> 20) & (HC_SLOTS - 1);  // high bits: independent of the HBM slot bits
  long long k0 = *(volatile long long*)&skeys[s];
  if (k0 != key) {
    if (k0 != BS_EMPTY_KEY) {
      table_insert<V, OP>(tkeys, tvals, mask, key, val, h, st);
      return;
    }
    long long prev = (long long)atomicCAS((unsigned long long*)&skeys[s],
                                          (unsigned long long)BS_EMPTY_KEY, (unsigned long long)key);
    if (prev != BS_EMPTY_KEY && prev != key) {
      table_insert<V, OP>(tkeys, tvals, mask, key, val, h, st);
      return;
    }
  }
  Combine<V, OP>::apply(&svals[s], val);
}

template <typename V, int OP>
__global__ void __launch_bounds__(HC_THREADS, 1) hash_combine_tma_kernel(
    const long long* __restrict__ keys, const V* __restrict__ vals, long long n,
    long long* __restrict__ tkeys, V* __restrict__ tvals, unsigned long long mask,
    BsCombineStatus* __restrict__ st) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  long long* stage_keys = (long long*)smem_raw;                                   // [STAGES][TILE]
  V* stage_vals = (V*)(smem_raw + (size_t)HC_STAGES * HC_TILE * sizeof(long long));  // [STAGES][TILE]
  long long* skeys = (long long*)((unsigned char*)stage_vals + (size_t)HC_STAGES * HC_TILE * sizeof(V));
  V* svals = (V*)(skeys + HC_SLOTS);
  uint64_t* full = (uint64_t*)(svals + HC_SLOTS);
  uint64_t* empty = full + HC_STAGES;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const V ident = Combine<V, OP>::identity();
  for (int s = tid; s < HC_SLOTS; s += HC_THREADS) {
    skeys[s] = BS_EMPTY_KEY;
    svals[s] = ident;
  }
  if (tid == 0) {
    for (int s = 0; s < HC_STAGES; ++s) {
      bs_mbar_init(&full[s], 1);
      bs_mbar_init(&empty[s], HC_CONSUMERS / 32);
    }
    bs_fence_barrier_init();
  }
  __syncthreads();

  const long long ntiles = (n + HC_TILE - 1) / HC_TILE;
  if (warp == HC_CONSUMERS / 32) {
    // ===== producer warp: one elected lane issues the TMA bulk loads =====
    if (lane == 0) {
      long long it = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        int s = (int)(it % HC_STAGES);
        uint32_t ph = (uint32_t)((it / HC_STAGES) & 1);
        bs_mbar_wait(&empty[s], ph ^ 1u);
        long long row0 = tile * HC_TILE;
        long long rows = n - row0 < HC_TILE ? n - row0 : HC_TILE;
        uint32_t kb = (uint32_t)(rows * sizeof(long long)), vb = (uint32_t)(rows * sizeof(V));
        bs_mbar_expect_tx(&full[s], kb + vb);
        bs_tma_load_1d(stage_keys + (size_t)s * HC_TILE, keys + row0, kb, &full[s]);
        bs_tma_load_1d(stage_vals + (size_t)s * HC_TILE, vals + row0, vb, &full[s]);
      }
    }
  } else {
    // ===== consumer warps =====
    long long it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      int s = (int)(it % HC_STAGES);
      uint32_t ph = (uint32_t)((it / HC_STAGES) & 1);
      bs_mbar_wait(&full[s], ph);
      long long row0 = tile * HC_TILE;
      int rows = (int)(n - row0 < HC_TILE ? n - row0 : HC_TILE);
      const long long* sk = stage_keys + (size_t)s * HC_TILE;
      const V* sv = stage_vals + (size_t)s * HC_TILE;
      long long k[HC_TILE / HC_CONSUMERS];
      V v[HC_TILE / HC_CONSUMERS];
#pragma unroll
      for (int j = 0; j < HC_TILE / HC_CONSUMERS; ++j) {
        int idx = tid + j * HC_CONSUMERS;
        if (idx < rows) {
          k[j] = sk[idx];
          v[j] = sv[idx];
        }
      }
      // stage drained into registers: hand the slot back before the atomics
      __syncwarp();
      if (lane == 0) bs_mbar_arrive(&empty[s]);
#pragma unroll
      for (int j = 0; j < HC_TILE / HC_CONSUMERS; ++j) {
        int idx = tid + j * HC_CONSUMERS;
        if (idx < rows) combine_row<V, OP>(k[j], v[j], skeys, svals, tkeys, tvals, mask, st);
      }
    }
  }
  __syncthreads();
  // ===== flush the shared hot-key table into the HBM table =====
  for (int s = tid; s < HC_SLOTS; s += HC_THREADS) {
    long long key = skeys[s];
    if (key != BS_EMPTY_KEY)
      table_insert<V, OP>(tkeys, tvals, mask, key, svals[s], bs_mix_u64((uint64_t)key), st);
  }
}

// Plain (non-TMA) variant: unaligned inputs and tiny batches; also the
// cross-check for the TMA pipeline in the numerics tests.
template <typename V, int OP>
__global__ void __launch_bounds__(512) hash_combine_simple_kernel(
    const long long* __restrict__ keys, const V* __restrict__ vals, long long n,
    long long* __restrict__ tkeys, V* __restrict__ tvals, unsigned long long mask,
    BsCombineStatus* __restrict__ st) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    long long key = keys[i];
    V val = vals[i];
    if (key == BS_EMPTY_KEY) {
      special_insert<V, OP>(val, st);
      continue;
    }
    table_insert<V, OP>(tkeys, tvals, mask, key, val, bs_mix_u64((uint64_t)key), st);
  }
}

// ---- compaction: table -> dense <key, value> ---------------------------------
template <typename V>
__global__ void __launch_bounds__(256) table_compact_kernel(const long long* __restrict__ tkeys,
                                                            const V* __restrict__ tvals,
                                                            long long cap,
                                                            long long* __restrict__ out_keys,
                                                            V* __restrict__ out_vals,
                                                            BsCombineStatus* __restrict__ st) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  // grid-stride over whole warps so the ballot is always convergent
  long long limit = ((cap + 31) / 32) * 32;
  for (; i < limit; i += stride) {
    long long key = i < cap ? tkeys[i] : BS_EMPTY_KEY;
    bool occ = key != BS_EMPTY_KEY;
    unsigned m = __ballot_sync(0xffffffffu, occ);
    if (m == 0) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&st->out_count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (occ) {
      unsigned long long o = base + __popc(m & ((1u << lane) - 1));
      out_keys[o] = key;
      out_vals[o] = tvals[i];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && st->has_special) {
    unsigned long long o = atomicAdd(&st->out_count, 1ULL);
    out_keys[o] = BS_EMPTY_KEY;
    out_vals[o] = *(const V*)&st->special_val;
  }
}

// ---- host wrappers --------------------------------------------------------------
template <typename V>
static size_t hc_smem_bytes() {
  return (size_t)HC_STAGES * HC_TILE * (sizeof(long long) + sizeof(V)) +
         (size_t)HC_SLOTS * (sizeof(long long) + sizeof(V)) + 2 * HC_STAGES * sizeof(uint64_t) + 16;
}

template <typename V, int OP>
static int launch_combine(const long long* keys, const V* vals, long long n, long long* tkeys,
                          V* tvals, long long cap, BsCombineStatus* st, int force_simple,
                          cudaStream_t stream) {
  if (n <= 0) return 0;
  unsigned long long mask = (unsigned long long)cap - 1;
  int sms = bs_num_sms();
  bool aligned = (((uintptr_t)keys | (uintptr_t)vals) & 15) == 0;
  // TMA sizes must be multiples of 16 bytes: peel a ragged tail.
  long long per16 = 16 / (long long)sizeof(V);
  long long n_main = aligned ? (n / per16) * per16 : 0;
  if (force_simple || n_main < (long long)HC_TILE * 4) n_main = 0;
  if (n_main > 0) {
    size_t smem = hc_smem_bytes<V>();
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(hash_combine_tma_kernel<V, OP>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      attr_set = true;
    }
    long long ntiles = (n_main + HC_TILE - 1) / HC_TILE;
    int blocks = (int)(ntiles < sms ? ntiles : sms);
    hash_combine_tma_kernel<V, OP><<<blocks, HC_THREADS, smem, stream>>>(keys, vals, n_main, tkeys,
                                                                         tvals, mask, st);
    BS_LAUNCH_CHECK();
  }
  long long rest = n - n_main;
  if (rest > 0) {
    long long want = (rest + 511) / 512;
    int blocks = (int)(want < (long long)sms * 4 ? want : (long long)sms * 4);
    hash_combine_simple_kernel<V, OP><<<blocks, 512, 0, stream>>>(keys + n_main, vals + n_main, rest,
                                                                  tkeys, tvals, mask, st);
    BS_LAUNCH_CHECK();
  }
  return 0;
}

#define DISPATCH_VOP(FN, ...)                                                          \
  switch (vtype * 3 + op) {                                                            \
    case BS_I32 * 3 + BS_ADD: return FN<int, BS_ADD>(__VA_ARGS__);                     \
    case BS_I32 * 3 + BS_MIN: return FN<int, BS_MIN>(__VA_ARGS__);                     \
    case BS_I32 * 3 + BS_MAX: return FN<int, BS_MAX>(__VA_ARGS__);                     \
    case BS_I64 * 3 + BS_ADD: return FN<long long, BS_ADD>(__VA_ARGS__);               \
    case BS_I64 * 3 + BS_MIN: return FN<long long, BS_MIN>(__VA_ARGS__);               \
    case BS_I64 * 3 + BS_MAX: return FN<long long, BS_MAX>(__VA_ARGS__);               \
    case BS_F32 * 3 + BS_ADD: return FN<float, BS_ADD>(__VA_ARGS__);                   \
    case BS_F32 * 3 + BS_MIN: return FN<float, BS_MIN>(__VA_ARGS__);                   \
    case BS_F32 * 3 + BS_MAX: return FN<float, BS_MAX>(__VA_ARGS__);                   \
    case BS_F64 * 3 + BS_ADD: return FN<double, BS_ADD>(__VA_ARGS__);                  \
    case BS_F64 * 3 + BS_MIN: return FN<double, BS_MIN>(__VA_ARGS__);                  \
    case BS_F64 * 3 + BS_MAX: return FN<double, BS_MAX>(__VA_ARGS__);                  \
    default: return -1;                                                                \
  }

template <typename V, int OP>
static int launch_init(long long* tkeys, void* tvals, long long cap, BsCombineStatus* st,
                       cudaStream_t stream) {
  int blocks = (int)((cap + 255) / 256 < (long long)bs_num_sms() * 8 ? (cap + 255) / 256
                                                                    : (long long)bs_num_sms() * 8);
  table_init_kernel<V, OP><<<blocks, 256, 0, stream>>>(tkeys, (V*)tvals, cap);
  BS_LAUNCH_CHECK();
  BsCombineStatus z;
  z.failed = 0;
  z.has_special = 0;
  z.out_count = 0;
  V id = OP == BS_ADD ? (V)0 : (V)0;
  (void)id;
  // special_val starts at the identity bit pattern
  unsigned long long bits = 0;
  if (OP == BS_MIN) {
    if (sizeof(V) == 8) bits = std::is_same<V, double>::value ? 0x7ff0000000000000ULL : 0x7fffffffffffffffULL;
    else bits = std::is_same<V, float>::value ? 0x7f800000ULL : 0x7fffffffULL;
  } else if (OP == BS_MAX) {
    if (sizeof(V) == 8) bits = std::is_same<V, double>::value ? 0xfff0000000000000ULL : 0x8000000000000000ULL;
    else bits = std::is_same<V, float>::value ? 0xff800000ULL : 0x80000000ULL;
  }
  z.special_val = bits;
  cudaError_t e = cudaMemcpyAsync(st, &z, sizeof(z), cudaMemcpyHostToDevice, stream);
  return (int)e;
}

BS_EXPORT int bs_table_init(long long* tkeys, void* tvals, long long cap, int vtype, int op,
                            void* status, cudaStream_t stream) {
  if (cap <= 0 || (cap & (cap - 1))) return -2;
  DISPATCH_VOP(launch_init, tkeys, tvals, cap, (BsCombineStatus*)status, stream)
}

template <typename V, int OP>
static int combine_entry(const long long* keys, const void* vals, long long n, long long* tkeys,
                         void* tvals, long long cap, void* st, int force_simple,
                         cudaStream_t stream) {
  return launch_combine<V, OP>(keys, (const V*)vals, n, tkeys, (V*)tvals, cap, (BsCombineStatus*)st,
                               force_simple, stream);
}

BS_EXPORT int bs_hash_combine(const long long* keys, const void* vals, long long n, int vtype, int op,
                              long long* tkeys, void* tvals, long long cap, void* status,
                              int force_simple, cudaStream_t stream) {
  if (cap <= 0 || (cap & (cap - 1))) return -2;
  DISPATCH_VOP(combine_entry, keys, vals, n, tkeys, tvals, cap, status, force_simple, stream)
}

BS_EXPORT int bs_table_compact(const long long* tkeys, const void* tvals, long long cap, int vtype,
                               long long* out_keys, void* out_vals, void* status,
                               cudaStream_t stream) {
  int blocks = (int)((cap + 255) / 256 < (long long)bs_num_sms() * 8 ? (cap + 255) / 256
                                                                    : (long long)bs_num_sms() * 8);
  BsCombineStatus* st = (BsCombineStatus*)status;
  switch (vtype) {
    case BS_I32: table_compact_kernel<int><<<blocks, 256, 0, stream>>>(tkeys, (const int*)tvals, cap, out_keys, (int*)out_vals, st); break;
    case BS_I64: table_compact_kernel<long long><<<blocks, 256, 0, stream>>>(tkeys, (const long long*)tvals, cap, out_keys, (long long*)out_vals, st); break;
    case BS_F32: table_compact_kernel<float><<<blocks, 256, 0, stream>>>(tkeys, (const float*)tvals, cap, out_keys, (float*)out_vals, st); break;
    case BS_F64: table_compact_kernel<double><<<blocks, 256, 0, stream>>>(tkeys, (const double*)tvals, cap, out_keys, (double*)out_vals, st); break;
    default: return -1;
  }
  BS_LAUNCH_CHECK();
  return 0;
}

BS_EXPORT int bs_combine_status_size() { return (int)sizeof(BsCombineStatus); }
