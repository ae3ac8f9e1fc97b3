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
#include "table.cuh"
#include <stdlib.h>

// ---- shared-memory accumulators ------------------------------------------------
// 64-bit shared atomics compile to CAS spin loops (ATOMS.CAST.SPIN.64); for the
// common int64 sum the accumulator is split into two 32-bit words combined with
// native ATOMS.ADD and a carry.
template <typename V, int OP>
struct SmemAcc {
  __device__ static __forceinline__ void init(void* base, int s) { ((V*)base)[s] = Combine<V, OP>::identity(); }
  __device__ static __forceinline__ void apply(void* base, int s, V v) { Combine<V, OP>::apply(&((V*)base)[s], v); }
  __device__ static __forceinline__ V read(void* base, int s) { return ((V*)base)[s]; }
};
#define HC_SLOTS 4096  // shared hot-key table entries (power of two)
template <>
struct SmemAcc<long long, BS_ADD> {
  __device__ static __forceinline__ void init(void* base, int s) {
    ((unsigned int*)base)[s] = 0;
    ((unsigned int*)base)[HC_SLOTS + s] = 0;
  }
  __device__ static __forceinline__ void apply(void* base, int s, long long v) {
    unsigned int lo = (unsigned int)v, hi = (unsigned int)((unsigned long long)v >> 32);
    unsigned int old = atomicAdd(&((unsigned int*)base)[s], lo);
    hi += ((unsigned int)(old + lo) < old) ? 1u : 0u;  // carry out of the low word
    if (hi) atomicAdd(&((unsigned int*)base)[HC_SLOTS + s], hi);
  }
  __device__ static __forceinline__ long long read(void* base, int s) {
    return (long long)(((unsigned long long)((unsigned int*)base)[HC_SLOTS + s] << 32) |
                       ((unsigned int*)base)[s]);
  }
};

// ---- the combine kernel -------------------------------------------------------
#define HC_CONSUMERS 992  // 31 consumer warps + 1 producer warp = 1024 threads (the CTA limit)
#define HC_THREADS (HC_CONSUMERS + 32)
static_assert(HC_THREADS <= 1024, "a CTA has at most 1024 threads");
#define HC_ROWS 4
#define HC_TILE (HC_CONSUMERS * HC_ROWS)
#define HC_STAGES 2
#define HC_DEST 256  // max destination shards tracked (XC_MAXP of the exchange)
static_assert((HC_TILE * 4) % 16 == 0, "TMA sizes must be multiples of 16 bytes");

template <typename V, int OP>
__global__ void __launch_bounds__(HC_THREADS, 1) hash_combine_tma_kernel(
    const long long* __restrict__ keys, const V* __restrict__ vals, long long n,
    BsTable tab, BsCombineStatus* __restrict__ st) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  long long* stage_keys = (long long*)smem_raw;                                   // [STAGES][TILE]
  V* stage_vals = (V*)(smem_raw + (size_t)HC_STAGES * HC_TILE * sizeof(long long));  // [STAGES][TILE]
  long long* skeys = (long long*)((unsigned char*)stage_vals + (size_t)HC_STAGES * HC_TILE * sizeof(V));
  void* sacc = (void*)(skeys + HC_SLOTS);  // HC_SLOTS * 8 bytes
  uint64_t* full = (uint64_t*)((unsigned char*)sacc + (size_t)HC_SLOTS * 8);
  uint64_t* empty = full + HC_STAGES;
  int* dhist = (int*)(empty + HC_STAGES);  // [HC_DEST] destination histogram of new keys

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  for (int s = tid; s < HC_SLOTS; s += HC_THREADS) {
    skeys[s] = BS_EMPTY_KEY;
    SmemAcc<V, OP>::init(sacc, s);
  }
  if (tid < HC_DEST) dhist[tid] = 0;
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
      const unsigned long long stream_pol = bs_policy_evict_first();
      long long it = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        int s = (int)(it % HC_STAGES);
        uint32_t ph = (uint32_t)((it / HC_STAGES) & 1);
        bs_mbar_wait(&empty[s], ph ^ 1u);
        long long row0 = tile * HC_TILE;
        long long rows = n - row0 < HC_TILE ? n - row0 : HC_TILE;
        uint32_t kb = (uint32_t)(rows * sizeof(long long)), vb = (uint32_t)(rows * sizeof(V));
        bs_mbar_expect_tx(&full[s], kb + vb);
        bs_tma_load_1d_hint(stage_keys + (size_t)s * HC_TILE, keys + row0, kb, &full[s], stream_pol);
        bs_tma_load_1d_hint(stage_vals + (size_t)s * HC_TILE, vals + row0, vb, &full[s], stream_pol);
      }
    }
  } else {
    // ===== consumer warps =====
    const unsigned long long keep_pol = bs_policy_evict_last();
    long long it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      int s = (int)(it % HC_STAGES);
      uint32_t ph = (uint32_t)((it / HC_STAGES) & 1);
      bs_mbar_wait(&full[s], ph);
      long long row0 = tile * HC_TILE;
      int rows = (int)(n - row0 < HC_TILE ? n - row0 : HC_TILE);
      const long long* sk = stage_keys + (size_t)s * HC_TILE;
      const V* sv = stage_vals + (size_t)s * HC_TILE;
      long long k[HC_ROWS];
      V v[HC_ROWS];
#pragma unroll
      for (int j = 0; j < HC_ROWS; ++j) {
        int idx = tid + j * HC_CONSUMERS;
        k[j] = idx < rows ? sk[idx] : 0;
        v[j] = idx < rows ? sv[idx] : Combine<V, OP>::identity();
      }
#ifndef HC_DBG_LATE_ARRIVE
      // Stage drained into registers: hand the slot back before the atomics.
      // The arrive must not become visible while any lane's LDS is still in
      // flight (the refilling TMA write is an async-proxy access that the
      // mbarrier release does not order against outstanding generic-proxy
      // reads; measured: rows of the *next* tile were read when the LSU queue
      // was backed up behind global atomics).  The arriving lane is therefore
      // elected from a ballot over a predicate computed from every loaded
      // register: the vote cannot issue before all eight loads of all lanes
      // have landed, and the arrive depends on the vote.
      {
        unsigned long long x = 0;
#pragma unroll
        for (int j = 0; j < HC_ROWS; ++j) {
          unsigned long long vb = 0;
          memcpy(&vb, &v[j], sizeof(V));
          x ^= (unsigned long long)k[j] ^ vb;
        }
        unsigned b = __ballot_sync(0xffffffffu, (unsigned)(x ^ (x >> 32)) & 1u);
        int leader = b ? __ffs((int)b) - 1 : 0;
        if (lane == leader) bs_mbar_arrive(&empty[s]);
      }
#endif
      // --- phase A: shared-memory hot table ---
      uint32_t h[HC_ROWS];
      unsigned miss = 0;
#pragma unroll
      for (int j = 0; j < HC_ROWS; ++j) {
        int idx = tid + j * HC_CONSUMERS;
        if (idx >= rows) continue;
        if (k[j] == BS_EMPTY_KEY) {
          special_insert<V, OP>(v[j], st);
          continue;
        }
        h[j] = bs_mix_u64((uint64_t)k[j]);
        uint32_t sl = (h[j] >> 20) & (HC_SLOTS - 1);  // high bits: independent of the HBM slot bits
        long long k0 = *(volatile long long*)&skeys[sl];
        if (k0 != k[j]) {
          if (k0 != BS_EMPTY_KEY) {
            miss |= 1u << j;
            continue;
          }
          long long prev = (long long)atomicCAS((unsigned long long*)&skeys[sl],
                                                (unsigned long long)BS_EMPTY_KEY,
                                                (unsigned long long)k[j]);
          if (prev != BS_EMPTY_KEY && prev != k[j]) {
            miss |= 1u << j;
            continue;
          }
        }
        SmemAcc<V, OP>::apply(sacc, (int)sl, v[j]);
      }
      if (tab.nparts == -1) miss = 0;  // profiling aid: shared-memory phase only (rows dropped)
      // --- phase B: warm level (L2-resident): issue all probes back to back ---
      long long cur[HC_ROWS];
      if (tab.has_warm) {
#pragma unroll
        for (int j = 0; j < HC_ROWS; ++j)
          cur[j] = (miss >> j) & 1u
                       ? bs_ld_hint_s64(&tab.warm[(unsigned long long)h[j] & tab.warm_mask].key, keep_pol)
                       : 0;
        // Slots seen EMPTY are claimed with one CAS per row, all issued before any
        // result is consumed (a claim is a dependent round trip to L2; first-touch
        // claims are ~8 % of the rows of every step).
        long long prev[HC_ROWS];
        unsigned claim = 0;
#pragma unroll
        for (int j = 0; j < HC_ROWS; ++j) {
          prev[j] = cur[j];
          if (((miss >> j) & 1u) && cur[j] == BS_EMPTY_KEY) {
            claim |= 1u << j;
            prev[j] = (long long)atomicCAS(
                (unsigned long long*)&tab.warm[(unsigned long long)h[j] & tab.warm_mask].key,
                (unsigned long long)BS_EMPTY_KEY, (unsigned long long)k[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < HC_ROWS; ++j) {
          if (!((miss >> j) & 1u)) continue;
          const bool won = ((claim >> j) & 1u) && prev[j] == BS_EMPTY_KEY;
          if (won) count_claim(tab, dhist, k[j]);
          if (won || prev[j] == k[j]) {
            Combine<V, OP>::apply(slot_val<V>(&tab.warm[(unsigned long long)h[j] & tab.warm_mask]), v[j]);
            miss &= ~(1u << j);
          }
        }
      }
      if (tab.nparts == -2) miss = 0;
      // --- phase C: cold tail -> main level (DRAM): batched probes again ---
      // (probing with the claiming CAS itself instead of load -> CAS was measured
      // slower: 2.09 vs 1.95 ms — a CAS on an occupied slot costs more than a load)
#pragma unroll
      for (int j = 0; j < HC_ROWS; ++j) {
        if ((miss >> j) & 1u) {
          h[j] = bs_hash_main(h[j]);
          cur[j] = __ldcg(&tab.main[bs_main_slot(tab, h[j])].key);
        }
      }
      {
        long long prev[HC_ROWS];
        unsigned claim = 0;
#pragma unroll
        for (int j = 0; j < HC_ROWS; ++j) {
          prev[j] = cur[j];
          if (((miss >> j) & 1u) && cur[j] == BS_EMPTY_KEY) {
            claim |= 1u << j;
            prev[j] = (long long)atomicCAS(
                (unsigned long long*)&tab.main[bs_main_slot(tab, h[j])].key,
                (unsigned long long)BS_EMPTY_KEY, (unsigned long long)k[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < HC_ROWS; ++j) {
          if (!((miss >> j) & 1u)) continue;
          const bool won = ((claim >> j) & 1u) && prev[j] == BS_EMPTY_KEY;
          if (won) count_claim(tab, dhist, k[j]);
          if (won || prev[j] == k[j]) {
            Combine<V, OP>::apply(slot_val<V>(&tab.main[bs_main_slot(tab, h[j])]), v[j]);
          } else {
            // first slot belongs to another key: continue the probe sequence
            main_resolve<V, OP>(tab, k[j], v[j], h[j], prev[j], st, dhist);
          }
        }
      }
#ifdef HC_DBG_LATE_ARRIVE
      __syncwarp();
      if (lane == 0) bs_mbar_arrive(&empty[s]);
#endif
    }
  }
  __syncthreads();
  // ===== flush the shared hot-key table into the HBM table =====
  for (int s = tid; s < HC_SLOTS; s += HC_THREADS) {
    long long key = skeys[s];
    if (key != BS_EMPTY_KEY)
      table_insert<V, OP>(tab, key, SmemAcc<V, OP>::read(sacc, s), bs_mix_u64((uint64_t)key), st, dhist);
  }
  if (tab.nparts > 0) {
    __syncthreads();
    if (tid < tab.nparts && dhist[tid]) atomicAdd(&tab.dest_counts[tid], (unsigned long long)dhist[tid]);
  }
}

// Plain (non-TMA) variant: unaligned inputs and tiny batches; also the
// cross-check for the TMA pipeline in the numerics tests.
template <typename V, int OP>
__global__ void __launch_bounds__(512) hash_combine_simple_kernel(
    const long long* __restrict__ keys, const V* __restrict__ vals, long long n,
    BsTable tab, BsCombineStatus* __restrict__ st) {
  __shared__ int dhist[HC_DEST];
  if (tab.nparts > 0) {
    for (int d = threadIdx.x; d < HC_DEST; d += blockDim.x) dhist[d] = 0;
    __syncthreads();
  }
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    long long key = keys[i];
    V val = vals[i];
    if (key == BS_EMPTY_KEY) {
      special_insert<V, OP>(val, st);
      continue;
    }
    table_insert<V, OP>(tab, key, val, bs_mix_u64((uint64_t)key), st, dhist);
  }
  if (tab.nparts > 0) {
    __syncthreads();
    for (int d = threadIdx.x; d < tab.nparts; d += blockDim.x)
      if (dhist[d]) atomicAdd(&tab.dest_counts[d], (unsigned long long)dhist[d]);
  }
}

// ---- compaction: table -> dense <key, value> ---------------------------------
// One global atomic per 2048-slot chunk (block-aggregated), 16-byte coalesced
// slot loads.  Output order is unspecified (as after any shuffle).
#define TC_THREADS 256
#define TC_ITEMS 8
template <typename V>
__global__ void __launch_bounds__(TC_THREADS) table_compact_kernel(BsSlot* __restrict__ tab,
                                                                   long long cap,
                                                                   long long* __restrict__ out_keys,
                                                                   V* __restrict__ out_vals,
                                                                   BsCombineStatus* __restrict__ st,
                                                                   int clear,
                                                                   unsigned long long clear_bits,
                                                                   long long out_cap) {
  __shared__ unsigned int wcnt[TC_ITEMS][TC_THREADS / 32];
  __shared__ unsigned long long chunk_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long chunk = (long long)TC_THREADS * TC_ITEMS;
  const long long nchunks = (cap + chunk - 1) / chunk;
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    BsSlot sl[TC_ITEMS];
    unsigned bal[TC_ITEMS];
#pragma unroll
    for (int j = 0; j < TC_ITEMS; ++j) {
      long long i = c * chunk + (long long)j * TC_THREADS + threadIdx.x;
      if (i < cap) {
        const int4 raw = *(const int4*)&tab[i];
        memcpy(&sl[j], &raw, 16);
        // scan-and-clear: leave the table ready for its next use, so a pooled
        // table never needs a separate init pass
        if (clear && sl[j].key != BS_EMPTY_KEY) {
          BsSlot z;
          z.key = BS_EMPTY_KEY;
          z.val = clear_bits;
          tab[i] = z;
        }
      } else {
        sl[j].key = BS_EMPTY_KEY;
      }
      bal[j] = __ballot_sync(0xffffffffu, sl[j].key != BS_EMPTY_KEY);
      if (lane == 0) wcnt[j][warp] = __popc(bal[j]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned run = 0;
      for (int j = 0; j < TC_ITEMS; ++j)
        for (int w = 0; w < TC_THREADS / 32; ++w) {
          unsigned t = wcnt[j][w];
          wcnt[j][w] = run;
          run += t;
        }
      chunk_base = run ? atomicAdd(&st->out_count, (unsigned long long)run) : 0ULL;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TC_ITEMS; ++j) {
      if (sl[j].key != BS_EMPTY_KEY) {
        unsigned long long o = chunk_base + wcnt[j][warp] + __popc(bal[j] & ((1u << lane) - 1u));
        if ((long long)o < out_cap) {  // out_count still reports the rows needed
          out_keys[o] = sl[j].key;
          V vv;
          memcpy(&vv, &sl[j].val, sizeof(V));
          out_vals[o] = vv;
        }
      }
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && st->has_special) {
    unsigned long long o = atomicAdd(&st->out_count, 1ULL);
    if ((long long)o < out_cap) {
      out_keys[o] = BS_EMPTY_KEY;
      out_vals[o] = *(const V*)&st->special_val;
    }
  }
}

// ---- host wrappers --------------------------------------------------------------
template <typename V>
static size_t hc_smem_bytes() {
  return (size_t)HC_STAGES * HC_TILE * (sizeof(long long) + sizeof(V)) +
         (size_t)HC_SLOTS * (sizeof(long long) + 8) + 2 * HC_STAGES * sizeof(uint64_t) + HC_DEST * sizeof(int) + 16;
}

template <typename V, int OP>
static int launch_combine(const long long* keys, const V* vals, long long n, BsSlot* slots,
                          long long warm_cap, long long cap, BsCombineStatus* st, int force_simple,
                          int nparts, unsigned long long* dest_counts, cudaStream_t stream) {
  if (n <= 0) return 0;
  BsTable tab;
  tab.warm = slots;
  tab.warm_mask = warm_cap > 0 ? (unsigned long long)warm_cap - 1 : 0;
  tab.has_warm = warm_cap > 0;
  tab.main = slots + warm_cap;
  tab.main_cap = (unsigned long long)cap;
  tab.nparts = dest_counts != nullptr ? nparts : 0;
  tab.dest_counts = dest_counts;
  {
    static int dbg = -2;
    if (dbg == -2) {
      const char* e = getenv("BIGSLICE_HC_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    if (dbg == 1) tab.nparts = -1;  // phase A only
    if (dbg == 2) tab.nparts = -2;  // phases A+B only (cold rows dropped)
  }
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
    hash_combine_tma_kernel<V, OP><<<blocks, HC_THREADS, smem, stream>>>(keys, vals, n_main, tab, st);
    BS_LAUNCH_CHECK();
  }
  long long rest = n - n_main;
  if (rest > 0) {
    long long want = (rest + 511) / 512;
    int blocks = (int)(want < (long long)sms * 4 ? want : (long long)sms * 4);
    hash_combine_simple_kernel<V, OP><<<blocks, 512, 0, stream>>>(keys + n_main, vals + n_main, rest,
                                                                  tab, st);
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

// Bit pattern of the combine identity (low sizeof(V) bytes), as stored in empty slots.
BS_EXPORT unsigned long long bs_identity_bits(int vtype, int op) {
  const bool wide = vtype == BS_I64 || vtype == BS_F64;
  const bool fp = vtype == BS_F32 || vtype == BS_F64;
  if (op == BS_MIN) return wide ? (fp ? 0x7ff0000000000000ULL : 0x7fffffffffffffffULL)
                                : (fp ? 0x7f800000ULL : 0x7fffffffULL);
  if (op == BS_MAX) return wide ? (fp ? 0xfff0000000000000ULL : 0x8000000000000000ULL)
                                : (fp ? 0xff800000ULL : 0x80000000ULL);
  return 0ULL;
}

// Reset the status block of a (clean) table before it is reused.
BS_EXPORT int bs_table_status_reset(void* status, int vtype, int op, cudaStream_t stream) {
  BsCombineStatus z;
  z.failed = 0;
  z.has_special = 0;
  z.out_count = 0;
  z.special_val = bs_identity_bits(vtype, op);
  return (int)cudaMemcpyAsync(status, &z, sizeof(z), cudaMemcpyHostToDevice, stream);
}

template <typename V, int OP>
static int launch_init(void* tab, long long cap, BsCombineStatus* st, cudaStream_t stream) {
  int blocks = (int)((cap + 255) / 256 < (long long)bs_num_sms() * 8 ? (cap + 255) / 256
                                                                    : (long long)bs_num_sms() * 8);
  table_init_kernel<V, OP><<<blocks, 256, 0, stream>>>((BsSlot*)tab, cap);
  BS_LAUNCH_CHECK();
  BsCombineStatus z;
  z.failed = 0;
  z.has_special = 0;
  z.out_count = 0;
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

// `tab` is `total` x 16 bytes (BsSlot): the warm level followed by the main level.
BS_EXPORT int bs_table_init(void* tab, long long total, int vtype, int op, void* status,
                            cudaStream_t stream) {
  if (total <= 0) return -2;
  DISPATCH_VOP(launch_init, tab, total, (BsCombineStatus*)status, stream)
}

template <typename V, int OP>
static int combine_entry(const long long* keys, const void* vals, long long n, void* tab,
                         long long warm_cap, long long cap, void* st, int force_simple,
                         int nparts, unsigned long long* dest_counts, cudaStream_t stream) {
  return launch_combine<V, OP>(keys, (const V*)vals, n, (BsSlot*)tab, warm_cap, cap,
                               (BsCombineStatus*)st, force_simple, nparts, dest_counts, stream);
}

// warm_cap: 0 or a power of two; cap: any size below 2^32.  tab holds warm_cap + cap slots.
// nparts/dest_counts (optional): accumulate, per destination shard
// murmur3(key) % nparts, the number of keys inserted for the first time.
BS_EXPORT int bs_hash_combine(const long long* keys, const void* vals, long long n, int vtype, int op,
                              void* tab, long long warm_cap, long long cap, void* status,
                              int force_simple, int nparts, unsigned long long* dest_counts,
                              cudaStream_t stream) {
  if (cap <= 0 || cap >= (1LL << 32) || (warm_cap & (warm_cap - 1)) || warm_cap < 0) return -2;
  if (dest_counts != nullptr && (nparts < 1 || nparts > HC_DEST)) return -2;
  DISPATCH_VOP(combine_entry, keys, vals, n, tab, warm_cap, cap, status, force_simple, nparts,
               dest_counts, stream)
}

BS_EXPORT int bs_table_compact(void* tab, long long cap, int vtype, long long* out_keys,
                               void* out_vals, void* status, int clear, unsigned long long clear_bits,
                               long long out_cap, cudaStream_t stream) {
  long long nchunks = (cap + (long long)TC_THREADS * TC_ITEMS - 1) / ((long long)TC_THREADS * TC_ITEMS);
  int blocks = (int)(nchunks < (long long)bs_num_sms() * 8 ? nchunks : (long long)bs_num_sms() * 8);
  BsCombineStatus* st = (BsCombineStatus*)status;
  BsSlot* t = (BsSlot*)tab;
  switch (vtype) {
    case BS_I32: table_compact_kernel<int><<<blocks, TC_THREADS, 0, stream>>>(t, cap, out_keys, (int*)out_vals, st, clear, clear_bits, out_cap); break;
    case BS_I64: table_compact_kernel<long long><<<blocks, TC_THREADS, 0, stream>>>(t, cap, out_keys, (long long*)out_vals, st, clear, clear_bits, out_cap); break;
    case BS_F32: table_compact_kernel<float><<<blocks, TC_THREADS, 0, stream>>>(t, cap, out_keys, (float*)out_vals, st, clear, clear_bits, out_cap); break;
    case BS_F64: table_compact_kernel<double><<<blocks, TC_THREADS, 0, stream>>>(t, cap, out_keys, (double*)out_vals, st, clear, clear_bits, out_cap); break;
    default: return -1;
  }
  BS_LAUNCH_CHECK();
  return 0;
}

BS_EXPORT int bs_combine_status_size() { return (int)sizeof(BsCombineStatus); }
