// K9 stream compaction (Filter), K11/K4 segmented reduce over sorted keys,
// K10 expansion indices (Flatmap), K12 CRC32 of device buffers.
#include "common.cuh"
#include <limits>
#include <type_traits>
#include <string.h>

// ---------------------------------------------------------------------------
// compact_indices: out[0..count) = indices i with mask[i] != 0, in order.
// Three small phases in two launches: per-CTA counts -> single-CTA scan ->
// ordered write with warp ballots (stable, unlike an atomic cursor).
#define CP_THREADS 512
#define CP_ITEMS 8
#define CP_TILE (CP_THREADS * CP_ITEMS)

__global__ void __launch_bounds__(CP_THREADS) compact_count_kernel(const unsigned char* __restrict__ mask,
                                                                    long long n,
                                                                    unsigned long long* __restrict__ tile_counts) {
  __shared__ unsigned int wsum[CP_THREADS / 32];
  long long t0 = (long long)blockIdx.x * CP_TILE;
  unsigned c = 0;
#pragma unroll
  for (int j = 0; j < CP_ITEMS; ++j) {
    long long i = t0 + (long long)j * CP_THREADS + threadIdx.x;
    c += (i < n && mask[i]) ? 1u : 0u;
  }
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned s = 0;
    for (int w = 0; w < CP_THREADS / 32; ++w) s += wsum[w];
    tile_counts[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(1024) scan_u64_kernel(unsigned long long* __restrict__ a, long long m,
                                                        unsigned long long* __restrict__ total) {
  __shared__ unsigned long long part[1024];
  const int tid = threadIdx.x;
  long long per = (m + 1023) / 1024;
  long long beg = (long long)tid * per, end = beg + per < m ? beg + per : m;
  unsigned long long s = 0;
  for (long long i = beg; i < end; ++i) s += a[i];
  part[tid] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    unsigned long long t = tid >= o ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += t;
    __syncthreads();
  }
  unsigned long long run = part[tid] - s;
  for (long long i = beg; i < end; ++i) {
    unsigned long long v = a[i];
    a[i] = run;
    run += v;
  }
  if (tid == 1023 && total) *total = part[1023];
}

__global__ void __launch_bounds__(CP_THREADS) compact_write_kernel(const unsigned char* __restrict__ mask,
                                                                    long long n,
                                                                    const unsigned long long* __restrict__ tile_off,
                                                                    long long* __restrict__ out) {
  __shared__ unsigned int wcnt[CP_ITEMS][CP_THREADS / 32];
  __shared__ unsigned int wbase[CP_ITEMS][CP_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = (long long)blockIdx.x * CP_TILE;
  unsigned bal[CP_ITEMS];
  bool keep[CP_ITEMS];
#pragma unroll
  for (int j = 0; j < CP_ITEMS; ++j) {
    long long i = t0 + (long long)j * CP_THREADS + threadIdx.x;
    keep[j] = i < n && mask[i];
    bal[j] = __ballot_sync(0xffffffffu, keep[j]);
    if (lane == 0) wcnt[j][warp] = __popc(bal[j]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned run = 0;
    for (int j = 0; j < CP_ITEMS; ++j)
      for (int w = 0; w < CP_THREADS / 32; ++w) {
        wbase[j][w] = run;
        run += wcnt[j][w];
      }
  }
  __syncthreads();
  unsigned long long base = tile_off[blockIdx.x];
#pragma unroll
  for (int j = 0; j < CP_ITEMS; ++j) {
    if (keep[j]) {
      long long i = t0 + (long long)j * CP_THREADS + threadIdx.x;
      out[base + wbase[j][warp] + __popc(bal[j] & ((1u << lane) - 1u))] = i;
    }
  }
}

BS_EXPORT long long bs_compact_scratch_bytes(long long n) {
  return ((n + CP_TILE - 1) / CP_TILE + 2) * 8;
}

// count_out: device u64 receiving the number of selected rows.
BS_EXPORT int bs_compact_indices(const unsigned char* mask, long long n, long long* out,
                                 void* scratch, unsigned long long* count_out, cudaStream_t stream) {
  if (n <= 0) return (int)cudaMemsetAsync(count_out, 0, 8, stream);
  long long ntiles = (n + CP_TILE - 1) / CP_TILE;
  unsigned long long* tc = (unsigned long long*)scratch;
  compact_count_kernel<<<(int)ntiles, CP_THREADS, 0, stream>>>(mask, n, tc);
  BS_LAUNCH_CHECK();
  scan_u64_kernel<<<1, 1024, 0, stream>>>(tc, ntiles, count_out);
  BS_LAUNCH_CHECK();
  compact_write_kernel<<<(int)ntiles, CP_THREADS, 0, stream>>>(mask, n, tc, out);
  BS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// segmented reduce: out[seg[i]] (op)= vals[i]; seg ids are non-decreasing
// (sorted input), so a warp first folds runs of equal ids with shuffles and
// only run leaders touch memory.
template <typename V, int OP>
__device__ __forceinline__ V sr_op(V a, V b) {
  if (OP == BS_ADD) return a + b;
  if (OP == BS_MIN) return a < b ? a : b;
  return a > b ? a : b;
}
template <typename V, int OP>
__device__ __forceinline__ void sr_atomic(V* p, V v);
#define SR_ATOMIC(V, CASTT)                                                                     \
  template <>                                                                                   \
  __device__ __forceinline__ void sr_atomic<V, BS_ADD>(V * p, V v) { atomicAdd((CASTT*)p, (CASTT)v); }
SR_ATOMIC(int, int)
SR_ATOMIC(float, float)
SR_ATOMIC(double, double)
template <>
__device__ __forceinline__ void sr_atomic<long long, BS_ADD>(long long* p, long long v) {
  atomicAdd((unsigned long long*)p, (unsigned long long)v);
}
template <>
__device__ __forceinline__ void sr_atomic<int, BS_MIN>(int* p, int v) { atomicMin(p, v); }
template <>
__device__ __forceinline__ void sr_atomic<int, BS_MAX>(int* p, int v) { atomicMax(p, v); }
template <>
__device__ __forceinline__ void sr_atomic<long long, BS_MIN>(long long* p, long long v) { atomicMin(p, v); }
template <>
__device__ __forceinline__ void sr_atomic<long long, BS_MAX>(long long* p, long long v) { atomicMax(p, v); }
template <typename V, int OP>
__device__ __forceinline__ void sr_cas(V* p, V v) {
  typedef typename std::conditional<sizeof(V) == 8, unsigned long long, unsigned int>::type U;
  U* ip = (U*)p;
  U old = *ip;
  while (true) {
    V f;
    memcpy(&f, &old, sizeof(V));
    V nv = sr_op<V, OP>(f, v);
    if (nv == f) break;
    U nb;
    memcpy(&nb, &nv, sizeof(V));
    U prev = atomicCAS(ip, old, nb);
    if (prev == old) break;
    old = prev;
  }
}
template <>
__device__ __forceinline__ void sr_atomic<float, BS_MIN>(float* p, float v) { sr_cas<float, BS_MIN>(p, v); }
template <>
__device__ __forceinline__ void sr_atomic<float, BS_MAX>(float* p, float v) { sr_cas<float, BS_MAX>(p, v); }
template <>
__device__ __forceinline__ void sr_atomic<double, BS_MIN>(double* p, double v) { sr_cas<double, BS_MIN>(p, v); }
template <>
__device__ __forceinline__ void sr_atomic<double, BS_MAX>(double* p, double v) { sr_cas<double, BS_MAX>(p, v); }

template <typename V, int OP>
__global__ void __launch_bounds__(256) segreduce_kernel(const V* __restrict__ vals,
                                                        const long long* __restrict__ seg, long long n,
                                                        V* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  long long limit = ((n + 31) / 32) * 32;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < limit; i += stride) {
    bool ok = i < n;
    long long s = ok ? seg[i] : -1 - lane;  // unique ids for padding lanes
    V v = ok ? vals[i] : (V)0;
    // segmented inclusive scan (by equal seg id) toward lower lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      V ov = __shfl_down_sync(0xffffffffu, v, o);
      long long os = __shfl_down_sync(0xffffffffu, s, o);
      if (lane + o < 32 && os == s) v = sr_op<V, OP>(v, ov);
    }
    long long prev = __shfl_up_sync(0xffffffffu, s, 1);
    bool leader = lane == 0 || prev != s;
    if (ok && leader) sr_atomic<V, OP>(&out[s], v);
  }
}

template <typename V, int OP>
__global__ void __launch_bounds__(256) fill_identity_kernel(V* out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  V id;
  if (OP == BS_ADD) id = (V)0;
  else if (OP == BS_MIN) id = std::numeric_limits<V>::has_infinity ? std::numeric_limits<V>::infinity() : std::numeric_limits<V>::max();
  else id = std::numeric_limits<V>::has_infinity ? -std::numeric_limits<V>::infinity() : std::numeric_limits<V>::lowest();
  for (; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = id;
}

template <typename V, int OP>
static int segreduce_launch(const void* vals, const long long* seg, long long n, void* out,
                            long long nseg, cudaStream_t stream) {
  int sms = bs_num_sms();
  int fb = (int)((nseg + 255) / 256 < (long long)sms * 8 ? (nseg + 255) / 256 : (long long)sms * 8);
  if (fb < 1) fb = 1;
  fill_identity_kernel<V, OP><<<fb, 256, 0, stream>>>((V*)out, nseg);
  BS_LAUNCH_CHECK();
  if (n <= 0) return 0;
  long long want = (n + 255) / 256;
  int blocks = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
  segreduce_kernel<V, OP><<<blocks, 256, 0, stream>>>((const V*)vals, seg, n, (V*)out);
  BS_LAUNCH_CHECK();
  return 0;
}

BS_EXPORT int bs_segreduce(const void* vals, const long long* seg, long long n, void* out,
                           long long nseg, int vtype, int op, cudaStream_t stream) {
  switch (vtype * 3 + op) {
    case BS_I32 * 3 + BS_ADD: return segreduce_launch<int, BS_ADD>(vals, seg, n, out, nseg, stream);
    case BS_I32 * 3 + BS_MIN: return segreduce_launch<int, BS_MIN>(vals, seg, n, out, nseg, stream);
    case BS_I32 * 3 + BS_MAX: return segreduce_launch<int, BS_MAX>(vals, seg, n, out, nseg, stream);
    case BS_I64 * 3 + BS_ADD: return segreduce_launch<long long, BS_ADD>(vals, seg, n, out, nseg, stream);
    case BS_I64 * 3 + BS_MIN: return segreduce_launch<long long, BS_MIN>(vals, seg, n, out, nseg, stream);
    case BS_I64 * 3 + BS_MAX: return segreduce_launch<long long, BS_MAX>(vals, seg, n, out, nseg, stream);
    case BS_F32 * 3 + BS_ADD: return segreduce_launch<float, BS_ADD>(vals, seg, n, out, nseg, stream);
    case BS_F32 * 3 + BS_MIN: return segreduce_launch<float, BS_MIN>(vals, seg, n, out, nseg, stream);
    case BS_F32 * 3 + BS_MAX: return segreduce_launch<float, BS_MAX>(vals, seg, n, out, nseg, stream);
    case BS_F64 * 3 + BS_ADD: return segreduce_launch<double, BS_ADD>(vals, seg, n, out, nseg, stream);
    case BS_F64 * 3 + BS_MIN: return segreduce_launch<double, BS_MIN>(vals, seg, n, out, nseg, stream);
    case BS_F64 * 3 + BS_MAX: return segreduce_launch<double, BS_MAX>(vals, seg, n, out, nseg, stream);
  }
  return -1;
}

// ---------------------------------------------------------------------------
// expand_indices: given exclusive offsets[n+1] of per-row output counts, write
// src[j] = i for every output slot j in [offsets[i], offsets[i+1])  (Flatmap /
// list-column gathers; the device form of repeat_interleave).
__global__ void __launch_bounds__(256) expand_indices_kernel(const long long* __restrict__ offsets,
                                                             long long n, long long total,
                                                             long long* __restrict__ src) {
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; j < total; j += stride) {
    // largest i with offsets[i] <= j
    long long lo = 0, hi = n;
    while (hi - lo > 1) {
      long long mid = (lo + hi) >> 1;
      if (offsets[mid] <= j) lo = mid; else hi = mid;
    }
    src[j] = lo;
  }
}

BS_EXPORT int bs_expand_indices(const long long* offsets, long long n, long long total,
                                long long* src, cudaStream_t stream) {
  if (total <= 0) return 0;
  long long want = (total + 255) / 256;
  int blocks = (int)(want < (long long)bs_num_sms() * 8 ? want : (long long)bs_num_sms() * 8);
  expand_indices_kernel<<<blocks, 256, 0, stream>>>(offsets, n, total, src);
  BS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// CRC32 (IEEE, reflected 0xEDB88320) of a device buffer: every thread CRCs a
// contiguous 4 KiB block bytewise; the per-block CRCs and lengths are combined
// on the host with the GF(2) matrix trick (crc32_combine).  Used to checksum
// device-resident frames before they are spilled / cached.
__constant__ unsigned int crc_table_dev[256];
static unsigned int crc_table_host[256];
static bool crc_ready = false;

static void crc_init() {
  if (crc_ready) return;
  for (unsigned i = 0; i < 256; ++i) {
    unsigned c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
    crc_table_host[i] = c;
  }
  cudaMemcpyToSymbol(crc_table_dev, crc_table_host, sizeof(crc_table_host));
  crc_ready = true;
}

#define CRC_BLOCK 4096
__global__ void __launch_bounds__(128) crc32_blocks_kernel(const unsigned char* __restrict__ data,
                                                           long long n, unsigned int* __restrict__ out) {
  long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nblocks = (n + CRC_BLOCK - 1) / CRC_BLOCK;
  if (b >= nblocks) return;
  long long beg = b * CRC_BLOCK, end = beg + CRC_BLOCK < n ? beg + CRC_BLOCK : n;
  unsigned c = 0xffffffffu;
  for (long long i = beg; i < end; ++i) c = crc_table_dev[(c ^ data[i]) & 0xff] ^ (c >> 8);
  out[b] = c ^ 0xffffffffu;
}

BS_EXPORT long long bs_crc32_nblocks(long long n) { return (n + CRC_BLOCK - 1) / CRC_BLOCK; }
BS_EXPORT int bs_crc32_block_bytes() { return CRC_BLOCK; }
BS_EXPORT int bs_crc32_blocks(const void* data, long long n, unsigned int* out, cudaStream_t stream) {
  crc_init();
  long long nblocks = (n + CRC_BLOCK - 1) / CRC_BLOCK;
  if (nblocks <= 0) return 0;
  crc32_blocks_kernel<<<(int)((nblocks + 127) / 128), 128, 0, stream>>>((const unsigned char*)data, n, out);
  BS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// K15: top-k by radix select (cmd/urls top-N, exec/topn.go).  Eight 8-bit passes
// from the most significant digit narrow down the k-th largest value: each pass
// histograms the digit of the values that still match the prefix found so far
// (shared-memory histogram, one global flush per CTA); the host walks the 256
// counts (one 1 KB D2H per pass).  A final pass writes the indices of all values
// above the threshold and as many ties as are needed.  Values are int64 in an
// order-preserving encoding chosen by the caller.
#define TK_THREADS 512
__device__ __forceinline__ unsigned long long tk_key(long long v) {
  return (unsigned long long)v ^ 0x8000000000000000ULL;
}

__global__ void __launch_bounds__(TK_THREADS) topk_hist_kernel(const long long* __restrict__ vals,
                                                               long long n, unsigned long long prefix,
                                                               unsigned long long mask, int shift,
                                                               unsigned int* __restrict__ hist) {
  __shared__ unsigned int h[256];
  if (threadIdx.x < 256) h[threadIdx.x] = 0;
  __syncthreads();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const unsigned long long k = tk_key(vals[i]);
    if ((k & mask) == prefix) atomicAdd(&h[(unsigned)(k >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 256 && h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

// out_idx[0 .. n_gt) = indices of values above the threshold (any order),
// out_idx[n_gt .. n_gt + n_eq) = indices of the first ties that win the race.
__global__ void __launch_bounds__(TK_THREADS) topk_select_kernel(const long long* __restrict__ vals,
                                                                 long long n, unsigned long long thresh,
                                                                 long long n_gt, long long n_eq,
                                                                 long long* __restrict__ out_idx,
                                                                 unsigned long long* __restrict__ cursors) {
  const unsigned lane = threadIdx.x & 31u;
  const long long limit = ((n + 31) / 32) * 32;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < limit; i += stride) {
    const bool ok = i < n;
    const unsigned long long k = ok ? tk_key(vals[i]) : 0ULL;
    const bool gt = ok && k > thresh, eq = ok && k == thresh;
    const unsigned bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
    unsigned long long baseg = 0, basee = 0;
    if (lane == 0) {
      if (bg) baseg = atomicAdd(&cursors[0], (unsigned long long)__popc(bg));
      if (be) basee = atomicAdd(&cursors[1], (unsigned long long)__popc(be));
    }
    baseg = __shfl_sync(0xffffffffu, baseg, 0);
    basee = __shfl_sync(0xffffffffu, basee, 0);
    const unsigned lt = (1u << lane) - 1u;
    if (gt) {
      const unsigned long long p = baseg + __popc(bg & lt);
      if (p < (unsigned long long)n_gt) out_idx[p] = i;
    }
    if (eq) {
      const unsigned long long p = basee + __popc(be & lt);
      if (p < (unsigned long long)n_eq) out_idx[n_gt + p] = i;
    }
  }
}

// scratch: 256 uint32 histogram + 2 uint64 cursors (caller-provided, >= 1040 bytes).
// Writes min(k, n) indices to out_idx; returns <0 on error.
BS_EXPORT int bs_topk_i64(const long long* vals, long long n, long long k, long long* out_idx,
                          void* scratch, cudaStream_t stream) {
  if (n <= 0 || k <= 0) return 0;
  if (k > n) k = n;
  unsigned int* hist = (unsigned int*)scratch;
  unsigned long long* cursors = (unsigned long long*)((unsigned char*)scratch + 1024);
  const int sms = bs_num_sms();
  long long want = (n + TK_THREADS - 1) / TK_THREADS;
  const int blocks = (int)(want < (long long)sms * 4 ? want : (long long)sms * 4);
  unsigned long long prefix = 0, mask = 0;
  long long remaining = k;  // how many of the values matching the prefix are still wanted
  unsigned int h[256];
  for (int shift = 56; shift >= 0; shift -= 8) {
    cudaError_t e = cudaMemsetAsync(hist, 0, 1024, stream);
    if (e != cudaSuccess) return -(int)e;
    topk_hist_kernel<<<blocks, TK_THREADS, 0, stream>>>(vals, n, prefix, mask, shift, hist);
    BS_COUNT_LAUNCH();
    e = cudaMemcpyAsync(h, hist, sizeof(h), cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return -(int)e;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return -(int)e;
    int d = 255;
    for (; d > 0; --d) {
      if ((long long)h[d] >= remaining) break;
      remaining -= h[d];
    }
    prefix |= (unsigned long long)d << shift;
    mask |= 0xffULL << shift;
  }
  // prefix is now the k-th largest key; `remaining` ties are needed
  const long long n_gt = k - remaining;
  cudaError_t e = cudaMemsetAsync(cursors, 0, 16, stream);
  if (e != cudaSuccess) return -(int)e;
  topk_select_kernel<<<blocks, TK_THREADS, 0, stream>>>(vals, n, prefix, n_gt, remaining, out_idx, cursors);
  BS_COUNT_LAUNCH();
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
