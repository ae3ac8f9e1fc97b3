// K1: bit-exact seeded murmur3 over key columns, and the single-GPU partition
// path (hash -> histogram -> reserve -> scatter permutation) + row gather.
//
// Replaces the reference's per-row defaultPartitioner + Frame.Hash
// (exec/compile.go:20-24, frame/frame.go:389-401) and the row-at-a-time copy
// into per-partition buffers (exec/local.go:217-230).
#include "common.cuh"

long long bs_launch_counter = 0;

BS_EXPORT long long bs_launches() { return bs_launch_counter; }

// ---------------------------------------------------------------------------
// hash_column: out[i] (int64 holding a uint32) = murmur3(col[i], seed),
// optionally XOR-accumulated into out (multi-column key prefixes).
template <int ES>
__global__ void __launch_bounds__(256) hash_column_kernel(const void* __restrict__ col, long long n,
                                                          uint32_t seed, long long* __restrict__ out,
                                                          int xor_into, int narrow_signed) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint32_t h;
    if (ES == 8) {
      h = bs_murmur3_u64(((const uint64_t*)col)[i], seed);
    } else if (ES == 4) {
      h = bs_murmur3_u32(((const uint32_t*)col)[i], seed);
    } else if (ES == 2) {
      uint32_t v = narrow_signed ? (uint32_t)(int32_t)((const int16_t*)col)[i]
                                 : (uint32_t)((const uint16_t*)col)[i];
      h = bs_murmur3_u32(v, seed);
    } else {
      uint32_t v = narrow_signed ? (uint32_t)(int32_t)((const int8_t*)col)[i]
                                 : (uint32_t)((const uint8_t*)col)[i];
      h = bs_murmur3_u32(v, seed);
    }
    long long r = (long long)h;
    if (xor_into) r ^= out[i];
    out[i] = r;
  }
}

BS_EXPORT int bs_hash_column(const void* col, int elem_bytes, int narrow_signed, long long n,
                             unsigned seed, long long* out, int xor_into, cudaStream_t stream) {
  if (n <= 0) return 0;
  int threads = 256;
  long long want = (n + threads - 1) / threads;
  int blocks = (int)(want < (long long)bs_num_sms() * 16 ? want : (long long)bs_num_sms() * 16);
  switch (elem_bytes) {
    case 8: hash_column_kernel<8><<<blocks, threads, 0, stream>>>(col, n, seed, out, xor_into, narrow_signed); break;
    case 4: hash_column_kernel<4><<<blocks, threads, 0, stream>>>(col, n, seed, out, xor_into, narrow_signed); break;
    case 2: hash_column_kernel<2><<<blocks, threads, 0, stream>>>(col, n, seed, out, xor_into, narrow_signed); break;
    case 1: hash_column_kernel<1><<<blocks, threads, 0, stream>>>(col, n, seed, out, xor_into, narrow_signed); break;
    default: return -1;
  }
  BS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// Partition permutation.  dest(i) = ids ? ids[i] : murmur3(keys[i], 0) % npart.
// Pass 1 counts rows per destination (shared-memory histogram, one global
// atomic per destination per CTA); an exclusive scan turns counts into
// cursors; pass 2 ranks rows inside the tile with shared atomics, reserves one
// contiguous range per destination per tile and writes the permutation.
#define PART_TILE 4096
#define PART_THREADS 512
#define PART_MAX 4096

__device__ __forceinline__ int part_dest(const long long* ids, const long long* keys, long long i,
                                         int npart) {
  if (ids) return (int)ids[i];
  return (int)(bs_murmur3_u64((uint64_t)keys[i], 0u) % (uint32_t)npart);
}

__global__ void __launch_bounds__(PART_THREADS) partition_count_kernel(
    const long long* __restrict__ ids, const long long* __restrict__ keys, long long n, int npart,
    long long* __restrict__ counts) {
  extern __shared__ int hist[];
  for (int d = threadIdx.x; d < npart; d += blockDim.x) hist[d] = 0;
  __syncthreads();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) atomicAdd(&hist[part_dest(ids, keys, i, npart)], 1);
  __syncthreads();
  for (int d = threadIdx.x; d < npart; d += blockDim.x)
    if (hist[d]) atomicAdd((unsigned long long*)&counts[d], (unsigned long long)hist[d]);
}

__global__ void partition_scan_kernel(const long long* __restrict__ counts, int npart,
                                      long long* __restrict__ cursors) {
  // npart is small (<= 4096): one thread does the exclusive scan.
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long acc = 0;
    for (int d = 0; d < npart; ++d) {
      cursors[d] = acc;
      acc += counts[d];
    }
  }
}

__global__ void __launch_bounds__(PART_THREADS) partition_scatter_kernel(
    const long long* __restrict__ ids, const long long* __restrict__ keys, long long n, int npart,
    long long* __restrict__ cursors, long long* __restrict__ perm) {
  extern __shared__ int sm[];
  int* hist = sm;                 // [npart] rows of this tile per destination
  int* base_lo = sm + npart;      // [npart*2] 64-bit reserved base per destination
  long long* base = (long long*)(base_lo + (npart & 1));  // 8-byte aligned
  long long ntiles = (n + PART_TILE - 1) / PART_TILE;
  constexpr int PER = PART_TILE / PART_THREADS;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    for (int d = threadIdx.x; d < npart; d += blockDim.x) hist[d] = 0;
    __syncthreads();
    int dest[PER], rank[PER];
    long long row0 = tile * PART_TILE;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      long long i = row0 + threadIdx.x + (long long)j * PART_THREADS;
      if (i < n) {
        dest[j] = part_dest(ids, keys, i, npart);
        rank[j] = atomicAdd(&hist[dest[j]], 1);
      } else {
        dest[j] = -1;
      }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < npart; d += blockDim.x)
      if (hist[d])
        base[d] = (long long)atomicAdd((unsigned long long*)&cursors[d], (unsigned long long)hist[d]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (dest[j] >= 0) {
        long long i = row0 + threadIdx.x + (long long)j * PART_THREADS;
        perm[base[dest[j]] + rank[j]] = i;
      }
    }
    __syncthreads();
  }
}

// counts[npart] (zeroed by the caller), cursors[npart] scratch, perm[n] out.
BS_EXPORT int bs_partition_perm(const long long* ids, const long long* keys, long long n, int npart,
                                long long* counts, long long* cursors, long long* perm,
                                cudaStream_t stream) {
  if (npart > PART_MAX || npart <= 0) return -1;
  if (n <= 0) return 0;
  int sms = bs_num_sms();
  long long want = (n + PART_THREADS - 1) / PART_THREADS;
  int blocks = (int)(want < (long long)sms * 4 ? want : (long long)sms * 4);
  partition_count_kernel<<<blocks, PART_THREADS, npart * sizeof(int), stream>>>(ids, keys, n, npart, counts);
  BS_LAUNCH_CHECK();
  partition_scan_kernel<<<1, 32, 0, stream>>>(counts, npart, cursors);
  BS_LAUNCH_CHECK();
  long long ntiles = (n + PART_TILE - 1) / PART_TILE;
  int blocks2 = (int)(ntiles < (long long)sms * 4 ? ntiles : (long long)sms * 4);
  size_t smem = (size_t)npart * sizeof(int) + (size_t)(npart + 1) * sizeof(long long) + 8;
  partition_scatter_kernel<<<blocks2, PART_THREADS, smem, stream>>>(ids, keys, n, npart, cursors, perm);
  BS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// gather_rows: dst[i] = src[idx[i]] for rows of ES bytes (16-byte vectorised
// when the row size allows).
template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const T* __restrict__ src,
                                                          const long long* __restrict__ idx,
                                                          long long n, int per_row,
                                                          T* __restrict__ dst) {
  long long total = n * per_row;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    long long r = i / per_row;
    int c = (int)(i - r * per_row);
    dst[i] = src[idx[r] * per_row + c];
  }
}

BS_EXPORT int bs_gather_rows(const void* src, int row_bytes, const long long* idx, long long n,
                             void* dst, cudaStream_t stream) {
  if (n <= 0) return 0;
  int threads = 256;
  int sms = bs_num_sms();
#define GATHER(T)                                                                                   \
  {                                                                                                 \
    int per = row_bytes / (int)sizeof(T);                                                           \
    long long want = (n * per + threads - 1) / threads;                                             \
    int blocks = (int)(want < (long long)sms * 16 ? want : (long long)sms * 16);                    \
    gather_rows_kernel<T><<<blocks, threads, 0, stream>>>((const T*)src, idx, n, per, (T*)dst);     \
  }
  bool a16 = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
  if (row_bytes % 16 == 0 && a16) GATHER(int4)
  else if (row_bytes % 8 == 0) GATHER(long long)
  else if (row_bytes % 4 == 0) GATHER(int)
  else if (row_bytes % 2 == 0) GATHER(short)
  else GATHER(char)
#undef GATHER
  BS_LAUNCH_CHECK();
  return 0;
}
