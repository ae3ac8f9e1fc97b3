// K6: merge-path merge of two key-sorted runs (int64 keys), the device half of
// the tiered external sort: runs that do not fit in the HBM budget are parked in
// pinned host memory (then NVMe) by bigslice_b200/sliceio/tiers.py and merged
// back pairwise with this kernel.
//
//   merge_partition  one thread per output tile boundary: binary search along the
//                    merge-path diagonal for the (a, b) split (stable: on ties the
//                    row of A comes first).
//   merge_tiles      one CTA per 2048-element output tile: both input slices are
//                    staged in shared memory, every thread finds its own split by a
//                    second diagonal search and merges 8 elements sequentially.
//                    Emits the merged keys and, for the payload columns, the source
//                    position of every output row (index into A, or |A| + index into
//                    B) — the caller gathers wide columns once with that permutation.
//
// Replaces the reference's heap merge, one row per heap operation
// (sortio/sort.go:161-222).
#include "common.cuh"

#define MG_THREADS 256
#define MG_VT 8
#define MG_TILE (MG_THREADS * MG_VT)

// Largest a in [lo, hi] such that the first `a` elements of A and `d - a` of B are
// the d smallest of the union (A wins ties).
__device__ __forceinline__ long long mg_split(const long long* __restrict__ A, long long na,
                                              const long long* __restrict__ B, long long nb, long long d) {
  long long lo = d > nb ? d - nb : 0, hi = d < na ? d : na;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (A[mid] <= B[d - 1 - mid]) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) merge_partition_kernel(const long long* __restrict__ A, long long na,
                                                              const long long* __restrict__ B, long long nb,
                                                              long long ntiles, long long* __restrict__ part) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > ntiles) return;
  long long d = i * MG_TILE;
  if (d > na + nb) d = na + nb;
  part[i] = mg_split(A, na, B, nb, d);
}

__global__ void __launch_bounds__(MG_THREADS) merge_tiles_kernel(const long long* __restrict__ A, long long na,
                                                                 const long long* __restrict__ B, long long nb,
                                                                 const long long* __restrict__ part,
                                                                 long long* __restrict__ out_keys,
                                                                 long long* __restrict__ out_src) {
  __shared__ long long sa[MG_TILE + 1];
  __shared__ long long sb[MG_TILE + 1];
  const long long tile = blockIdx.x;
  const long long d0 = tile * MG_TILE;
  const long long total = na + nb;
  const long long d1 = d0 + MG_TILE < total ? d0 + MG_TILE : total;
  const long long a0 = part[tile], a1 = part[tile + 1];
  const long long b0 = d0 - a0, b1 = d1 - a1;
  const int ca = (int)(a1 - a0), cb = (int)(b1 - b0);
  for (int i = threadIdx.x; i < ca; i += MG_THREADS) sa[i] = A[a0 + i];
  for (int i = threadIdx.x; i < cb; i += MG_THREADS) sb[i] = B[b0 + i];
  __syncthreads();
  const int n = ca + cb;
  int d = threadIdx.x * MG_VT;
  if (d > n) d = n;
  // split of this thread's first output inside the tile
  int lo = d > cb ? d - cb : 0, hi = d < ca ? d : ca;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sa[mid] <= sb[d - 1 - mid]) lo = mid + 1;
    else hi = mid;
  }
  int ia = lo, ib = d - lo;
#pragma unroll
  for (int k = 0; k < MG_VT; ++k) {
    const int o = d + k;
    if (o >= n) break;
    const bool take_a = ib >= cb || (ia < ca && sa[ia] <= sb[ib]);
    if (take_a) {
      out_keys[d0 + o] = sa[ia];
      out_src[d0 + o] = a0 + ia;
      ++ia;
    } else {
      out_keys[d0 + o] = sb[ib];
      out_src[d0 + o] = na + b0 + ib;
      ++ib;
    }
  }
}

BS_EXPORT long long bs_merge_scratch_elems(long long na, long long nb) {
  return (na + nb + MG_TILE - 1) / MG_TILE + 2;
}

// out_keys/out_src: na + nb elements; scratch: bs_merge_scratch_elems(na, nb) int64.
BS_EXPORT int bs_merge_sorted_i64(const long long* A, long long na, const long long* B, long long nb,
                                  long long* out_keys, long long* out_src, long long* scratch,
                                  cudaStream_t stream) {
  const long long total = na + nb;
  if (total <= 0) return 0;
  const long long ntiles = (total + MG_TILE - 1) / MG_TILE;
  const int pblocks = (int)((ntiles + 1 + 255) / 256);
  merge_partition_kernel<<<pblocks, 256, 0, stream>>>(A, na, B, nb, ntiles, scratch);
  BS_LAUNCH_CHECK();
  merge_tiles_kernel<<<(unsigned int)ntiles, MG_THREADS, 0, stream>>>(A, na, B, nb, scratch, out_keys, out_src);
  BS_LAUNCH_CHECK();
  return 0;
}
