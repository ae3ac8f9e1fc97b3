// K7: Cogroup on the device — run detection, sort-merge union of two key-sorted
// inputs emitting CSR group sizes, and the replicated-small-side (broadcast)
// variant that multicasts rows to every GPU through NVSwitch (multimem.st).
//
// Cogroup of inputs A and B (each already radix-sorted by its int64 key):
//
//   cg_run_flags   flags[i] = (i == 0 || key[i] != key[i-1])        (per input)
//   (stream compaction of the flags gives the distinct keys + run starts)
//   cg_search      for every distinct key of one input: its lower-bound rank in
//                  the other input's distinct keys and whether it occurs there
//                  (binary search over an L2-resident array: both lists are a
//                  few tens of MB)
//   (exclusive scan of the match flags = number of common keys below each key)
//   cg_place       union position of key i of A  =  i + rankB(i) - commonBelow(i)
//                  (symmetrically for B; a common key gets the same position from
//                  both sides): writes the union key and the group sizes of both
//                  inputs — the CSR offsets are an exclusive scan away and the
//                  value columns are the sorted input columns, untouched.
//
// No per-row heap, no host loop: replaces the reference's k-way heap merge
// (cogroup.go:191-260).
#include "common.cuh"

__global__ void __launch_bounds__(256) cg_run_flags_kernel(const long long* __restrict__ key, long long n,
                                                           unsigned char* __restrict__ flags) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) flags[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

__device__ __forceinline__ long long cg_lower_bound(const long long* __restrict__ a, long long n, long long x) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (__ldg(&a[mid]) < x) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) cg_search_kernel(const long long* __restrict__ mine, long long n,
                                                        const long long* __restrict__ other, long long m,
                                                        long long* __restrict__ rank,
                                                        long long* __restrict__ match) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const long long x = mine[i];
    const long long r = cg_lower_bound(other, m, x);
    rank[i] = r;
    match[i] = (r < m && __ldg(&other[r]) == x) ? 1 : 0;
  }
}

// side 0 writes the keys and its sizes; side 1 writes its sizes (and the keys it
// alone owns).  `common_below` is the exclusive scan of `match`.
__global__ void __launch_bounds__(256) cg_place_kernel(const long long* __restrict__ mine, long long n,
                                                       const long long* __restrict__ rank,
                                                       const long long* __restrict__ match,
                                                       const long long* __restrict__ common_below,
                                                       const long long* __restrict__ run_start,
                                                       long long nrows, long long* __restrict__ union_keys,
                                                       long long* __restrict__ sizes, int write_common_key) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const long long pos = i + rank[i] - common_below[i];
    const long long end = i + 1 < n ? run_start[i + 1] : nrows;
    sizes[pos] = end - run_start[i];
    if (write_common_key || !match[i]) union_keys[pos] = mine[i];
  }
}

// ---- broadcast of a (small) table to every GPU through the switch -----------------
// `mc` is the multicast (NVLS) mapping of a symmetric buffer: one multimem.st lands
// the 16 bytes in every GPU's copy.  Rank r writes its rows at row offset `at`.
__global__ void __launch_bounds__(256) cg_multicast_rows_kernel(const int4* __restrict__ src, long long n16,
                                                                int4* mc, long long at16) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) {
    const int4 v = src[i];
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + at16 + i),
                 "f"(__int_as_float(v.x)), "f"(__int_as_float(v.y)), "f"(__int_as_float(v.z)),
                 "f"(__int_as_float(v.w))
                 : "memory");
  }
  __threadfence_system();
}

// Fallback without multicast hardware/driver support: the same all-gather with
// plain peer stores, one pass per destination.
__global__ void __launch_bounds__(256) cg_peer_rows_kernel(const int4* __restrict__ src, long long n16,
                                                           int4* const* peers, int world, long long at16) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) {
    const int4 v = src[i];
    for (int p = 0; p < world; ++p) peers[p][at16 + i] = v;
  }
  __threadfence_system();
}

// ---------------------------------------------------------------------------
static int cg_blocks(long long n, int per = 256) {
  long long want = (n + per - 1) / per;
  long long cap = (long long)bs_num_sms() * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

BS_EXPORT int bs_cg_run_flags(const long long* key, long long n, unsigned char* flags, cudaStream_t stream) {
  if (n <= 0) return 0;
  cg_run_flags_kernel<<<cg_blocks(n), 256, 0, stream>>>(key, n, flags);
  BS_LAUNCH_CHECK();
  return 0;
}

BS_EXPORT int bs_cg_search(const long long* mine, long long n, const long long* other, long long m,
                           long long* rank, long long* match, cudaStream_t stream) {
  if (n <= 0) return 0;
  cg_search_kernel<<<cg_blocks(n), 256, 0, stream>>>(mine, n, other, m, rank, match);
  BS_LAUNCH_CHECK();
  return 0;
}

BS_EXPORT int bs_cg_place(const long long* mine, long long n, const long long* rank, const long long* match,
                          const long long* common_below, const long long* run_start, long long nrows,
                          long long* union_keys, long long* sizes, int write_common_key,
                          cudaStream_t stream) {
  if (n <= 0) return 0;
  cg_place_kernel<<<cg_blocks(n), 256, 0, stream>>>(mine, n, rank, match, common_below, run_start, nrows,
                                                    union_keys, sizes, write_common_key);
  BS_LAUNCH_CHECK();
  return 0;
}

BS_EXPORT int bs_cg_multicast_rows(const void* src, long long nbytes, void* mc_base, long long at_bytes,
                                   cudaStream_t stream) {
  if (nbytes <= 0) return 0;
  if ((nbytes | at_bytes) & 15) return -2;
  cg_multicast_rows_kernel<<<cg_blocks(nbytes / 16), 256, 0, stream>>>((const int4*)src, nbytes / 16,
                                                                       (int4*)mc_base, at_bytes / 16);
  BS_LAUNCH_CHECK();
  return 0;
}

BS_EXPORT int bs_cg_peer_rows(const void* src, long long nbytes, void* const* peers_d, int world,
                              long long at_bytes, cudaStream_t stream) {
  if (nbytes <= 0) return 0;
  if ((nbytes | at_bytes) & 15) return -2;
  cg_peer_rows_kernel<<<cg_blocks(nbytes / 16), 256, 0, stream>>>((const int4*)src, nbytes / 16,
                                                                  (int4* const*)peers_d, world, at_bytes / 16);
  BS_LAUNCH_CHECK();
  return 0;
}
