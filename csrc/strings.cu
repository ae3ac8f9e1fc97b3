// K10: variable-length (string) columns on the device.
//
// A string column is CSR: offsets[n+1] (int64) + data (uint8).  Kernels:
//
//   str_hash        seeded MurmurHash3_x86_32 over each string's bytes — bit-exact
//                   with the host reference (frame/hashing.py: murmur3_bytes; the
//                   reference hashes strings the same way, frame/ops_builtin.go:13-20),
//                   so shard placement of string keys computed on the GPU equals the
//                   CPU's.  `str_key64` packs two differently seeded hashes into one
//                   int64: the dictionary id string keys ride the int64 table path with.
//   str_split_*     tokeniser (count pass -> host-side scan of the counts -> fill pass):
//                   splits every row on a set of delimiter bytes, emitting token
//                   (start, length) pairs into the parent data plus each token's parent
//                   row — the Flatmap of the tutorial wordcount (docs/index.md:164-278).
//   str_domain      per-row URL -> host part ("scheme://host[:port]/..." or "host/..."),
//                   lower-cased in place, as (start, length) — cmd/urls' extraction
//                   (cmd/urls/urls.go:64-84).
//   str_copy_spans  gathers spans (start, length) of a data buffer into a new dense
//                   data buffer (one warp per span, coalesced byte copies) — used to
//                   materialise tokens/domains and to take rows of a string column.
//   str_equal_spans compares span pairs (dictionary verification: equal ids must be
//                   equal strings).
#include "common.cuh"

__device__ __forceinline__ uint32_t str_murmur3(const unsigned char* __restrict__ p, long long len,
                                                uint32_t seed) {
  uint32_t h = seed;
  const long long nblocks = len >> 2;
  for (long long i = 0; i < nblocks; ++i) {
    const unsigned char* q = p + (i << 2);
    const uint32_t k = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
    h = bs_mm3_h(h, k);
  }
  const unsigned char* tail = p + (nblocks << 2);
  uint32_t k1 = 0;
  switch (len & 3) {
    case 3: k1 ^= (uint32_t)tail[2] << 16;
    case 2: k1 ^= (uint32_t)tail[1] << 8;
    case 1: k1 ^= (uint32_t)tail[0];
            h ^= bs_mm3_k(k1);
  }
  return bs_fmix32(h ^ (uint32_t)len);
}

__global__ void __launch_bounds__(256) str_hash_kernel(const long long* __restrict__ off,
                                                       const unsigned char* __restrict__ data, long long n,
                                                       uint32_t seed, long long* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (long long)str_murmur3(data + off[i], off[i + 1] - off[i], seed);
}

__global__ void __launch_bounds__(256) str_key64_kernel(const long long* __restrict__ off,
                                                        const unsigned char* __restrict__ data, long long n,
                                                        long long* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const unsigned char* p = data + off[i];
    const long long len = off[i + 1] - off[i];
    const uint64_t lo = str_murmur3(p, len, 0u), hi = str_murmur3(p, len, 0x9747b28cu);
    long long k = (long long)(lo | (hi << 32));
    if (k == BS_EMPTY_KEY) k = 0x7fffffffffffffffLL;  // keep the table's reserved key free
    out[i] = k;
  }
}

__device__ __forceinline__ bool str_is_delim(unsigned char c, const unsigned char* __restrict__ delim, int nd) {
  for (int j = 0; j < nd; ++j)
    if (c == delim[j]) return true;
  return false;
}

// pass 1: tokens per row
__global__ void __launch_bounds__(256) str_split_count_kernel(const long long* __restrict__ off,
                                                              const unsigned char* __restrict__ data,
                                                              long long n, const unsigned char* __restrict__ delim,
                                                              int nd, long long* __restrict__ counts) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    long long c = 0;
    bool in = false;
    for (long long p = off[i]; p < off[i + 1]; ++p) {
      const bool d = str_is_delim(data[p], delim, nd);
      if (!d && !in) ++c;
      in = !d;
    }
    counts[i] = c;
  }
}

// pass 2: token spans; tok_base[i] = first token index of row i (exclusive scan of counts)
__global__ void __launch_bounds__(256) str_split_fill_kernel(const long long* __restrict__ off,
                                                             const unsigned char* __restrict__ data,
                                                             long long n, const unsigned char* __restrict__ delim,
                                                             int nd, const long long* __restrict__ tok_base,
                                                             long long* __restrict__ start,
                                                             long long* __restrict__ len,
                                                             long long* __restrict__ parent) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    long long t = tok_base[i];
    long long s = -1;
    for (long long p = off[i]; p <= off[i + 1]; ++p) {
      const bool d = p == off[i + 1] || str_is_delim(data[p], delim, nd);
      if (!d && s < 0) s = p;
      if (d && s >= 0) {
        start[t] = s;
        len[t] = p - s;
        parent[t] = i;
        ++t;
        s = -1;
      }
    }
  }
}

// URL -> host part, lower-cased in place.  len 0 when the row holds no host.
__global__ void __launch_bounds__(256) str_domain_kernel(const long long* __restrict__ off, unsigned char* data,
                                                         long long n, long long* __restrict__ start,
                                                         long long* __restrict__ len) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const long long a = off[i], b = off[i + 1];
    long long s = a;
    for (long long p = a; p + 2 < b; ++p)
      if (data[p] == ':' && data[p + 1] == '/' && data[p + 2] == '/') {
        s = p + 3;
        break;
      }
    long long e = s;
    while (e < b && data[e] != '/' && data[e] != '?' && data[e] != '#') ++e;
    // drop userinfo ("user@host")
    for (long long p = s; p < e; ++p)
      if (data[p] == '@') s = p + 1;
    for (long long p = s; p < e; ++p)
      if (data[p] >= 'A' && data[p] <= 'Z') data[p] += 32;
    start[i] = s;
    len[i] = e - s;
  }
}

// one warp per span: dst[dst_off[i] .. +len[i]) = src[start[i] .. +len[i])
__global__ void __launch_bounds__(256) str_copy_spans_kernel(const unsigned char* __restrict__ src,
                                                             const long long* __restrict__ start,
                                                             const long long* __restrict__ dst_off, long long n,
                                                             unsigned char* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (; w < n; w += nw) {
    const long long s = start[w], d = dst_off[w], L = dst_off[w + 1] - d;
    for (long long j = lane; j < L; j += 32) dst[d + j] = src[s + j];
  }
}

__global__ void __launch_bounds__(256) str_equal_spans_kernel(const unsigned char* __restrict__ a,
                                                              const long long* __restrict__ a_start,
                                                              const long long* __restrict__ a_len,
                                                              const unsigned char* __restrict__ b,
                                                              const long long* __restrict__ b_start,
                                                              const long long* __restrict__ b_len, long long n,
                                                              unsigned char* __restrict__ eq) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    bool same = a_len[i] == b_len[i];
    if (same) {
      const unsigned char* p = a + a_start[i];
      const unsigned char* q = b + b_start[i];
      for (long long j = 0; j < a_len[i]; ++j)
        if (p[j] != q[j]) {
          same = false;
          break;
        }
    }
    eq[i] = same ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------
static int str_blocks(long long n) {
  long long want = (n + 255) / 256, cap = (long long)bs_num_sms() * 16;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

BS_EXPORT int bs_str_hash(const long long* off, const void* data, long long n, unsigned int seed,
                          long long* out, cudaStream_t stream) {
  if (n <= 0) return 0;
  str_hash_kernel<<<str_blocks(n), 256, 0, stream>>>(off, (const unsigned char*)data, n, seed, out);
  BS_LAUNCH_CHECK();
  return 0;
}
BS_EXPORT int bs_str_key64(const long long* off, const void* data, long long n, long long* out,
                           cudaStream_t stream) {
  if (n <= 0) return 0;
  str_key64_kernel<<<str_blocks(n), 256, 0, stream>>>(off, (const unsigned char*)data, n, out);
  BS_LAUNCH_CHECK();
  return 0;
}
BS_EXPORT int bs_str_split_count(const long long* off, const void* data, long long n, const void* delim, int nd,
                                 long long* counts, cudaStream_t stream) {
  if (n <= 0) return 0;
  str_split_count_kernel<<<str_blocks(n), 256, 0, stream>>>(off, (const unsigned char*)data, n,
                                                            (const unsigned char*)delim, nd, counts);
  BS_LAUNCH_CHECK();
  return 0;
}
BS_EXPORT int bs_str_split_fill(const long long* off, const void* data, long long n, const void* delim, int nd,
                                const long long* tok_base, long long* start, long long* len, long long* parent,
                                cudaStream_t stream) {
  if (n <= 0) return 0;
  str_split_fill_kernel<<<str_blocks(n), 256, 0, stream>>>(off, (const unsigned char*)data, n,
                                                           (const unsigned char*)delim, nd, tok_base, start, len,
                                                           parent);
  BS_LAUNCH_CHECK();
  return 0;
}
BS_EXPORT int bs_str_domain(const long long* off, void* data, long long n, long long* start, long long* len,
                            cudaStream_t stream) {
  if (n <= 0) return 0;
  str_domain_kernel<<<str_blocks(n), 256, 0, stream>>>(off, (unsigned char*)data, n, start, len);
  BS_LAUNCH_CHECK();
  return 0;
}
BS_EXPORT int bs_str_copy_spans(const void* src, const long long* start, const long long* dst_off, long long n,
                                void* dst, cudaStream_t stream) {
  if (n <= 0) return 0;
  long long want = (n * 32 + 255) / 256, cap = (long long)bs_num_sms() * 16;
  int blocks = (int)(want < cap ? want : cap);
  str_copy_spans_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>((const unsigned char*)src, start, dst_off, n,
                                                                     (unsigned char*)dst);
  BS_LAUNCH_CHECK();
  return 0;
}
BS_EXPORT int bs_str_equal_spans(const void* a, const long long* a_start, const long long* a_len, const void* b,
                                 const long long* b_start, const long long* b_len, long long n, void* eq,
                                 cudaStream_t stream) {
  if (n <= 0) return 0;
  str_equal_spans_kernel<<<str_blocks(n), 256, 0, stream>>>((const unsigned char*)a, a_start, a_len,
                                                            (const unsigned char*)b, b_start, b_len, n,
                                                            (unsigned char*)eq);
  BS_LAUNCH_CHECK();
  return 0;
}
