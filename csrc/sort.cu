// K5: stable LSD radix sort of <int64 key, 64-bit payload> pairs.
//
// 8-bit digits; one up-front kernel builds all eight digit histograms so that
// passes whose digit is constant across the input are skipped (keys below 2^32
// cost four passes).  Each pass is chunked: a fixed grid of CTAs owns
// contiguous chunks, so the per-CTA digit counts form a small matrix that one
// CTA scans, and the scatter keeps running per-digit offsets in shared memory.
// Ranking inside a tile is warp-synchronous (__match_any_sync) and therefore
// stable.  Payload is typically the row index; columns are then moved with
// bs_gather_rows.
//
// Replaces sort.Sort(frame) / the external sorter's in-memory runs in the
// reference (sortio/sort.go:52, exec/combiner.go:262).
#include "common.cuh"

#define RS_THREADS 512
#define RS_WARPS (RS_THREADS / 32)
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)

__device__ __forceinline__ unsigned rs_digit(long long key, int shift) {
  unsigned long long u = (unsigned long long)key ^ 0x8000000000000000ULL;  // signed order
  return (unsigned)(u >> shift) & 0xffu;
}

__global__ void __launch_bounds__(RS_THREADS) rs_global_hist_kernel(const long long* __restrict__ keys,
                                                                     long long n,
                                                                     unsigned long long* __restrict__ ghist) {
  __shared__ unsigned int h[8][256];
  for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&h[0][0])[i] = 0;
  __syncthreads();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    long long k = keys[i];
#pragma unroll
    for (int p = 0; p < 8; ++p) atomicAdd(&h[p][rs_digit(k, p * 8)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) {
    unsigned v = (&h[0][0])[i];
    if (v) atomicAdd(&ghist[i], (unsigned long long)v);
  }
}

// counts[digit * nchunks + chunk]
__global__ void __launch_bounds__(RS_THREADS) rs_chunk_hist_kernel(const long long* __restrict__ keys,
                                                                    long long n, long long chunk_rows,
                                                                    int shift,
                                                                    unsigned long long* __restrict__ counts) {
  __shared__ unsigned int h[256];
  if (threadIdx.x < 256) h[threadIdx.x] = 0;
  __syncthreads();
  long long beg = (long long)blockIdx.x * chunk_rows;
  long long end = beg + chunk_rows < n ? beg + chunk_rows : n;
  for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) atomicAdd(&h[rs_digit(keys[i], shift)], 1u);
  __syncthreads();
  if (threadIdx.x < 256) counts[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of m values in place by a single CTA
__global__ void __launch_bounds__(1024) rs_scan_kernel(unsigned long long* __restrict__ a, long long m) {
  __shared__ unsigned long long part[1024];
  const int tid = threadIdx.x;
  long long per = (m + 1023) / 1024;
  long long beg = (long long)tid * per, end = beg + per < m ? beg + per : m;
  unsigned long long s = 0;
  for (long long i = beg; i < end; ++i) s += a[i];
  part[tid] = s;
  __syncthreads();
  // Hillis-Steele over 1024 partials
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
}

// Scatter pass.  Items are first placed in shared memory in tile-sorted order
// (digit-major, stable), then written out so that consecutive threads store to
// consecutive addresses inside each digit's run: coalesced 8-byte stores
// instead of 256-way scattered ones.
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(
    const long long* __restrict__ keys_in, const long long* __restrict__ vals_in,
    long long* __restrict__ keys_out, long long* __restrict__ vals_out, long long n,
    long long chunk_rows, int shift, const unsigned long long* __restrict__ offsets) {
  __shared__ unsigned long long run[256];           // next output slot per digit for this chunk
  __shared__ unsigned int tstart[257];              // tile-local start of each digit's run
  __shared__ unsigned int wcount[RS_WARPS][256];    // per-warp digit counts -> warp bases
  extern __shared__ __align__(16) unsigned char rs_smem[];
  long long* skey = (long long*)rs_smem;            // [RS_TILE]
  long long* sval = skey + RS_TILE;                 // [RS_TILE]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 256) run[tid] = offsets[(size_t)tid * gridDim.x + blockIdx.x];
  long long beg = (long long)blockIdx.x * chunk_rows;
  long long end = beg + chunk_rows < n ? beg + chunk_rows : n;
  const unsigned lt = (1u << lane) - 1u;
  for (long long t0 = beg; t0 < end; t0 += RS_TILE) {
    for (int i = tid; i < RS_WARPS * 256; i += RS_THREADS) (&wcount[0][0])[i] = 0;
    __syncthreads();
    long long k[RS_ITEMS], v[RS_ITEMS];
    unsigned dg[RS_ITEMS], rk[RS_ITEMS];
    const long long wbase = t0 + (long long)warp * 32 * RS_ITEMS;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
      long long i = wbase + j * 32 + lane;
      bool ok = i < end;
      k[j] = 0;
      v[j] = 0;
      if (ok) {
        k[j] = keys_in[i];
        v[j] = vals_in[i];
      }
      dg[j] = ok ? rs_digit(k[j], shift) : 0xffffffffu;
      unsigned m = __match_any_sync(0xffffffffu, dg[j]);
      unsigned c = ok ? wcount[warp][dg[j]] : 0;
      __syncwarp();
      rk[j] = c + __popc(m & lt);
      if (ok && (m & lt) == 0) wcount[warp][dg[j]] = c + __popc(m);
      __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive scan over warps -> warp bases; digit totals in tstart[d+1]
    if (tid < 256) {
      unsigned acc = 0;
      for (int w = 0; w < RS_WARPS; ++w) {
        unsigned c = wcount[w][tid];
        wcount[w][tid] = acc;
        acc += c;
      }
      tstart[tid + 1] = acc;
      if (tid == 0) tstart[0] = 0;
    }
    __syncthreads();
    // exclusive scan of the 256 digit totals (one warp, 8 per lane)
    if (tid < 32) {
      unsigned loc[8], sum = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        loc[q] = tstart[tid * 8 + q + 1];
        sum += loc[q];
      }
      unsigned incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      unsigned runv = incl - sum;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        runv += loc[q];
        tstart[tid * 8 + q + 1] = runv;  // inclusive end of digit (tid*8+q) == start of the next
      }
    }
    __syncthreads();
    // place items at their tile-sorted position
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
      if (dg[j] != 0xffffffffu) {
        unsigned lpos = tstart[dg[j]] + wcount[warp][dg[j]] + rk[j];
        skey[lpos] = k[j];
        sval[lpos] = v[j];
      }
    }
    __syncthreads();
    const unsigned ntile = tstart[256];
    for (unsigned p = tid; p < ntile; p += RS_THREADS) {
      long long kk = skey[p];
      unsigned d = rs_digit(kk, shift);
      unsigned long long pos = run[d] + (p - tstart[d]);
      keys_out[pos] = kk;
      vals_out[pos] = sval[p];
    }
    __syncthreads();
    if (tid < 256) run[tid] += tstart[tid + 1] - tstart[tid];
    // (the zeroing loop + barrier at the top of the next iteration orders this)
  }
}

// scratch layout (unsigned long long units): [8*256 global hist][256*nchunks counts]
BS_EXPORT long long bs_sort_scratch_bytes(long long n) {
  int nchunks = bs_num_sms() * 4;
  (void)n;
  return (long long)(8 * 256 + 256 * (long long)nchunks) * 8;
}

// Sorts (keys_a, vals_a) using (keys_b, vals_b) as ping-pong buffers.  Returns
// <0 on error, else 0 if the result is in the *_a buffers and 1 if in *_b.
BS_EXPORT int bs_radix_sort_pairs(long long* keys_a, long long* vals_a, long long* keys_b,
                                  long long* vals_b, long long n, void* scratch,
                                  cudaStream_t stream) {
  if (n <= 1) return 0;
  int sms = bs_num_sms();
  int nchunks = sms * 4;
  long long chunk_rows = (n + nchunks - 1) / nchunks;
  chunk_rows = ((chunk_rows + RS_TILE - 1) / RS_TILE) * RS_TILE;
  nchunks = (int)((n + chunk_rows - 1) / chunk_rows);
  unsigned long long* ghist = (unsigned long long*)scratch;
  unsigned long long* counts = ghist + 8 * 256;
  cudaError_t e = cudaMemsetAsync(ghist, 0, 8 * 256 * sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return -(int)e;
  {
    long long want = (n + RS_THREADS - 1) / RS_THREADS;
    int blocks = (int)(want < (long long)sms * 4 ? want : (long long)sms * 4);
    rs_global_hist_kernel<<<blocks, RS_THREADS, 0, stream>>>(keys_a, n, ghist);
    BS_COUNT_LAUNCH();
  }
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(rs_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_TILE * 16);
    if (e != cudaSuccess) return -(int)e;
    attr_set = true;
  }
  unsigned long long h[8 * 256];
  e = cudaMemcpyAsync(h, ghist, sizeof(h), cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess) return -(int)e;
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return -(int)e;
  int flip = 0;
  for (int p = 0; p < 8; ++p) {
    bool constant = false;
    for (int d = 0; d < 256; ++d)
      if (h[p * 256 + d] == (unsigned long long)n) constant = true;
    if (constant) continue;  // every key has the same digit: pass is the identity
    const long long* kin = flip ? keys_b : keys_a;
    const long long* vin = flip ? vals_b : vals_a;
    long long* kout = flip ? keys_a : keys_b;
    long long* vout = flip ? vals_a : vals_b;
    rs_chunk_hist_kernel<<<nchunks, RS_THREADS, 0, stream>>>(kin, n, chunk_rows, p * 8, counts);
    BS_COUNT_LAUNCH();
    rs_scan_kernel<<<1, 1024, 0, stream>>>(counts, 256LL * nchunks);
    BS_COUNT_LAUNCH();
    rs_scatter_kernel<<<nchunks, RS_THREADS, RS_TILE * 16, stream>>>(kin, vin, kout, vout, n, chunk_rows, p * 8, counts);
    BS_COUNT_LAUNCH();
    flip ^= 1;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return -(int)e;
  return flip;
}
