// K2: fused hash -> partition -> all-to-all over NVLink peer memory.
//
// Every rank maps every peer's control block and receive arena (CUDA IPC).  A
// shuffle is four launches on each rank's stream and **no NCCL call**:
//
//   xchg_count   per-destination-shard row counts of the local rows
//                (murmur3 of the key prefix, or caller-provided ids; optional
//                "skip EMPTY" predicate so a hash table can be shuffled in
//                place — fused compaction).
//   xchg_plan    one CTA: stores this rank's counts row (+ its arena base) into
//                every peer's control block with P2P stores, releases a flag at
//                system scope, acquire-spins for all peers' flags, then turns
//                the W x P count matrix into exact, densely packed destination
//                addresses for every (dest shard, column).  This is the
//                all-gather of counts done with peer stores + flags.
//   xchg_scatter per 2048-row tile: rank rows by destination in shared memory,
//                reserve a slot range per destination with one local atomic,
//                stage each column in destination order in shared memory and
//                write it out with coalesced stores that land directly in the
//                destination GPU's HBM (NVLink).  The last CTA releases the
//                "data complete" flag on every peer.
//   xchg_wait    acquire-spins until every peer's data flag reached the epoch.
//
// Spins are bounded (timeout -> abort flag) so a dead peer cannot hang the GPU.
// Replaces the reference's pull-based file+RPC shuffle (exec/bigmachine.go:
// 857-1005) and the in-memory row copies of exec/local.go:217-230.
#include "common.cuh"
#include <string.h>

#define XC_MAXW 16    // ranks
#define XC_MAXP 256   // destination shards
#define XC_MAXC 8     // columns per row
#define XC_MAXK 4     // key columns hashed
#define XC_TILE 2048
#define XC_THREADS 512
#define XC_TIMEOUT_NS 30000000000ULL

// Symmetric control block, one per rank, mapped by every peer.
struct XcCtrl {
  unsigned long long count_flag[XC_MAXW];   // epoch of the last counts row from rank s
  unsigned long long data_flag[XC_MAXW];    // epoch of the last completed scatter from rank s
  unsigned long long arena_base[2][XC_MAXW];
  long long cnt[2][XC_MAXW][XC_MAXP];
  unsigned long long abort_flag;
  unsigned long long barrier_flag[XC_MAXW];
  // device-side vote published with the counts row: [0] producer status (0 ok),
  // [1] fingerprint of the invocations issued since the last exchange
  unsigned long long vote[2][XC_MAXW][2];
};

struct XcPeers {
  int world, rank;
  XcCtrl* ctrl[XC_MAXW];            // peer-mapped control blocks
  unsigned char* arena[XC_MAXW];    // peer-mapped receive arenas
  unsigned long long arena_bytes;
};

struct XcSpec {
  long long n;                       // local rows (or table slots)
  int nparts;                        // destination shards; shard d lives on rank d % world
  int ncols;
  int nkeys;                         // key columns (first nkeys columns of cols) when ids == null
  int skip_empty;                    // skip rows whose first column equals BS_EMPTY_KEY
  const long long* ids;              // optional explicit destination per row
  const unsigned char* cols[XC_MAXC];
  int es[XC_MAXC];                   // bytes per row of each column
  int stride[XC_MAXC];               // bytes between consecutive rows (== es when dense);
                                     // larger for columns interleaved in a hash-table slot
};

// Per-exchange plan, written by xchg_plan, read by xchg_scatter and the host.
struct XcPlan {
  unsigned char* dst[XC_MAXP][XC_MAXC];          // where this rank's rows of shard d, col c go
  long long recv_total[XC_MAXP];                 // total rows of shard d (all sources)
  unsigned long long recv_off[XC_MAXP][XC_MAXC]; // byte offset of (d, c) region in the owner's arena
  unsigned long long arena_end[XC_MAXW];         // bytes used in each rank's arena
  long long send_count[XC_MAXP];                 // this rank's rows per shard
  unsigned long long cursor[XC_MAXP];            // running slot reservation (scatter)
  unsigned int done_ctas;
  int error;                                     // 1 overflow, 2 timeout/abort, 3 a rank reported
                                                 // failed producers, 4 ranks issued different invocations
};

__device__ __forceinline__ int xc_dest(const XcSpec& sp, long long i) {
  if (sp.ids) return (int)sp.ids[i];
  uint32_t h = 0;
#pragma unroll
  for (int k = 0; k < XC_MAXK; ++k) {
    if (k < sp.nkeys) {
      const unsigned char* p = sp.cols[k] + (size_t)i * sp.stride[k];
      if (sp.es[k] == 8) h ^= bs_murmur3_u64(*(const uint64_t*)p, 0u);
      else h ^= bs_murmur3_u32(*(const uint32_t*)p, 0u);
    }
  }
  return (int)(h % (uint32_t)sp.nparts);
}
__device__ __forceinline__ bool xc_valid(const XcSpec& sp, long long i) {
  return !sp.skip_empty || *(const long long*)(sp.cols[0] + (size_t)i * sp.stride[0]) != BS_EMPTY_KEY;
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(XC_THREADS) xchg_count_kernel(XcSpec sp, XcPlan* plan) {
  __shared__ int hist[XC_MAXP];
  for (int d = threadIdx.x; d < sp.nparts; d += blockDim.x) hist[d] = 0;
  __syncthreads();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < sp.n; i += stride)
    if (xc_valid(sp, i)) atomicAdd(&hist[xc_dest(sp, i)], 1);
  __syncthreads();
  for (int d = threadIdx.x; d < sp.nparts; d += blockDim.x)
    if (hist[d]) atomicAdd((unsigned long long*)&plan->send_count[d], (unsigned long long)hist[d]);
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) xchg_plan_kernel(XcPeers pe, XcSpec sp, XcPlan* plan,
                                                        unsigned long long epoch,
                                                        unsigned long long my_arena_base,
                                                        unsigned long long my_status,
                                                        unsigned long long my_fp) {
  const int W = pe.world, me = pe.rank, P = sp.nparts;
  const int par = (int)(epoch & 1);
  const int tid = threadIdx.x;
  // two verdicts, each written in one phase and read only after the barrier that ends it
  // (one variable for both let a thread that was already voting overwrite what a slower
  // thread was still reading as the flag-wait verdict — found by the CPU emulator)
  __shared__ int bad, vote_bad;
  if (tid == 0) {
    bad = 0;
    vote_bad = 0;
  }
  // 1. publish my counts row and arena base to every peer (P2P stores)
  for (int x = tid; x < W * P; x += blockDim.x) {
    int p = x / P, d = x - p * P;
    pe.ctrl[p]->cnt[par][me][d] = plan->send_count[d];
  }
  if (tid < W) {
    pe.ctrl[tid]->arena_base[par][me] = my_arena_base;
    pe.ctrl[tid]->vote[par][me][0] = my_status;
    pe.ctrl[tid]->vote[par][me][1] = my_fp;
  }
  __threadfence_system();
  __syncthreads();
  // 2. release flag on every peer, then 3. acquire all peers' flags
  if (tid < W) bs_st_release_sys(&pe.ctrl[tid]->count_flag[me], epoch);
  if (tid < W) {
    const unsigned long long* f = &pe.ctrl[me]->count_flag[tid];
    unsigned long long t0 = bs_globaltimer_ns();
    while (bs_ld_acquire_sys(f) < epoch) {
      if (bs_ld_relaxed_sys(&pe.ctrl[me]->abort_flag) != 0 ||
          bs_globaltimer_ns() - t0 > XC_TIMEOUT_NS) {
        atomicExch(&bad, 2);
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (bad) {
    if (tid == 0) {
      plan->error = 2;
      for (int p = 0; p < W; ++p) bs_st_release_sys(&pe.ctrl[p]->abort_flag, 1ULL);
    }
    return;
  }
  XcCtrl* mine = pe.ctrl[me];
  // 3b. the vote that came with the counts: every rank sees every rank's status
  //     and fingerprint, so all of them fail (or proceed) together — no host
  //     collective before the exchange.
  if (tid < W) {
    if (mine->vote[par][tid][0] != 0) atomicMax(&vote_bad, 3);
    else if (mine->vote[par][tid][1] != my_fp) atomicMax(&vote_bad, 4);
  }
  __syncthreads();
  if (vote_bad) {
    if (tid == 0) plan->error = vote_bad;
    return;
  }
  // 4. layout: shards of rank p are packed in increasing d, each (d, c) region
  //    256-byte aligned.  Every rank computes the same layout from the matrix.
  if (tid < P) {
    long long total = 0, before = 0;
    for (int s = 0; s < W; ++s) {
      long long c = mine->cnt[par][s][tid];
      if (s < me) before += c;
      total += c;
    }
    plan->recv_total[tid] = total;
    plan->cursor[tid] = (unsigned long long)before;  // my first slot inside shard tid
  }
  __syncthreads();
  if (tid < W) {
    const int p = tid;
    unsigned long long off = mine->arena_base[par][p];
    for (int d = p; d < P; d += W) {
      for (int c = 0; c < sp.ncols; ++c) {
        plan->recv_off[d][c] = off;
        plan->dst[d][c] = pe.arena[p] + off;
        off += (((unsigned long long)plan->recv_total[d] * (unsigned long long)sp.es[c]) + 255ULL) & ~255ULL;
      }
    }
    plan->arena_end[p] = off;
    if (off > pe.arena_bytes) atomicExch(&plan->error, 1);
  }
  if (tid == 0) plan->done_ctas = 0;
}

// ---------------------------------------------------------------------------
template <int VW>
__device__ __forceinline__ void xc_copy_chunk(unsigned char* dst, const unsigned char* src) {
  if (VW == 16) *(int4*)dst = *(const int4*)src;
  else if (VW == 8) *(long long*)dst = *(const long long*)src;
  else if (VW == 4) *(int*)dst = *(const int*)src;
  else if (VW == 2) *(short*)dst = *(const short*)src;
  else *dst = *src;
}

template <int VW>
__device__ __forceinline__ void xc_move_column(const XcSpec& sp, const XcPlan* plan, int c,
                                               long long row0, int rows, const int* sortpos,
                                               const unsigned short* pos_dest, const int* segstart,
                                               const unsigned long long* base,
                                               unsigned char* stage, unsigned char** sptr) {
  constexpr int PER = XC_TILE / XC_THREADS;
  const int es = sp.es[c];
  const int stride = sp.stride[c];
  const int nchunk = es / VW;
  // Per destination: where sorted position 0 of this tile would land in the
  // destination column (one pointer per destination, computed once per column),
  // so the flush loop is a table lookup + one multiply-add per element.
  for (int d = threadIdx.x; d < sp.nparts; d += XC_THREADS)
    sptr[d] = plan->dst[d][c] + ((long long)base[d] - (long long)segstart[d]) * (long long)es;
  for (int q = 0; q < nchunk; ++q) {
    // stage this chunk of the column in destination-sorted order
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      int r = threadIdx.x + j * XC_THREADS;
      if (r < rows && sortpos[j] >= 0)
        xc_copy_chunk<VW>(stage + (size_t)sortpos[j] * VW,
                          sp.cols[c] + (size_t)(row0 + r) * stride + (size_t)q * VW);
    }
    __syncthreads();
    const int nsorted = segstart[sp.nparts];
    const int qoff = q * VW;
#pragma unroll 2
    for (int pos = threadIdx.x; pos < nsorted; pos += XC_THREADS)
      xc_copy_chunk<VW>(sptr[pos_dest[pos]] + (long long)pos * es + qoff, stage + (size_t)pos * VW);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(XC_THREADS, 2) xchg_scatter_kernel(XcPeers pe, XcSpec sp, XcPlan* plan,
                                                                 unsigned long long epoch) {
  __shared__ int hist[XC_MAXP];
  __shared__ int segstart[XC_MAXP + 1];
  __shared__ unsigned long long base[XC_MAXP];
  __shared__ unsigned short pos_dest[XC_TILE];
  __shared__ unsigned char* sptr[XC_MAXP];
  __shared__ __align__(16) unsigned char stage[XC_TILE * 16];
  constexpr int PER = XC_TILE / XC_THREADS;
  const int P = sp.nparts;
  if (plan->error == 0) {
    const long long ntiles = (sp.n + XC_TILE - 1) / XC_TILE;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int d = threadIdx.x; d < P; d += blockDim.x) hist[d] = 0;
      __syncthreads();
      const long long row0 = tile * XC_TILE;
      const int rows = (int)(sp.n - row0 < XC_TILE ? sp.n - row0 : XC_TILE);
      int dest[PER], rank[PER], sortpos[PER];
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        int r = threadIdx.x + j * XC_THREADS;
        dest[j] = -1;
        if (r < rows && xc_valid(sp, row0 + r)) {
          dest[j] = xc_dest(sp, row0 + r);
          rank[j] = atomicAdd(&hist[dest[j]], 1);
        }
      }
      __syncthreads();
      // exclusive scan of hist (P <= 256: one warp, 8 entries per lane)
      if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int per = (P + 31) / 32;
        int local = 0;
        for (int k = 0; k < per; ++k) {
          int d = lane * per + k;
          if (d < P) local += hist[d];
        }
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        int run = incl - local;
        for (int k = 0; k < per; ++k) {
          int d = lane * per + k;
          if (d < P) {
            segstart[d] = run;
            run += hist[d];
          }
        }
        if (lane == 31) segstart[P] = incl;
      }
      __syncthreads();
      // one local atomic per destination reserves this tile's slot range
      for (int d = threadIdx.x; d < P; d += blockDim.x)
        if (hist[d]) base[d] = atomicAdd(&plan->cursor[d], (unsigned long long)hist[d]);
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        sortpos[j] = -1;
        if (dest[j] >= 0) {
          sortpos[j] = segstart[dest[j]] + rank[j];
          pos_dest[sortpos[j]] = (unsigned short)dest[j];
        }
      }
      __syncthreads();
      for (int c = 0; c < sp.ncols; ++c) {
        const int es = sp.es[c];
        const uintptr_t al = (uintptr_t)sp.cols[c] | (uintptr_t)sp.stride[c];
        const bool a16 = (al & 15) == 0;
        if (es % 16 == 0 && a16) xc_move_column<16>(sp, plan, c, row0, rows, sortpos, pos_dest, segstart, base, stage, sptr);
        else if (es % 8 == 0 && (al & 7) == 0) xc_move_column<8>(sp, plan, c, row0, rows, sortpos, pos_dest, segstart, base, stage, sptr);
        else if (es % 4 == 0 && (al & 3) == 0) xc_move_column<4>(sp, plan, c, row0, rows, sortpos, pos_dest, segstart, base, stage, sptr);
        else if (es % 2 == 0 && (al & 1) == 0) xc_move_column<2>(sp, plan, c, row0, rows, sortpos, pos_dest, segstart, base, stage, sptr);
        else xc_move_column<1>(sp, plan, c, row0, rows, sortpos, pos_dest, segstart, base, stage, sptr);
      }
    }
  }
  // completion: every CTA fences its peer stores; the last one signals peers.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned int prev = atomicAdd(&plan->done_ctas, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence_system();
      for (int p = 0; p < pe.world; ++p) bs_st_release_sys(&pe.ctrl[p]->data_flag[pe.rank], epoch);
    }
  }
}

__global__ void xchg_wait_kernel(XcPeers pe, XcPlan* plan, unsigned long long epoch) {
  const int tid = threadIdx.x;
  if (tid < pe.world) {
    const unsigned long long* f = &pe.ctrl[pe.rank]->data_flag[tid];
    unsigned long long t0 = bs_globaltimer_ns();
    while (bs_ld_acquire_sys(f) < epoch) {
      if (bs_ld_relaxed_sys(&pe.ctrl[pe.rank]->abort_flag) != 0 ||
          bs_globaltimer_ns() - t0 > XC_TIMEOUT_NS) {
        plan->error = 2;
        break;
      }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

// ---------------------------------------------------------------------------
// Hash-table exchange: the Reduce path shuffles an open-addressing table of
// 16-byte <key, value> slots in place.  Specialised kernels (one 16-byte load per
// slot, warp-aggregated ranking, optional scan-and-clear so the table is ready
// for its next use without an init pass) instead of the generic column mover.
#define XT_THREADS 256
#define XT_PER 4
#define XT_TILE (XT_THREADS * XT_PER)
#define XT_CTAS_PER_SM 4

struct __align__(16) XtSlot {
  long long key;
  unsigned long long val;
};

__device__ __forceinline__ int xt_dest(long long key, int P) {
  return (int)(bs_murmur3_u64((uint64_t)key, 0u) % (uint32_t)P);
}

// Adds 1 to hist[d] for every lane with a valid d, one shared atomic per
// distinct destination in the warp; returns the lane's rank among the rows of
// its destination seen so far in this CTA.
__device__ __forceinline__ int xt_rank(int* hist, int d, bool ok) {
  const unsigned lane = threadIdx.x & 31u;
  unsigned m = __match_any_sync(0xffffffffu, ok ? d : -1 - (int)lane);
  int basev = 0;
  const int leader = __ffs((int)m) - 1;
  if (ok && (int)lane == leader) basev = atomicAdd(&hist[d], __popc(m));
  basev = __shfl_sync(0xffffffffu, basev, leader);
  return basev + __popc(m & ((1u << lane) - 1u));
}

// Counting pass for tables whose combine kernels did not track destinations.
__global__ void __launch_bounds__(512) xchg_table_count_kernel(const XtSlot* __restrict__ slots,
                                                                long long total, int P, XcPlan* plan) {
  __shared__ int hist[XC_MAXP];
  for (int d = threadIdx.x; d < P; d += blockDim.x) hist[d] = 0;
  __syncthreads();
  const long long chunk = (long long)blockDim.x * 4;
  const long long nchunks = (total + chunk - 1) / chunk;
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    long long key[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long i = c * chunk + (long long)j * blockDim.x + threadIdx.x;
      key[j] = i < total ? slots[i].key : BS_EMPTY_KEY;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bool ok = key[j] != BS_EMPTY_KEY;
      if (__any_sync(0xffffffffu, ok)) xt_rank(hist, ok ? xt_dest(key[j], P) : 0, ok);
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < P; d += blockDim.x)
    if (hist[d]) atomicAdd((unsigned long long*)&plan->send_count[d], (unsigned long long)hist[d]);
}

// Scatter pass.  Four 256-thread CTAs per SM, each streaming 1024-slot tiles:
// the next tile's slots are loaded (one 16-byte load per slot) while the
// current tile is ranked, staged and flushed, so the DRAM latency of the scan
// is off the per-tile barrier chain.
template <int VB>
__global__ void __launch_bounds__(XT_THREADS, XT_CTAS_PER_SM) xchg_table_scatter_kernel(
    XcPeers pe, XtSlot* __restrict__ slots, long long total, int P, XcPlan* plan,
    unsigned long long epoch, int clear, unsigned long long clear_bits) {
  __shared__ int hist[XC_MAXP];
  __shared__ int segstart[XC_MAXP + 1];
  __shared__ unsigned char* kptr[XC_MAXP];
  __shared__ unsigned char* vptr[XC_MAXP];
  __shared__ long long skey[XT_TILE];
  __shared__ unsigned long long sval[XT_TILE];
  __shared__ unsigned short sdest[XT_TILE];
  if (plan->error == 0) {
    const long long ntiles = (total + XT_TILE - 1) / XT_TILE;
    int4 nxt[XT_PER];
    auto load_tile = [&](long long tile) {
#pragma unroll
      for (int j = 0; j < XT_PER; ++j) {
        const long long i = tile * XT_TILE + threadIdx.x + j * XT_THREADS;
        if (tile < ntiles && i < total) {
          nxt[j] = *(const int4*)&slots[i];
        } else {
          XtSlot e;
          e.key = BS_EMPTY_KEY;
          e.val = 0;
          memcpy(&nxt[j], &e, 16);
        }
      }
    };
    load_tile(blockIdx.x);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int d = threadIdx.x; d < P; d += blockDim.x) hist[d] = 0;
      __syncthreads();
      const long long row0 = tile * XT_TILE;
      long long key[XT_PER];
      unsigned long long val[XT_PER];
      int dest[XT_PER], rank[XT_PER];
#pragma unroll
      for (int j = 0; j < XT_PER; ++j) {
        XtSlot sl;
        memcpy(&sl, &nxt[j], 16);
        key[j] = sl.key;
        val[j] = sl.val;
      }
#pragma unroll
      for (int j = 0; j < XT_PER; ++j) {
        const bool ok = key[j] != BS_EMPTY_KEY;
        dest[j] = -1;
        if (__any_sync(0xffffffffu, ok)) {
          const int d = ok ? xt_dest(key[j], P) : 0;
          rank[j] = xt_rank(hist, d, ok);
          if (ok) {
            dest[j] = d;
            if (clear) {
              XtSlot z;
              z.key = BS_EMPTY_KEY;
              z.val = clear_bits;
              slots[row0 + threadIdx.x + j * XT_THREADS] = z;
            }
          }
        }
      }
      load_tile(tile + gridDim.x);  // prefetch: consumed in the next iteration
      __syncthreads();
      // exclusive scan of hist (P <= 256: one warp, <= 8 entries per lane)
      if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int per = (P + 31) / 32;
        int local = 0;
        for (int k = 0; k < per; ++k) {
          int d = lane * per + k;
          if (d < P) local += hist[d];
        }
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        int run = incl - local;
        for (int k = 0; k < per; ++k) {
          int d = lane * per + k;
          if (d < P) {
            segstart[d] = run;
            run += hist[d];
          }
        }
        if (lane == 31) segstart[P] = incl;
      }
      __syncthreads();
      // one local atomic per destination reserves this tile's slot range; the
      // per-destination pointers already include the (-segstart) shift so the
      // flush below is pointer[dest] + pos * width
      for (int d = threadIdx.x; d < P; d += blockDim.x) {
        if (hist[d]) {
          const long long b = (long long)atomicAdd(&plan->cursor[d], (unsigned long long)hist[d]) -
                              (long long)segstart[d];
          kptr[d] = plan->dst[d][0] + b * 8;
          vptr[d] = plan->dst[d][1] + b * VB;
        }
      }
#pragma unroll
      for (int j = 0; j < XT_PER; ++j) {
        if (dest[j] >= 0) {
          const int pos = segstart[dest[j]] + rank[j];
          skey[pos] = key[j];
          sval[pos] = val[j];
          sdest[pos] = (unsigned short)dest[j];
        }
      }
      __syncthreads();
      const int nsorted = segstart[P];
      for (int pos = threadIdx.x; pos < nsorted; pos += XT_THREADS) {
        const int d = sdest[pos];
        *(long long*)(kptr[d] + (long long)pos * 8) = skey[pos];
        if (VB == 8) *(unsigned long long*)(vptr[d] + (long long)pos * 8) = sval[pos];
        else *(unsigned int*)(vptr[d] + (long long)pos * 4) = (unsigned int)sval[pos];
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned int prev = atomicAdd(&plan->done_ctas, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence_system();
      for (int p = 0; p < pe.world; ++p) bs_st_release_sys(&pe.ctrl[p]->data_flag[pe.rank], epoch);
    }
  }
}

// Buffered scatter for few destinations (P <= 8, the one-shard-per-GPU case).
// Reserving a slot range per (tile, destination) costs one returning atomic on
// the same P cursor addresses per tile; those serialise in L2 (measured ~7 ns
// each: 35 k tiles -> 0.5 ms for a 36 M-slot table).  Here one persistent CTA per
// SM appends the rows it finds to per-destination buffers in shared memory
// (software write combining) and only reserves + writes out a destination's
// buffer when the next tile might not fit: one atomic per ~B rows, long
// coalesced stores over NVLink, and no global-latency operation on the
// per-tile path (the next tile's slots are prefetched into registers).
#define XB_THREADS 1024
#define XB_PER 4
#define XB_TILE (XB_THREADS * XB_PER)
#define XB_ROWS 8192  // staged rows per CTA: 8192 x 16 B = 128 KB of shared memory

template <int VB>
__device__ __forceinline__ void xb_store(const XcPlan* plan, int d, unsigned long long at, long long k,
                                         unsigned long long v) {
  ((long long*)plan->dst[d][0])[at] = k;
  if (VB == 8) ((unsigned long long*)plan->dst[d][1])[at] = v;
  else ((unsigned int*)plan->dst[d][1])[at] = (unsigned int)v;
}

template <int VB>
__global__ void __launch_bounds__(XB_THREADS, 1) xchg_table_scatter_buf_kernel(
    XcPeers pe, XtSlot* __restrict__ slots, long long total, int P, XcPlan* plan,
    unsigned long long epoch, int clear, unsigned long long clear_bits) {
  extern __shared__ __align__(16) unsigned char xb_smem[];
  long long* bkey = (long long*)xb_smem;                        // [P][B]
  unsigned long long* bval = (unsigned long long*)(bkey + XB_ROWS);
  __shared__ int hist[8], fill[8], off[8], flushn[8], direct[8];
  __shared__ unsigned long long gbase[8], dbase[8];
  const int B = XB_ROWS / P;
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u;
  const unsigned lt = (1u << lane) - 1u;
  if (tid < 8) fill[tid] = 0;
  if (plan->error == 0) {
    const long long ntiles = (total + XB_TILE - 1) / XB_TILE;
    int4 nxt[XB_PER];
    auto load_tile = [&](long long tile) {
#pragma unroll
      for (int j = 0; j < XB_PER; ++j) {
        const long long i = tile * XB_TILE + tid + j * XB_THREADS;
        if (tile < ntiles && i < total) {
          nxt[j] = *(const int4*)&slots[i];
        } else {
          XtSlot e;
          e.key = BS_EMPTY_KEY;
          e.val = 0;
          memcpy(&nxt[j], &e, 16);
        }
      }
    };
    load_tile(blockIdx.x);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      if (tid < 8) hist[tid] = 0;
      __syncthreads();
      long long key[XB_PER];
      unsigned long long val[XB_PER];
      int dest[XB_PER], rank[XB_PER];
#pragma unroll
      for (int j = 0; j < XB_PER; ++j) {
        XtSlot sl;
        memcpy(&sl, &nxt[j], 16);
        key[j] = sl.key;
        val[j] = sl.val;
      }
      load_tile(tile + gridDim.x);  // prefetch: consumed in the next iteration
      // Warp-local ranking with ballots only (P <= 8): every lane learns how many
      // of the warp's rows so far went to each destination; one shared atomic per
      // (warp, destination) for the whole tile reserves the warp's range.
      int mine = 0;  // rows of lane-th destination in this warp's part of the tile
#pragma unroll
      for (int j = 0; j < XB_PER; ++j) {
        const bool ok = key[j] != BS_EMPTY_KEY;
        const int d = ok ? xt_dest(key[j], P) : -1;
        dest[j] = d;
        rank[j] = 0;
        for (int q = 0; q < P; ++q) {
          const unsigned bal = __ballot_sync(0xffffffffu, d == q);
          // rows of destination q seen in earlier j's are tracked by lane q
          const int before = __shfl_sync(0xffffffffu, mine, q);
          if (d == q) rank[j] = before + __popc(bal & lt);
          if ((int)lane == q) mine += __popc(bal);
        }
        if (ok && clear) {
          XtSlot z;
          z.key = BS_EMPTY_KEY;
          z.val = clear_bits;
          slots[tile * XB_TILE + tid + j * XB_THREADS] = z;
        }
      }
      int wbase = 0;
      if ((int)lane < P && mine) wbase = atomicAdd(&hist[lane], mine);
#pragma unroll
      for (int j = 0; j < XB_PER; ++j) {
        const int bq = __shfl_sync(0xffffffffu, wbase, dest[j] < 0 ? 0 : dest[j]);
        rank[j] += bq;
      }
      __syncthreads();
      // per destination: append behind what is buffered, flush first if it would
      // not fit, or (a tile with more rows for one destination than a whole
      // buffer) write this tile's rows directly
      int need = 0;
      if (tid < P) {
        const int f = fill[tid], h = hist[tid];
        const int dir = h > B;             // does not fit even in an empty buffer
        need = (f + h > B) && f > 0;       // flush what is buffered first
        flushn[tid] = need ? f : 0;
        if (need) gbase[tid] = atomicAdd(&plan->cursor[tid], (unsigned long long)f);
        direct[tid] = dir;
        if (dir) dbase[tid] = atomicAdd(&plan->cursor[tid], (unsigned long long)h);
        off[tid] = need ? 0 : f;
        fill[tid] = dir ? 0 : (need ? 0 : f) + h;
      }
      if (__syncthreads_or(need)) {
        for (int d = 0; d < P; ++d) {
          const int n = flushn[d];
          if (n == 0) continue;
          const long long* ks = bkey + (size_t)d * B;
          const unsigned long long* vs = bval + (size_t)d * B;
          for (int i = tid; i < n; i += XB_THREADS) xb_store<VB>(plan, d, gbase[d] + i, ks[i], vs[i]);
        }
        __syncthreads();
      }
#pragma unroll
      for (int j = 0; j < XB_PER; ++j) {
        const int d = dest[j];
        if (d >= 0) {
          if (direct[d]) {
            xb_store<VB>(plan, d, dbase[d] + rank[j], key[j], val[j]);
          } else {
            const int pos = d * B + off[d] + rank[j];
            bkey[pos] = key[j];
            bval[pos] = val[j];
          }
        }
      }
    }
    // drain what is still buffered
    __syncthreads();
    if (tid < P) {
      flushn[tid] = fill[tid];
      if (fill[tid]) gbase[tid] = atomicAdd(&plan->cursor[tid], (unsigned long long)fill[tid]);
    }
    __syncthreads();
    for (int d = 0; d < P; ++d) {
      const int n = flushn[d];
      if (n == 0) continue;
      const long long* ks = bkey + (size_t)d * B;
      const unsigned long long* vs = bval + (size_t)d * B;
      for (int i = tid; i < n; i += XB_THREADS) xb_store<VB>(plan, d, gbase[d] + i, ks[i], vs[i]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned int prev = atomicAdd(&plan->done_ctas, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence_system();
      for (int p = 0; p < pe.world; ++p) bs_st_release_sys(&pe.ctrl[p]->data_flag[pe.rank], epoch);
    }
  }
}

template <int VB>
static int xb_launch(const XcPeers& pe, XtSlot* tab, long long total, int nparts, XcPlan* plan,
                     unsigned long long epoch, int clear, unsigned long long clear_bits, int sms,
                     cudaStream_t stream) {
  const int smem = XB_ROWS * 16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(xchg_table_scatter_buf_kernel<VB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  long long ntiles = (total + XB_TILE - 1) / XB_TILE;
  int blocks = (int)(ntiles < sms ? ntiles : sms);
  if (blocks < 1) blocks = 1;
  xchg_table_scatter_buf_kernel<VB><<<blocks, XB_THREADS, smem, stream>>>(pe, tab, total, nparts, plan, epoch,
                                                                           clear, clear_bits);
  BS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// Row-record exchange (generic frames, v2).  Measured on the column mover above:
// 371 GB/s = 48 % of the 770 GB/s peer-copy figure, issue-bound on 4-byte chunks.
// Here every row is assembled once, in shared memory, into a 16-byte-aligned
// *record* (all its columns back to back), appended to a per-destination staging
// buffer (software write combining, as in the Reduce scatter) and full buffers
// leave the SM as ONE TMA bulk store (cp.async.bulk shared -> global) into the
// destination GPU's arena: NVLink sees long, aligned bursts whatever the column
// widths are.  The receiver turns records back into columns with a local pass
// (xchg_unpack_rows: HBM-speed, coalesced column stores).
#define XR_THREADS 1024
#define XR_PER 2
#define XR_TILE (XR_THREADS * XR_PER)
#define XR_BYTES (128 * 1024)
#define XR_MAXP 64

struct XcRec {
  int rec_bytes;          // bytes per record (multiple of 16)
  int off[XC_MAXC];       // byte offset of every column inside a record
};

__global__ void __launch_bounds__(XR_THREADS, 1) xchg_scatter_rows_kernel(XcPeers pe, XcSpec sp, XcRec rc,
                                                                          XcPlan* plan,
                                                                          unsigned long long epoch) {
  extern __shared__ __align__(128) unsigned char xr_stage[];
  __shared__ int hist[XR_MAXP], fill[XR_MAXP], took[XR_MAXP];
  __shared__ int pending;
  const int P = sp.nparts;
  const int tid = threadIdx.x;
  const int rec = rc.rec_bytes;
  const int B = XR_BYTES / (P * rec);  // records per destination buffer
  bool small_cols = true;
  for (int c = 0; c < sp.ncols; ++c) small_cols = small_cols && (sp.es[c] == 4 || sp.es[c] == 8);
  if (tid < XR_MAXP) fill[tid] = 0;
  __syncthreads();
  auto flush = [&](int d, int n) {  // one thread: n staged records of destination d -> peer memory
    const unsigned long long at = atomicAdd(&plan->cursor[d], (unsigned long long)n);
    bs_tma_store_1d(plan->dst[d][0] + at * (unsigned long long)rec, xr_stage + (size_t)d * B * rec,
                    (uint32_t)n * (uint32_t)rec);
    bs_tma_store_commit();
    bs_tma_store_wait_read();
  };
  if (plan->error == 0) {
    const long long ntiles = (sp.n + XR_TILE - 1) / XR_TILE;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      if (tid < P) hist[tid] = 0;
      __syncthreads();
      const long long row0 = tile * XR_TILE;
      int dest[XR_PER], rank[XR_PER];
#pragma unroll
      for (int j = 0; j < XR_PER; ++j) {
        const long long r = row0 + tid + (long long)j * XR_THREADS;
        const bool ok = r < sp.n && xc_valid(sp, r);
        dest[j] = -1;
        rank[j] = 0;
        if (__any_sync(0xffffffffu, ok)) {
          const int d = ok ? xc_dest(sp, r) : 0;
          rank[j] = xt_rank(hist, d, ok);
          if (ok) dest[j] = d;
        }
      }
      __syncthreads();
      // Append in rounds: a round takes, per destination, as many of the tile's rows
      // as its buffer has room for; full buffers are flushed between rounds.  The
      // uniform case is one round and one flush every B/hist tiles.
      if (tid < P) took[tid] = 0;
      while (true) {
        __syncthreads();
        if (tid == 0) pending = 0;
        __syncthreads();
        int room = 0, lo = 0;
#pragma unroll
        for (int j = 0; j < XR_PER; ++j) {
          const int d = dest[j];
          if (d < 0) continue;
          lo = took[d];
          room = B - fill[d];
          const int k = rank[j] - lo;  // position among the rows of d not yet appended
          if (k >= 0 && k < room) {
            unsigned char* out = xr_stage + ((size_t)d * B + fill[d] + k) * rec;
            const long long r = row0 + tid + (long long)j * XR_THREADS;
            if (small_cols) {
              // all columns are 4 or 8 bytes wide: issue every column's load before
              // the first store (independent loads in flight instead of a chain of
              // dependent round trips through the loop)
              unsigned long long v[XC_MAXC];
#pragma unroll
              for (int c = 0; c < XC_MAXC; ++c) {
                if (c < sp.ncols) {
                  const unsigned char* src = sp.cols[c] + (size_t)r * sp.stride[c];
                  v[c] = sp.es[c] == 8 ? *(const unsigned long long*)src
                                       : (unsigned long long)*(const unsigned int*)src;
                }
              }
#pragma unroll
              for (int c = 0; c < XC_MAXC; ++c) {
                if (c < sp.ncols) {
                  if (sp.es[c] == 8) *(unsigned long long*)(out + rc.off[c]) = v[c];
                  else *(unsigned int*)(out + rc.off[c]) = (unsigned int)v[c];
                }
              }
            } else {
              for (int c = 0; c < sp.ncols; ++c) {
                const unsigned char* src = sp.cols[c] + (size_t)r * sp.stride[c];
                const int es = sp.es[c];
                unsigned char* o = out + rc.off[c];
                if (es == 4) *(int*)o = *(const int*)src;
                else if (es == 8) *(long long*)o = *(const long long*)src;
                else if (es == 16) *(int4*)o = *(const int4*)src;
                else if (es % 16 == 0 && ((uintptr_t)src & 15) == 0)
                  for (int q = 0; q < es; q += 16) *(int4*)(o + q) = *(const int4*)(src + q);
                else if (es % 4 == 0 && ((uintptr_t)src & 3) == 0)
                  for (int q = 0; q < es; q += 4) *(int*)(o + q) = *(const int*)(src + q);
                else
                  for (int q = 0; q < es; ++q) o[q] = src[q];
              }
            }
            dest[j] = -1;  // done
          } else if (k >= room) {
            pending = 1;
          }
        }
        bs_fence_proxy_async();
        __syncthreads();
        if (tid < P) {
          const int h = hist[tid] - took[tid];          // rows of this destination still to place
          const int r = B - fill[tid];
          const int t = h < r ? h : r;
          took[tid] += t;
          fill[tid] += t;
          if (fill[tid] == B || (pending && fill[tid] > 0 && h > r)) {
            flush(tid, fill[tid]);
            fill[tid] = 0;
          }
        }
        __syncthreads();
        if (!pending) break;
      }
    }
    __syncthreads();
    if (tid < P && fill[tid] > 0) flush(tid, fill[tid]);
  }
  if (tid < P) {
    bs_tma_store_wait_all();
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    const unsigned int prev = atomicAdd(&plan->done_ctas, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence_system();
      for (int p = 0; p < pe.world; ++p) bs_st_release_sys(&pe.ctrl[p]->data_flag[pe.rank], epoch);
    }
  }
}

// Receiver: records of one shard -> its columns.
__global__ void __launch_bounds__(256) xchg_unpack_rows_kernel(const unsigned char* __restrict__ recs,
                                                               long long n, XcRec rc, int ncols,
                                                               XcSpec out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const unsigned char* r = recs + (size_t)i * rc.rec_bytes;
    for (int c = 0; c < ncols; ++c) {
      const int es = out.es[c];
      unsigned char* o = (unsigned char*)out.cols[c] + (size_t)i * es;
      const unsigned char* src = r + rc.off[c];
      if (es == 4) *(int*)o = *(const int*)src;
      else if (es == 8) *(long long*)o = *(const long long*)src;
      else if (es % 16 == 0)
        for (int q = 0; q < es; q += 16) *(int4*)(o + q) = *(const int4*)(src + q);
      else if (es % 4 == 0)
        for (int q = 0; q < es; q += 4) *(int*)(o + q) = *(const int*)(src + q);
      else
        for (int q = 0; q < es; ++q) o[q] = src[q];
    }
  }
}

// In-kernel barrier across ranks (flag all-to-all), used to fence arena reuse
// and for device-side timing alignment.  No NCCL.
__global__ void xchg_barrier_kernel(XcPeers pe, unsigned long long epoch, int* err) {
  const int tid = threadIdx.x;
  __threadfence_system();
  if (tid < pe.world) bs_st_release_sys(&pe.ctrl[tid]->barrier_flag[pe.rank], epoch);
  if (tid < pe.world) {
    const unsigned long long* f = &pe.ctrl[pe.rank]->barrier_flag[tid];
    unsigned long long t0 = bs_globaltimer_ns();
    while (bs_ld_acquire_sys(f) < epoch) {
      if (bs_globaltimer_ns() - t0 > XC_TIMEOUT_NS) {
        if (err) *err = 2;
        break;
      }
      __nanosleep(200);
    }
  }
}

// ---------------------------------------------------------------------------
// host API
BS_EXPORT int bs_xc_sizes(int* ctrl_bytes, int* plan_bytes, int* peers_bytes, int* spec_bytes) {
  *ctrl_bytes = (int)sizeof(XcCtrl);
  *plan_bytes = (int)sizeof(XcPlan);
  *peers_bytes = (int)sizeof(XcPeers);
  *spec_bytes = (int)sizeof(XcSpec);
  return 0;
}
BS_EXPORT int bs_xc_limits(int* maxw, int* maxp, int* maxc, int* maxk) {
  *maxw = XC_MAXW; *maxp = XC_MAXP; *maxc = XC_MAXC; *maxk = XC_MAXK;
  return 0;
}

// Offsets inside XcPlan for the host to read results without a header.
BS_EXPORT int bs_xc_plan_offsets(long long* recv_total, long long* recv_off, long long* arena_end,
                                 long long* send_count, long long* error) {
  XcPlan* p = nullptr;
  *recv_total = (long long)(uintptr_t)&p->recv_total[0];
  *recv_off = (long long)(uintptr_t)&p->recv_off[0][0];
  *arena_end = (long long)(uintptr_t)&p->arena_end[0];
  *send_count = (long long)(uintptr_t)&p->send_count[0];
  *error = (long long)(uintptr_t)&p->error;
  return 0;
}

BS_EXPORT int bs_ipc_alloc(long long bytes, void** ptr, unsigned char* handle64) {
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(*ptr, 0, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) return (int)e;
  memcpy(handle64, &h, sizeof(h));
  return 0;
}
BS_EXPORT int bs_ipc_open(const unsigned char* handle64, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}
BS_EXPORT int bs_ipc_close(void* ptr) { return (int)cudaIpcCloseMemHandle(ptr); }
BS_EXPORT int bs_dev_free(void* ptr) { return (int)cudaFree(ptr); }
BS_EXPORT int bs_dev_alloc(long long bytes, void** ptr) {
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaMemset(*ptr, 0, (size_t)bytes);
}
BS_EXPORT int bs_enable_peer(int peer_device) {
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return 0;
  }
  return (int)e;
}
BS_EXPORT int bs_memcpy_d2d(void* dst, const void* src, long long bytes, cudaStream_t stream) {
  return (int)cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, stream);
}
BS_EXPORT int bs_memcpy_d2h(void* dst, const void* src, long long bytes, cudaStream_t stream) {
  return (int)cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, stream);
}
BS_EXPORT int bs_memset(void* dst, int v, long long bytes, cudaStream_t stream) {
  return (int)cudaMemsetAsync(dst, v, (size_t)bytes, stream);
}
BS_EXPORT int bs_stream_sync(cudaStream_t stream) { return (int)cudaStreamSynchronize(stream); }

// One full exchange enqueued on `stream`.  peers/spec are host structs laid out
// exactly as XcPeers/XcSpec (built by the python side with ctypes).
BS_EXPORT int bs_xc_exchange(const void* peers_h, const void* spec_h, void* plan_d,
                             unsigned long long epoch, unsigned long long arena_base,
                             unsigned long long my_status, unsigned long long my_fp,
                             cudaStream_t stream) {
  XcPeers pe = *(const XcPeers*)peers_h;
  XcSpec sp = *(const XcSpec*)spec_h;
  for (int c = 0; c < XC_MAXC; ++c)
    if (sp.stride[c] == 0) sp.stride[c] = sp.es[c];
  if (pe.world > XC_MAXW || sp.nparts > XC_MAXP || sp.ncols > XC_MAXC || sp.nkeys > XC_MAXK) return -1;
  XcPlan* plan = (XcPlan*)plan_d;
  cudaError_t e = cudaMemsetAsync(plan, 0, sizeof(XcPlan), stream);
  if (e != cudaSuccess) return (int)e;
  int sms = bs_num_sms();
  if (sp.n > 0) {
    long long want = (sp.n + XC_THREADS - 1) / XC_THREADS;
    int blocks = (int)(want < (long long)sms * 4 ? want : (long long)sms * 4);
    xchg_count_kernel<<<blocks, XC_THREADS, 0, stream>>>(sp, plan);
    BS_LAUNCH_CHECK();
  }
  xchg_plan_kernel<<<1, 256, 0, stream>>>(pe, sp, plan, epoch, arena_base, my_status, my_fp);
  BS_LAUNCH_CHECK();
  long long ntiles = (sp.n + XC_TILE - 1) / XC_TILE;
  int blocks = (int)(ntiles < (long long)sms * 2 ? ntiles : (long long)sms * 2);
  if (blocks < 1) blocks = 1;
  xchg_scatter_kernel<<<blocks, XC_THREADS, 0, stream>>>(pe, sp, plan, epoch);
  BS_LAUNCH_CHECK();
  xchg_wait_kernel<<<1, 32, 0, stream>>>(pe, plan, epoch);
  BS_LAUNCH_CHECK();
  return 0;
}

static void xc_make_rec(const XcSpec& sp, XcRec* rc) {
  int off = 0;
  for (int c = 0; c < sp.ncols; ++c) {
    const int es = sp.es[c];
    const int al = es % 16 == 0 ? 16 : (es % 8 == 0 ? 8 : (es % 4 == 0 ? 4 : 1));
    off = (off + al - 1) / al * al;
    rc->off[c] = off;
    off += es;
  }
  rc->rec_bytes = (off + 15) / 16 * 16;
}

BS_EXPORT int bs_xc_rec_bytes(const void* spec_h) {
  XcSpec sp = *(const XcSpec*)spec_h;
  XcRec rc;
  xc_make_rec(sp, &rc);
  return rc.rec_bytes;
}

// Row-record exchange: like bs_xc_exchange, but rows travel as records (see
// xchg_scatter_rows_kernel); shard d's records land at plan->recv_off[d][0].
BS_EXPORT int bs_xc_exchange_rows(const void* peers_h, const void* spec_h, void* plan_d,
                                  unsigned long long epoch, unsigned long long my_status,
                                  unsigned long long my_fp, cudaStream_t stream) {
  XcPeers pe = *(const XcPeers*)peers_h;
  XcSpec sp = *(const XcSpec*)spec_h;
  for (int c = 0; c < XC_MAXC; ++c)
    if (sp.stride[c] == 0) sp.stride[c] = sp.es[c];
  if (pe.world > XC_MAXW || sp.nparts > XR_MAXP || sp.ncols > XC_MAXC || sp.nkeys > XC_MAXK) return -1;
  XcRec rc;
  xc_make_rec(sp, &rc);
  if ((long long)sp.nparts * rc.rec_bytes * 8 > XR_BYTES) return -3;  // records too wide for the staging buffers
  XcPlan* plan = (XcPlan*)plan_d;
  cudaError_t e = cudaMemsetAsync(plan, 0, sizeof(XcPlan), stream);
  if (e != cudaSuccess) return (int)e;
  int sms = bs_num_sms();
  if (sp.n > 0) {
    long long want = (sp.n + XC_THREADS - 1) / XC_THREADS;
    int blocks = (int)(want < (long long)sms * 4 ? want : (long long)sms * 4);
    xchg_count_kernel<<<blocks, XC_THREADS, 0, stream>>>(sp, plan);
    BS_LAUNCH_CHECK();
  }
  XcSpec lay = sp;  // the layout the plan computes: one "column" of record bytes per shard
  lay.ncols = 1;
  lay.es[0] = rc.rec_bytes;
  xchg_plan_kernel<<<1, 256, 0, stream>>>(pe, lay, plan, epoch, 0ULL, my_status, my_fp);
  BS_LAUNCH_CHECK();
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(xchg_scatter_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XR_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  long long ntiles = (sp.n + XR_TILE - 1) / XR_TILE;
  int blocks = (int)(ntiles < sms ? ntiles : sms);
  if (blocks < 1) blocks = 1;
  xchg_scatter_rows_kernel<<<blocks, XR_THREADS, XR_BYTES, stream>>>(pe, sp, rc, plan, epoch);
  BS_LAUNCH_CHECK();
  xchg_wait_kernel<<<1, 32, 0, stream>>>(pe, plan, epoch);
  BS_LAUNCH_CHECK();
  return 0;
}

// Records of one received shard -> dense columns (out_spec.cols = destinations,
// out_spec.es = element sizes, same column order as the sender's spec).
BS_EXPORT int bs_xc_unpack_rows(const void* recs, long long n, const void* out_spec_h, cudaStream_t stream) {
  XcSpec out = *(const XcSpec*)out_spec_h;
  if (n <= 0) return 0;
  XcRec rc;
  xc_make_rec(out, &rc);
  long long want = (n + 255) / 256;
  long long cap = (long long)bs_num_sms() * 8;
  int blocks = (int)(want < cap ? want : cap);
  xchg_unpack_rows_kernel<<<blocks, 256, 0, stream>>>((const unsigned char*)recs, n, rc, out.ncols, out);
  BS_LAUNCH_CHECK();
  return 0;
}

// Exchange of a hash table of `total` 16-byte slots (key, value bits): rows of
// shard d = murmur3(key) % nparts land in the owner's arena as two dense columns
// (int64 keys, vbytes-wide values) exactly like bs_xc_exchange with
// skip_empty.  clear != 0 resets every occupied slot to <EMPTY, clear_bits>.
// precounts (optional, device): exact per-shard row counts gathered while the
// table was built; without them a counting pass scans the table first.
BS_EXPORT int bs_xc_exchange_table(const void* peers_h, void* slots, long long total, int vbytes,
                                   int nparts, void* plan_d, unsigned long long epoch,
                                   unsigned long long arena_base, int clear,
                                   unsigned long long clear_bits,
                                   const unsigned long long* precounts,
                                   unsigned long long my_status, unsigned long long my_fp,
                                   cudaStream_t stream) {
  XcPeers pe = *(const XcPeers*)peers_h;
  if (pe.world > XC_MAXW || nparts > XC_MAXP || nparts < 1 || (vbytes != 4 && vbytes != 8)) return -1;
  XcSpec sp;
  memset(&sp, 0, sizeof(sp));
  sp.n = total;
  sp.nparts = nparts;
  sp.ncols = 2;
  sp.nkeys = 1;
  sp.skip_empty = 1;
  sp.cols[0] = (const unsigned char*)slots;
  sp.cols[1] = (const unsigned char*)slots + 8;
  sp.es[0] = 8;
  sp.es[1] = vbytes;
  sp.stride[0] = sp.stride[1] = 16;
  XcPlan* plan = (XcPlan*)plan_d;
  cudaError_t e = cudaMemsetAsync(plan, 0, sizeof(XcPlan), stream);
  if (e != cudaSuccess) return (int)e;
  int sms = bs_num_sms();
  XtSlot* tab = (XtSlot*)slots;
  if (precounts != nullptr) {
    // destinations were counted while the table was built (bs_hash_combine)
    e = cudaMemcpyAsync(&plan->send_count[0], precounts, (size_t)nparts * sizeof(long long),
                        cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return (int)e;
  } else if (total > 0) {
    long long want = (total + 2047) / 2048;
    int blocks = (int)(want < (long long)sms * 4 ? want : (long long)sms * 4);
    xchg_table_count_kernel<<<blocks, 512, 0, stream>>>(tab, total, nparts, plan);
    BS_LAUNCH_CHECK();
  }
  xchg_plan_kernel<<<1, 256, 0, stream>>>(pe, sp, plan, epoch, arena_base, my_status, my_fp);
  BS_LAUNCH_CHECK();
  if (nparts <= 8) {
    // few destinations: software write-combining in shared memory
    int rc = vbytes == 8 ? xb_launch<8>(pe, tab, total, nparts, plan, epoch, clear, clear_bits, sms, stream)
                         : xb_launch<4>(pe, tab, total, nparts, plan, epoch, clear, clear_bits, sms, stream);
    if (rc != 0) return rc;
  } else {
    long long ntiles = (total + XT_TILE - 1) / XT_TILE;
    int blocks = (int)(ntiles < (long long)sms * XT_CTAS_PER_SM ? ntiles : (long long)sms * XT_CTAS_PER_SM);
    if (blocks < 1) blocks = 1;
    if (vbytes == 8)
      xchg_table_scatter_kernel<8><<<blocks, XT_THREADS, 0, stream>>>(pe, tab, total, nparts, plan, epoch, clear, clear_bits);
    else
      xchg_table_scatter_kernel<4><<<blocks, XT_THREADS, 0, stream>>>(pe, tab, total, nparts, plan, epoch, clear, clear_bits);
    BS_LAUNCH_CHECK();
  }
  xchg_wait_kernel<<<1, 32, 0, stream>>>(pe, plan, epoch);
  BS_LAUNCH_CHECK();
  return 0;
}

BS_EXPORT int bs_xc_barrier(const void* peers_h, unsigned long long epoch, int* err_d,
                            cudaStream_t stream) {
  XcPeers pe = *(const XcPeers*)peers_h;
  xchg_barrier_kernel<<<1, 32, 0, stream>>>(pe, epoch, err_d);
  BS_LAUNCH_CHECK();
  return 0;
}
