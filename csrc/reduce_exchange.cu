// K3: the Reduce shuffle — fused table scatter -> NVLink -> combine on arrival.
//
// The producer side of a Reduce has hash-aggregated its rows into an HBM table
// (hash_combine.cu).  This file moves that table to the owners of the keys and
// combines what arrives, with **no count exchange, no host round trip and no
// NCCL call**:
//
//   rx_begin    (1 CTA)   bumps the device-side epoch, resets the per-destination
//                         cursors and waits (bounded) until every destination has
//                         acknowledged the landing buffers of epoch-2 (same parity).
//   rx_scatter  (1 CTA/SM) scans the local table (one 16-byte load per slot, EMPTY
//                         slots skipped = fused compaction, occupied slots cleared =
//                         the table is ready for its next use), appends rows to
//                         per-destination staging buffers in 128 KB of shared
//                         memory and flushes full buffers with ONE TMA bulk store
//                         (cp.async.bulk shared -> global, SASS UBLKCP) straight into
//                         the destination GPU's landing region over NVLink.  Every
//                         (source, shard) pair owns a fixed-capacity region, so a
//                         source reserves space with a *local* cursor and never
//                         needs the other ranks' counts.  The last CTA publishes the
//                         row counts + the vote (producer status, invocation
//                         fingerprint) to the destinations and releases the data
//                         flags at system scope.
//   rx_wait     (1 CTA)   acquire-spins (bounded) for all sources' flags, folds the
//                         votes into one verdict every rank agrees on.
//   rx_combine  (many)    combine on arrival: rows of all sources are inserted into
//                         the destination's final table with L2 atomics.
//   (table_compact from hash_combine.cu turns the table into dense columns)
//   rx_end      (1 CTA)   acknowledges the landing buffers to every source and
//                         writes the report (error, counts) to pinned host memory,
//                         which the host reads only when somebody needs the sizes.
//
// Reference behaviour replaced: worker.runCombine + CommitCombiner + the pull
// shuffle + sortio.Reduce on the consumer (exec/bigmachine.go:1084-1210, 857-1005;
// sortio/reader.go:36-130).
#include "table.cuh"

#define RX_MAXW 16
#define RX_MAXP 64
// Two CTAs per SM (512 threads, 64 KB of staging each): the per-tile barrier chain
// (rank -> reserve -> stage) of one CTA overlaps the other's.
#define RX_THREADS 512
#define RX_PER 4
#define RX_TILE (RX_THREADS * RX_PER)
#define RX_ROWS 4096  // staged rows per CTA (16 B each = 64 KB of shared memory)
#define RX_CTAS_PER_SM 2

enum RxError {
  RX_OK = 0,
  RX_LANDING_OVERFLOW = 1,  // a (source, shard) landing region was too small
  RX_TABLE_OVERFLOW = 2,    // a producer-side table overflowed on some rank
  RX_PRODUCER_FAILED = 3,   // a rank reported failed producers
  RX_FINGERPRINT = 4,       // ranks issued different invocations
  RX_TIMEOUT = 5,           // a peer did not show up (or aborted)
  RX_DEST_OVERFLOW = 6,     // destination table / output buffer too small on some rank
};

// Symmetric control block: one per rank, mapped by every peer (CUDA IPC).
struct RxCtrl {
  unsigned long long data_flag[RX_MAXW];  // [src] last epoch whose rows from src are complete here
  unsigned long long ack[RX_MAXW];        // [dst] last epoch dst has finished reading (written by dst)
  unsigned long long fin[RX_MAXW];        // [src] (epoch << 8) | src's outcome of that epoch (see rx_end)
  unsigned long long abort_flag;
  unsigned long long pad[7];
  long long cnt[2][RX_MAXW][RX_MAXP];      // [parity][src][shard] rows written
  unsigned long long vote[2][RX_MAXW][4];  // [parity][src] {status, fingerprint, need_rows, -}
};

struct RxPeers {
  int world, rank;
  RxCtrl* ctrl[RX_MAXW];
  unsigned char* arena[RX_MAXW];
  unsigned long long region_rows;  // capacity C of one (parity, local shard, source) region
  unsigned long long half_bytes;   // the arena is split into two parity halves of this size
  int shards_per_rank;             // Pl * R: landing regions per (parity, source)
  int regions;                     // R (power of two): table regions per destination shard
};

// Per-rank private device state of the exchange in flight.
struct RxState {
  unsigned long long epoch;
  unsigned long long cursor[RX_MAXP];
  unsigned long long need_rows;
  unsigned int done_ctas;
  int error;
  long long recv[RX_MAXP][RX_MAXW];  // [local shard j][src]
  long long sent_rows;
  unsigned long long timeout_ns;
  unsigned long long vote_in[2];     // {producer status, invocation fingerprint} of the next exchange
  unsigned long long publish_seq;    // tickets of purely local publishes
};

// What the host learns about a finished exchange (pinned, written by rx_end).
struct RxReport {
  unsigned long long epoch;
  long long error;
  long long need_rows;
  long long sent_rows;
  long long recv_rows[RX_MAXP];
  long long out_rows[RX_MAXP];
  long long dest_failed[RX_MAXP];
};

// Virtual partitions.  A destination shard d is split into R *table regions*
// (r = a few bits of the table hash): v = d * R + r.  Rows of one region land
// together, and the destination combines region after region, each region's part
// of the table (tens of MB) staying L2-resident — instead of random DRAM accesses
// over the whole table when it does not fit in L2 (few ranks, many keys each).
__device__ __forceinline__ int rx_vowner(const RxPeers& pe, int v) { return (v / pe.regions) % pe.world; }
__device__ __forceinline__ int rx_vlocal(const RxPeers& pe, int v) {
  return ((v / pe.regions) / pe.world) * pe.regions + (v % pe.regions);
}
__device__ __forceinline__ int rx_table_region(long long key, int R) {
  return (int)(bs_mix_u64((uint64_t)key) & (uint32_t)(R - 1));
}

__device__ __forceinline__ unsigned char* rx_region(const RxPeers& pe, int dst_rank, int par, int j,
                                                    int src) {
  const unsigned long long idx = (unsigned long long)j * (unsigned long long)pe.world + src;
  return pe.arena[dst_rank] + (unsigned long long)par * pe.half_bytes + idx * pe.region_rows * 16ULL;
}

// Destination shard of a packed key (bigslice_b200/ops/keypack.py): the same
// value Frame.hash computes from the key columns — mode 0: one 8-byte column,
// mode 1: one column of <= 4 bytes (its 32-bit image), mode 2: two such columns
// (XOR of the per-column hashes, frame/frame.go:389-401).
__device__ __forceinline__ int rx_dest(long long key, int P, int mode) {
  uint32_t h;
  if (mode == 0) h = bs_murmur3_u64((uint64_t)key, 0u);
  else if (mode == 1) h = bs_murmur3_u32((uint32_t)key, 0u);
  else h = bs_murmur3_u32((uint32_t)key, 0u) ^ bs_murmur3_u32((uint32_t)((uint64_t)key >> 32), 0u);
  return (int)(h % (uint32_t)P);
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(64) rx_begin_kernel(RxPeers pe, RxState* st, int P) {
  __shared__ int bad;
  const int tid = threadIdx.x;
  if (tid == 0) {
    bad = 0;
    st->epoch += 1;
    st->error = 0;
    st->done_ctas = 0;
    st->need_rows = 0;
    st->sent_rows = 0;
  }
  if (tid < RX_MAXP) st->cursor[tid] = 0;
  __syncthreads();
  const unsigned long long epoch = st->epoch;
  if (tid < pe.world && epoch > 2) {
    const unsigned long long* a = &pe.ctrl[pe.rank]->ack[tid];
    const unsigned long long t0 = bs_globaltimer_ns();
    while (bs_ld_acquire_sys(a) + 2 < epoch) {
      if (bs_ld_relaxed_sys(&pe.ctrl[pe.rank]->abort_flag) != 0 ||
          bs_globaltimer_ns() - t0 > st->timeout_ns) {
        atomicExch(&bad, 1);
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
  if (bad && tid == 0) st->error = RX_TIMEOUT;
}

// Rank of this lane's row among the rows of its destination seen so far in the
// CTA's tile: one shared atomic per distinct destination in the warp.
__device__ __forceinline__ int rx_rank(int* hist, int d, bool ok) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned m = __match_any_sync(0xffffffffu, ok ? d : -1 - (int)lane);
  int basev = 0;
  const int leader = __ffs((int)m) - 1;
  if (ok && (int)lane == leader) basev = atomicAdd(&hist[d], __popc(m));
  basev = __shfl_sync(0xffffffffu, basev, leader);
  return basev + __popc(m & ((1u << lane) - 1u));
}

// One elected thread per destination flushes `n` staged rows with a single TMA
// bulk store into the destination GPU's landing region.
__device__ __forceinline__ void rx_flush(const RxPeers& pe, RxState* st, int par, int d,
                                         const int4* stage_d, int n, int* cta_err) {
  const unsigned long long at = atomicAdd(&st->cursor[d], (unsigned long long)n);
  if (at + (unsigned long long)n > pe.region_rows) {
    atomicMax(&st->need_rows, at + (unsigned long long)n);
    *cta_err = RX_LANDING_OVERFLOW;
    return;
  }
  unsigned char* dst = rx_region(pe, rx_vowner(pe, d), par, rx_vlocal(pe, d), pe.rank) + at * 16ULL;
  bs_tma_store_1d(dst, stage_d, (uint32_t)n * 16u);
  bs_tma_store_commit();
}

__global__ void __launch_bounds__(RX_THREADS, RX_CTAS_PER_SM) rx_scatter_kernel(
    RxPeers pe, RxState* st, int4* __restrict__ slots, long long total, int P, int clear,
    unsigned long long clear_bits, BsCombineStatus* tst, int keymode) {
  const unsigned long long my_status = st->vote_in[0], my_fp = st->vote_in[1];
  extern __shared__ __align__(128) unsigned char rx_smem[];
  int4* stage = (int4*)rx_smem;  // [P][B] rows
  __shared__ int hist[RX_MAXP], fill[RX_MAXP], off[RX_MAXP], direct[RX_MAXP];
  __shared__ unsigned long long dbase[RX_MAXP];
  __shared__ int cta_err;
  __shared__ unsigned int is_last;
  const int PV = P * pe.regions;  // virtual partitions: (destination shard, table region)
  const int B = RX_ROWS / PV;
  const int tid = threadIdx.x;
  const unsigned long long epoch = st->epoch;
  const int par = (int)(epoch & 1ULL);
  if (tid < RX_MAXP) fill[tid] = 0;
  if (tid == 0) cta_err = 0;
  __syncthreads();
  // Rows are only sent by a healthy exchange; the scan (and the clearing) happens
  // regardless, so the table always comes out empty and reusable.  A landing overflow
  // raised by a CTA that finished earlier does not stop the others: they keep reserving
  // (rx_flush drops what does not fit), so `need_rows` ends up exact and one retry with a
  // larger arena is enough.
  const int err0 = st->error;
  const bool send = (err0 == 0 || err0 == RX_LANDING_OVERFLOW) && my_status == 0 && tst->failed == 0;
  {
    const long long ntiles = (total + RX_TILE - 1) / RX_TILE;
    int4 nxt[RX_PER];
    auto load_tile = [&](long long tile) {
#pragma unroll
      for (int j = 0; j < RX_PER; ++j) {
        const long long i = tile * RX_TILE + tid + j * RX_THREADS;
        if (tile < ntiles && i < total) {
          nxt[j] = bs_ldg_nocache(&slots[i]);
        } else {
          nxt[j] = make_int4(0, (int)0x80000000, 0, 0);  // key = BS_EMPTY_KEY
        }
      }
    };
    load_tile(blockIdx.x);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      if (tid < PV) hist[tid] = 0;
      __syncthreads();
      int4 row[RX_PER];
      int dest[RX_PER], rank[RX_PER];
#pragma unroll
      for (int j = 0; j < RX_PER; ++j) row[j] = nxt[j];
      load_tile(tile + gridDim.x);  // prefetch: consumed in the next iteration
#pragma unroll
      for (int j = 0; j < RX_PER; ++j) {
        const long long key = ((long long)(unsigned int)row[j].y << 32) | (unsigned int)row[j].x;
        const bool occupied = key != BS_EMPTY_KEY;
        const bool ok = occupied && send;
        dest[j] = -1;
        rank[j] = 0;
        if (__any_sync(0xffffffffu, occupied)) {
          const int d = ok ? rx_dest(key, P, keymode) * pe.regions + rx_table_region(key, pe.regions) : 0;
          rank[j] = rx_rank(hist, d, ok);
          if (ok) dest[j] = d;
          if (occupied) {
            if (clear) {
              int4 z = make_int4(0, (int)0x80000000, (int)(unsigned int)clear_bits,
                                 (int)(unsigned int)(clear_bits >> 32));
              slots[tile * RX_TILE + tid + j * RX_THREADS] = z;
            }
          }
        }
      }
      __syncthreads();
      // per destination: append behind what is staged; flush first when the tile's
      // rows would not fit; a tile with more rows for one destination than a whole
      // buffer is written directly
      int need = 0;
      if (tid < PV) {
        const int f = fill[tid], h = hist[tid];
        const int dir = h > B;
        need = (f + h > B) && f > 0;
        direct[tid] = dir;
        if (need) rx_flush(pe, st, par, tid, stage + (size_t)tid * B, f, &cta_err);
        if (dir) {
          const unsigned long long at = atomicAdd(&st->cursor[tid], (unsigned long long)h);
          if (at + (unsigned long long)h > pe.region_rows) {
            atomicMax(&st->need_rows, at + (unsigned long long)h);
            cta_err = RX_LANDING_OVERFLOW;
            direct[tid] = 2;  // drop
          }
          dbase[tid] = at;
        }
        off[tid] = need ? 0 : f;
        fill[tid] = dir ? (need ? 0 : f) : (need ? 0 : f) + h;
        if (need) bs_tma_store_wait_read();  // staging buffer may be overwritten now
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < RX_PER; ++j) {
        const int d = dest[j];
        if (d >= 0) {
          if (direct[d] == 1) {
            int4* dst = (int4*)rx_region(pe, rx_vowner(pe, d), par, rx_vlocal(pe, d), pe.rank);
            dst[dbase[d] + (unsigned long long)rank[j]] = row[j];
          } else if (direct[d] == 0) {
            stage[(size_t)d * B + off[d] + rank[j]] = row[j];
          }
        }
      }
      bs_fence_proxy_async();  // staged rows -> visible to the TMA (async proxy)
    }
    // drain what is still staged
    __syncthreads();
    if (tid < PV && fill[tid] > 0) rx_flush(pe, st, par, tid, stage + (size_t)tid * B, fill[tid], &cta_err);
  }
  if (tid < PV) {
    bs_tma_store_wait_all();  // my bulk stores are complete (written, not just read)
    asm volatile("fence.proxy.async;" ::: "memory");  // async-proxy writes ordered before the flag
  }
  __syncthreads();
  if (tid == 0) {
    if (cta_err) atomicMax(&st->error, cta_err);
    __threadfence_system();
    const unsigned int prev = atomicAdd(&st->done_ctas, 1u);
    is_last = prev == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (!is_last) return;
  // ===== last CTA: special row, counts + vote to the destinations, data flags =====
  __threadfence_system();
  if (tid == 0) {
    unsigned long long status = my_status ? RX_PRODUCER_FAILED : 0ULL;
    if (!status && tst->failed) status = RX_TABLE_OVERFLOW;
    if (!status && st->error) status = (unsigned long long)st->error;
    if (!status && tst->has_special) {
      const int d = rx_dest(BS_EMPTY_KEY, P, keymode) * pe.regions + rx_table_region(BS_EMPTY_KEY, pe.regions);
      const unsigned long long at = atomicAdd(&st->cursor[d], 1ULL);
      if (at + 1 > pe.region_rows) {
        atomicMax(&st->need_rows, at + 1);
        status = RX_LANDING_OVERFLOW;
      } else {
        int4* dst = (int4*)rx_region(pe, rx_vowner(pe, d), par, rx_vlocal(pe, d), pe.rank);
        dst[at] = make_int4(0, (int)0x80000000, (int)(unsigned int)tst->special_val,
                            (int)(unsigned int)(tst->special_val >> 32));
      }
    }
    st->error = (int)status;  // republished through the vote: every rank sees it
    if (clear) {              // the table is empty again: ready for its next combine
      tst->failed = 0;
      tst->has_special = 0;
      tst->special_val = clear_bits;
      tst->out_count = 0;
    }
  }
  __syncthreads();
  long long sent = 0;
  if (tid < PV) {
    unsigned long long c = st->cursor[tid];
    if (c > pe.region_rows) c = pe.region_rows;
    pe.ctrl[rx_vowner(pe, tid)]->cnt[par][pe.rank][tid] = (long long)c;
    sent = (long long)c;
  }
  for (int o = 16; o > 0; o >>= 1) sent += __shfl_down_sync(0xffffffffu, sent, o);
  if ((tid & 31) == 0 && tid < RX_MAXP && sent) atomicAdd((unsigned long long*)&st->sent_rows, (unsigned long long)sent);
  if (tid < pe.world) {
    unsigned long long* v = pe.ctrl[tid]->vote[par][pe.rank];
    v[0] = (unsigned long long)st->error;
    v[1] = my_fp;
    v[2] = st->need_rows;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < pe.world) bs_st_release_sys(&pe.ctrl[tid]->data_flag[pe.rank], epoch);
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(64) rx_wait_kernel(RxPeers pe, RxState* st, int P) {
  const unsigned long long my_fp = st->vote_in[1];
  __shared__ int verdict;
  __shared__ unsigned long long need;
  const int tid = threadIdx.x;
  const unsigned long long epoch = st->epoch;
  const int par = (int)(epoch & 1ULL);
  RxCtrl* mine = pe.ctrl[pe.rank];
  if (tid == 0) {
    verdict = 0;
    need = 0;
  }
  __syncthreads();
  if (tid < pe.world) {
    const unsigned long long* f = &mine->data_flag[tid];
    const unsigned long long t0 = bs_globaltimer_ns();
    bool ok = true;
    while (bs_ld_acquire_sys(f) < epoch) {
      if (bs_ld_relaxed_sys(&mine->abort_flag) != 0 || bs_globaltimer_ns() - t0 > st->timeout_ns) {
        ok = false;
        break;
      }
      __nanosleep(100);
    }
    if (!ok) {
      atomicMax(&verdict, RX_TIMEOUT);
    } else {
      const unsigned long long* v = mine->vote[par][tid];
      if (v[0] != 0) atomicMax(&verdict, (int)v[0]);
      else if (v[1] != my_fp) atomicMax(&verdict, RX_FINGERPRINT);
      atomicMax(&need, v[2]);
    }
  }
  __syncthreads();
  if (tid == 0) {
    if (verdict == RX_TIMEOUT)
      for (int p = 0; p < pe.world; ++p) bs_st_release_sys(&pe.ctrl[p]->abort_flag, 1ULL);
    st->error = verdict;  // identical on every rank (each one saw every vote)
    if (need > st->need_rows) st->need_rows = need;
  }
  for (int x = tid; x < pe.shards_per_rank * pe.world; x += blockDim.x) {
    const int jv = x / pe.world, s = x - jv * pe.world;       // local virtual index, source
    const int d = pe.rank + (jv / pe.regions) * pe.world;    // destination shard
    const int v = d * pe.regions + (jv % pe.regions);
    st->recv[jv][s] = (d < P && verdict == 0) ? mine->cnt[par][s][v] : 0;
  }
}

// ---------------------------------------------------------------------------
// Combine on arrival: the rows every source wrote for local shard j go into the
// shard's final table.  Sources are pre-combined, so there are no heavy hitters
// here (a key arrives at most once per source): no shared-memory level, four
// probes per thread in flight, L2 atomics.
#define RC_THREADS 512
#define RC_PER 4
template <typename V, int OP>
__global__ void __launch_bounds__(RC_THREADS) rx_combine_kernel(RxPeers pe, const RxState* __restrict__ st,
                                                                int j, BsTable tab,
                                                                BsCombineStatus* __restrict__ dst_st) {
  if (st->error != 0) return;
  __shared__ long long pre[RX_MAXW + 1];
  const int par = (int)(st->epoch & 1ULL);
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int s = 0; s < pe.world; ++s) {
      pre[s] = run;
      run += st->recv[j][s];
    }
    pre[pe.world] = run;
  }
  __syncthreads();
  const long long total = pre[pe.world];
  const long long chunk = (long long)RC_THREADS * RC_PER;
  for (long long c0 = (long long)blockIdx.x * chunk; c0 < total; c0 += (long long)gridDim.x * chunk) {
    long long key[RC_PER];
    V val[RC_PER];
    uint32_t h2[RC_PER];
    long long cur[RC_PER] = {0, 0, 0, 0};
    unsigned live = 0;
#pragma unroll
    for (int r = 0; r < RC_PER; ++r) {
      const long long i = c0 + (long long)r * RC_THREADS + threadIdx.x;
      if (i < total) {
        int s = 0;
        while (i >= pre[s + 1]) ++s;
        const int4* src = (const int4*)rx_region(pe, pe.rank, par, j, s);
        const int4 raw = bs_ldg_nocache(&src[i - pre[s]]);
        key[r] = ((long long)(unsigned int)raw.y << 32) | (unsigned int)raw.x;
        unsigned long long bits = ((unsigned long long)(unsigned int)raw.w << 32) | (unsigned int)raw.z;
        memcpy(&val[r], &bits, sizeof(V));
        if (key[r] == BS_EMPTY_KEY) {
          special_insert<V, OP>(val[r], dst_st);
        } else {
          live |= 1u << r;
          h2[r] = bs_hash_main(bs_mix_u64((uint64_t)key[r]));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RC_PER; ++r)
      if ((live >> r) & 1u) cur[r] = __ldcg(&tab.main[bs_main_slot(tab, h2[r])].key);
    long long prev[RC_PER];
    unsigned claim = 0;
#pragma unroll
    for (int r = 0; r < RC_PER; ++r) {
      prev[r] = cur[r];
      if (((live >> r) & 1u) && cur[r] == BS_EMPTY_KEY) {
        claim |= 1u << r;
        prev[r] = (long long)atomicCAS((unsigned long long*)&tab.main[bs_main_slot(tab, h2[r])].key,
                                       (unsigned long long)BS_EMPTY_KEY, (unsigned long long)key[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < RC_PER; ++r) {
      if (!((live >> r) & 1u)) continue;
      const bool won = ((claim >> r) & 1u) && prev[r] == BS_EMPTY_KEY;
      if (won || prev[r] == key[r])
        Combine<V, OP>::apply(slot_val<V>(&tab.main[bs_main_slot(tab, h2[r])]), val[r]);
      else
        main_resolve<V, OP>(tab, key[r], val[r], h2[r], prev[r], dst_st, nullptr);
    }
  }
}

// ---------------------------------------------------------------------------
struct RxEndArgs {
  BsCombineStatus* st[RX_MAXP];  // destination tables' status blocks (null: none)
  long long caps[RX_MAXP];       // capacities of the output buffers
  unsigned long long clear_bits; // identity bit pattern: the tables were compacted-and-cleared
};

// Everything a rank can only find out after combining (destination table or output
// buffer too small) is rank-local knowledge; a rank that retried on its own would
// pair its repeated exchange with the peers' *next* one.  So the exchange ends with
// a second, tiny agreement: every rank stores its outcome into every peer's `fin`
// slot and reads the others' — the report carries the same verdict on every rank
// and a lost exchange is retried by all of them together.
__global__ void __launch_bounds__(64) rx_end_kernel(RxPeers pe, RxState* st, int nlocal, RxEndArgs a,
                                                    RxReport* ring, int nslots, int ack) {
  __shared__ int verdict;
  const int tid = threadIdx.x;
  const unsigned long long epoch = st->epoch;
  RxReport* rep = ring + (epoch % (unsigned long long)nslots);
  // the landing buffers of this epoch have been read: sources may reuse them
  __threadfence_system();
  if (ack && tid < pe.world) bs_st_release_sys(&pe.ctrl[tid]->ack[pe.rank], epoch);
  if (tid == 0) {
    long long err = st->error;
    for (int j = 0; j < nlocal; ++j) {
      long long rows = 0;
      for (int r = 0; r < pe.regions; ++r)
        for (int s = 0; s < pe.world; ++s) rows += st->recv[j * pe.regions + r][s];
      rep->recv_rows[j] = rows;
      BsCombineStatus* d = a.st[j];
      rep->out_rows[j] = d ? (long long)d->out_count : 0;
      rep->dest_failed[j] = d ? (long long)d->failed : 0;
      if (d && err == 0 && (d->failed != 0 || (long long)d->out_count > a.caps[j])) err = RX_DEST_OVERFLOW;
      if (d) {  // the destination table is empty again: ready for its next exchange
        d->failed = 0;
        d->has_special = 0;
        d->special_val = a.clear_bits;
        d->out_count = 0;
      }
    }
    verdict = (int)err;
  }
  __syncthreads();
  const int mine = verdict;
  __syncthreads();
  if (ack) {  // (the probe half of a first-run exchange ends without acknowledging: no agreement yet)
    if (tid < pe.world)
      bs_st_release_sys(&pe.ctrl[tid]->fin[pe.rank], (epoch << 8) | (unsigned long long)(mine & 0xff));
    if (tid < pe.world && mine != RX_TIMEOUT) {
      const unsigned long long* f = &pe.ctrl[pe.rank]->fin[tid];
      const unsigned long long t0 = bs_globaltimer_ns();
      unsigned long long v;
      bool ok = true;
      while (((v = bs_ld_acquire_sys(f)) >> 8) < epoch) {
        if (bs_ld_relaxed_sys(&pe.ctrl[pe.rank]->abort_flag) != 0 || bs_globaltimer_ns() - t0 > st->timeout_ns) {
          ok = false;
          break;
        }
        __nanosleep(100);
      }
      if (!ok) atomicMax(&verdict, RX_TIMEOUT);
      else if ((v >> 8) == epoch && (v & 0xffULL) != 0) atomicMax(&verdict, (int)(v & 0xffULL));
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (ack && verdict == RX_TIMEOUT && mine != RX_TIMEOUT)
      for (int p = 0; p < pe.world; ++p) bs_st_release_sys(&pe.ctrl[p]->abort_flag, 1ULL);
    rep->error = (long long)verdict;
    rep->need_rows = (long long)st->need_rows;
    rep->sent_rows = st->sent_rows;
    __threadfence_system();
    rep->epoch = epoch;  // last: the host polls this field
    __threadfence_system();
  }
}

// Publishes the outcome of a purely local combine->compact (no exchange) the
// same way: the host learns the row count from pinned memory when it needs it.
__global__ void table_publish_kernel(const BsCombineStatus* st, long long out_cap, RxState* xs,
                                     RxReport* ring, int nslots) {
  const unsigned long long ticket = ++xs->publish_seq;
  RxReport* rep = ring + (ticket % (unsigned long long)nslots);
  long long err = 0;
  if (st->failed != 0) err = RX_TABLE_OVERFLOW;
  else if ((long long)st->out_count > out_cap) err = RX_DEST_OVERFLOW;
  rep->error = err;
  rep->need_rows = 0;
  rep->sent_rows = 0;
  rep->recv_rows[0] = 0;
  rep->out_rows[0] = (long long)st->out_count;
  rep->dest_failed[0] = (long long)st->failed;
  __threadfence_system();
  rep->epoch = ticket;
  __threadfence_system();
}

// ---------------------------------------------------------------------------
// host API
BS_EXPORT int bs_rx_sizes(int* ctrl_bytes, int* state_bytes, int* report_bytes, int* peers_bytes) {
  *ctrl_bytes = (int)sizeof(RxCtrl);
  *state_bytes = (int)sizeof(RxState);
  *report_bytes = (int)sizeof(RxReport);
  *peers_bytes = (int)sizeof(RxPeers);
  return 0;
}
BS_EXPORT int bs_rx_limits(int* maxw, int* maxp) {
  *maxw = RX_MAXW;
  *maxp = RX_MAXP;
  return 0;
}
BS_EXPORT int bs_rx_report_offsets(long long* recv_rows, long long* out_rows, long long* dest_failed) {
  RxReport* r = nullptr;
  *recv_rows = (long long)(uintptr_t)&r->recv_rows[0];
  *out_rows = (long long)(uintptr_t)&r->out_rows[0];
  *dest_failed = (long long)(uintptr_t)&r->dest_failed[0];
  return 0;
}

// (Re)initialise the private state: epoch 0, timeout.
BS_EXPORT int bs_rx_state_init(void* state_d, unsigned long long timeout_ns, cudaStream_t stream) {
  RxState z;
  memset(&z, 0, sizeof(z));
  z.timeout_ns = timeout_ns;
  cudaError_t e = cudaMemcpyAsync(state_d, &z, sizeof(z), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaStreamSynchronize(stream);
}

// Pinned, device-mapped host memory for reports.
BS_EXPORT int bs_host_alloc(long long bytes, void** ptr) {
  cudaError_t e = cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) return (int)e;
  memset(*ptr, 0, (size_t)bytes);
  return 0;
}
BS_EXPORT int bs_host_free(void* ptr) { return (int)cudaFreeHost(ptr); }

// Producer half: begin + scatter of the local table.
__global__ void rx_set_vote_kernel(RxState* st, unsigned long long status, unsigned long long fp) {
  st->vote_in[0] = status;
  st->vote_in[1] = fp;
}
// The vote of the next exchange (kept out of the captured graph: it changes every step).
BS_EXPORT int bs_rx_set_vote(void* state_d, unsigned long long status, unsigned long long fp,
                             cudaStream_t stream) {
  rx_set_vote_kernel<<<1, 1, 0, stream>>>((RxState*)state_d, status, fp);
  BS_LAUNCH_CHECK();
  return 0;
}

BS_EXPORT int bs_rx_scatter(const void* peers_h, void* state_d, void* slots, long long total, int nparts,
                            int clear, unsigned long long clear_bits, void* table_status, int keymode,
                            cudaStream_t stream) {
  RxPeers pe = *(const RxPeers*)peers_h;
  if (pe.world > RX_MAXW || nparts < 1 || pe.regions < 1 || (pe.regions & (pe.regions - 1)) ||
      nparts * pe.regions > RX_MAXP)
    return -1;
  RxState* st = (RxState*)state_d;
  rx_begin_kernel<<<1, 64, 0, stream>>>(pe, st, nparts);
  BS_LAUNCH_CHECK();
  const int smem = RX_ROWS * 16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(rx_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int sms = bs_num_sms();
  long long ntiles = (total + RX_TILE - 1) / RX_TILE;
  int blocks = (int)(ntiles < (long long)sms * RX_CTAS_PER_SM ? ntiles : (long long)sms * RX_CTAS_PER_SM);
  if (blocks < 1) blocks = 1;
  rx_scatter_kernel<<<blocks, RX_THREADS, smem, stream>>>(pe, st, (int4*)slots, total, nparts, clear, clear_bits,
                                                          (BsCombineStatus*)table_status, keymode);
  BS_LAUNCH_CHECK();
  return 0;
}

BS_EXPORT int bs_rx_wait(const void* peers_h, void* state_d, int nparts, cudaStream_t stream) {
  RxPeers pe = *(const RxPeers*)peers_h;
  rx_wait_kernel<<<1, 64, 0, stream>>>(pe, (RxState*)state_d, nparts);
  BS_LAUNCH_CHECK();
  return 0;
}

template <typename V, int OP>
static int rx_launch_combine(const RxPeers& pe, const RxState* st, int j, void* slots, long long cap,
                             void* dst_status, long long max_rows, cudaStream_t stream) {
  BsTable tab;
  memset(&tab, 0, sizeof(tab));
  tab.main = (BsSlot*)slots;
  tab.main_cap = (unsigned long long)cap;
  const long long chunk = (long long)RC_THREADS * RC_PER;
  long long want = (max_rows + chunk - 1) / chunk;
  const long long cap_blocks = (long long)bs_num_sms() * 4;
  int blocks = (int)(want < cap_blocks ? want : cap_blocks);
  if (blocks < 1) blocks = 1;
  rx_combine_kernel<V, OP><<<blocks, RC_THREADS, 0, stream>>>(pe, st, j, tab, (BsCombineStatus*)dst_status);
  BS_LAUNCH_CHECK();
  return 0;
}

#define RX_DISPATCH_VOP(FN, ...)                                                       \
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

// Combine the landed rows of local shard j into the (clean) table `slots[cap]`.
// max_rows bounds the grid (worst case: every region full).
BS_EXPORT int bs_rx_combine(const void* peers_h, const void* state_d, int j, int vtype, int op, void* slots,
                            long long cap, void* dst_status, long long max_rows, cudaStream_t stream) {
  RxPeers pe = *(const RxPeers*)peers_h;
  if (cap <= 0 || cap >= (1LL << 32)) return -2;
  RX_DISPATCH_VOP(rx_launch_combine, pe, (const RxState*)state_d, j, slots, cap, dst_status, max_rows, stream)
}

// dst_status / out_caps: host arrays [nlocal] (device pointers to the destination
// tables' status blocks; capacities of the output buffers).
BS_EXPORT int bs_rx_end(const void* peers_h, void* state_d, int nlocal, const void* const* dst_status,
                        const long long* out_caps, unsigned long long clear_bits, void* report_ring,
                        int nslots, int ack, cudaStream_t stream) {
  RxPeers pe = *(const RxPeers*)peers_h;
  if (nlocal > RX_MAXP) return -1;
  RxEndArgs a;
  memset(&a, 0, sizeof(a));
  for (int j = 0; j < nlocal; ++j) {
    a.st[j] = (BsCombineStatus*)dst_status[j];
    a.caps[j] = out_caps[j];
  }
  a.clear_bits = clear_bits;
  rx_end_kernel<<<1, 64, 0, stream>>>(pe, (RxState*)state_d, nlocal, a, (RxReport*)report_ring, nslots, ack);
  BS_LAUNCH_CHECK();
  return 0;
}

BS_EXPORT int bs_table_publish(const void* status, long long out_cap, void* state_d, void* report_ring,
                               int nslots, cudaStream_t stream) {
  table_publish_kernel<<<1, 1, 0, stream>>>((const BsCombineStatus*)status, out_cap, (RxState*)state_d,
                                            (RxReport*)report_ring, nslots);
  BS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// CUDA graphs: a fixed launch sequence (scatter -> wait -> combine -> compact ->
// end) is captured once per shape and re-launched as ONE graph per step.
static long long g_graph_nodes[1];
BS_EXPORT int bs_graph_begin(cudaStream_t stream) {
  return (int)cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal);
}
BS_EXPORT int bs_graph_end(cudaStream_t stream, void** exec_out, long long* nodes_out) {
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(stream, &g);
  if (e != cudaSuccess || g == nullptr) {
    cudaGetLastError();
    return e != cudaSuccess ? (int)e : -1;
  }
  size_t n = 0;
  cudaGraphGetNodes(g, nullptr, &n);
  cudaGraphExec_t ex = nullptr;
  e = cudaGraphInstantiate(&ex, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return (int)e;
  *exec_out = (void*)ex;
  *nodes_out = (long long)n;
  (void)g_graph_nodes;
  return 0;
}
BS_EXPORT int bs_graph_launch(void* exec, long long nodes, cudaStream_t stream) {
  cudaError_t e = cudaGraphLaunch((cudaGraphExec_t)exec, stream);
  if (e == cudaSuccess) __sync_fetch_and_add(&bs_launch_counter, nodes);
  return (int)e;
}
BS_EXPORT int bs_graph_destroy(void* exec) { return (int)cudaGraphExecDestroy((cudaGraphExec_t)exec); }
