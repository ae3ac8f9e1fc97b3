// Open-addressing <key, value> tables shared by the combine kernels
// (hash_combine.cu) and the Reduce exchange (reduce_exchange.cu): typed atomic
// combiners, the 16-byte slot layout, the two-level (L2-resident warm + HBM main)
// table descriptor and the device-side insert path.
//
// Replaces the reference's combiningFrame (exec/combiner.go:62-223): probing and
// combining are lock-free atomics on HBM/L2 instead of a per-row reflective call.
#pragma once
#include "common.cuh"
#include <type_traits>
#include <string.h>

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

// ---- HBM table: array-of-structs, one 16-byte slot per key ------------------
// A probe touches a single 32-byte sector (key + value together), so a miss in
// the shared-memory table costs one DRAM line instead of two.
struct __align__(16) BsSlot {
  long long key;
  unsigned long long val;  // the value's bits live in the low sizeof(V) bytes
};

template <typename V>
__device__ __forceinline__ V* slot_val(BsSlot* s) { return (V*)&s->val; }

template <typename V, int OP>
__global__ void __launch_bounds__(256) table_init_kernel(BsSlot* __restrict__ tab, long long cap) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  V id = Combine<V, OP>::identity();
  unsigned long long bits = 0;
  memcpy(&bits, &id, sizeof(V));
  for (; i < cap; i += stride) {
    BsSlot s;
    s.key = BS_EMPTY_KEY;
    s.val = bits;
    tab[i] = s;  // one 16-byte store
  }
}

// ---- two-level HBM table ----------------------------------------------------
// slots = [ warm (warm_mask+1 slots, a power of two) | main (main_cap slots, any size) ].
//  * warm: direct-mapped, small enough to live in the 126 MB L2 (accessed with
//    evict_last hints).  A key that finds its warm slot free claims it for good;
//    a key whose warm slot belongs to another key always goes to `main`.  Hence
//    every key lives in exactly one level and the union of both levels is the
//    result — compaction and the exchange just scan the whole array.
//  * main: open addressing with linear probing (random DRAM accesses), reached
//    only by the cold tail of the key distribution.
struct BsTable {
  BsSlot* warm;
  unsigned long long warm_mask;  // 0 when there is no warm level
  BsSlot* main;
  unsigned long long main_cap;   // any size < 2^32: slots are picked by fast range reduction
  int has_warm;
  // Optional: count newly inserted keys per destination shard
  // (murmur3(key) % nparts, the exchange's partition function) while combining,
  // so the table exchange needs no counting pass over the slot array.
  int nparts;                       // 0: disabled
  unsigned long long* dest_counts;  // [nparts], accumulated across launches
};

// A key was inserted for the first time: bump its destination's counter in the
// CTA's shared histogram (flushed once per CTA).
__device__ __forceinline__ void count_claim(const BsTable& t, int* dhist, long long key) {
  if (t.nparts > 0) atomicAdd(&dhist[bs_murmur3_u64((uint64_t)key, 0u) % (uint32_t)t.nparts], 1);
}

__device__ __forceinline__ uint32_t bs_hash_main(uint32_t h) { return bs_fmix32(h + 0x9e3779b9u); }
// First probe slot in the main level: (h2 * cap) >> 32 — uniform over [0, cap) for
// any capacity, so tables are sized to the cardinality instead of the next power
// of two (init, compaction and the exchange all scan the whole slot array).
__device__ __forceinline__ unsigned long long bs_main_slot(const BsTable& t, uint32_t h2) {
  return ((unsigned long long)h2 * t.main_cap) >> 32;
}

template <typename V, int OP>
__device__ __forceinline__ void main_resolve(const BsTable& t, long long key, V val, uint32_t h2,
                                             long long cur, BsCombineStatus* st, int* dhist) {
  unsigned long long slot = bs_main_slot(t, h2);
  // Linear probing; bounded so a full table cannot hang the GPU.
  for (int probe = 0; probe < 4096; ++probe) {
    if (cur == key) {
      Combine<V, OP>::apply(slot_val<V>(&t.main[slot]), val);
      return;
    }
    if (cur == BS_EMPTY_KEY) {
      long long prev = (long long)atomicCAS((unsigned long long*)&t.main[slot].key,
                                            (unsigned long long)BS_EMPTY_KEY, (unsigned long long)key);
      if (prev == BS_EMPTY_KEY || prev == key) {
        if (prev == BS_EMPTY_KEY) count_claim(t, dhist, key);
        Combine<V, OP>::apply(slot_val<V>(&t.main[slot]), val);
        return;
      }
    }
    slot = slot + 1 == t.main_cap ? 0 : slot + 1;
    cur = __ldcg(&t.main[slot].key);
  }
  atomicAdd(&st->failed, 1ULL);
}

// Try the warm level given the already-loaded key of the warm slot; returns
// true when the row was absorbed.
template <typename V, int OP>
__device__ __forceinline__ bool warm_resolve(const BsTable& t, long long key, V val, uint32_t h,
                                             long long cur, int* dhist) {
  BsSlot* s = &t.warm[(unsigned long long)h & t.warm_mask];
  if (cur != key) {
    if (cur != BS_EMPTY_KEY) return false;
    long long prev = (long long)atomicCAS((unsigned long long*)&s->key,
                                          (unsigned long long)BS_EMPTY_KEY, (unsigned long long)key);
    if (prev != BS_EMPTY_KEY && prev != key) return false;
    if (prev == BS_EMPTY_KEY) count_claim(t, dhist, key);
  }
  Combine<V, OP>::apply(slot_val<V>(s), val);
  return true;
}

template <typename V, int OP>
__device__ __forceinline__ void table_insert(const BsTable& t, long long key, V val, uint32_t h,
                                             BsCombineStatus* st, int* dhist) {
  if (t.has_warm) {
    long long cur = __ldcg(&t.warm[(unsigned long long)h & t.warm_mask].key);
    if (warm_resolve<V, OP>(t, key, val, h, cur, dhist)) return;
  }
  uint32_t h2 = bs_hash_main(h);
  main_resolve<V, OP>(t, key, val, h2, __ldcg(&t.main[bs_main_slot(t, h2)].key), st, dhist);
}

template <typename V, int OP>
__device__ __forceinline__ void special_insert(V val, BsCombineStatus* st) {
  // The reserved EMPTY key itself: combined in a dedicated status slot
  // (special_val starts at the identity, see bs_table_init).
  atomicExch(&st->has_special, 1ULL);
  Combine<V, OP>::apply((V*)&st->special_val, val);
}

