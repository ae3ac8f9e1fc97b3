"""The generated compacting kernels (``jit.generate_compact``), executed on the CPU.

``bs_fused_count`` / ``bs_fused_write`` use warp ballots, a shared-memory scan and
block barriers, so defining the CUDA keywords away is not enough to run them on the
host.  This test compiles the *unmodified generated source* with g++ against a small
SIMT emulator: every CUDA thread of a CTA is a ``ucontext`` fiber, ``__syncthreads``
and ``__ballot_sync`` are cooperative barriers over the fibers of the block / warp,
``__shared__`` variables are statics (CTAs run one after the other).  The kept rows
must come out exactly as the host evaluation of the same chain produces them — row
order included — for chains with pass-through columns, type changes, bool outputs
and row counts that are not a multiple of the CTA chunk.
"""
import ctypes
import shutil
import subprocess

import numpy as np
import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200.ops import jit

EMULATOR = r"""
#include <ucontext.h>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <vector>
using std::fabs; using std::floor; using std::ceil; using std::sqrt; using std::exp; using std::log;
using std::pow; using std::fmod;
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(x)
#define __shared__ static
struct bs_dim3 { long long x; };
static bs_dim3 blockIdx{0}, blockDim{256}, threadIdx{0}, gridDim{1};

static ucontext_t bs_sched;
static std::vector<ucontext_t> bs_fib;
static std::vector<char> bs_done;
static int bs_cur;
static std::function<void()> bs_kernel;
static inline void bs_yield() { swapcontext(&bs_fib[bs_cur], &bs_sched); }

static int bs_bar_gen, bs_bar_count;
static inline void __syncthreads() {
  const int gen = bs_bar_gen;
  if (++bs_bar_count == 256) { bs_bar_count = 0; ++bs_bar_gen; }
  while (bs_bar_gen == gen) bs_yield();
}
struct bs_warp { unsigned bits, result; int count, gen; };
static bs_warp bs_warps[8];
static inline unsigned __ballot_sync(unsigned, bool p) {
  bs_warp& w = bs_warps[bs_cur >> 5];
  const int gen = w.gen;
  if (p) w.bits |= 1u << (bs_cur & 31);
  if (++w.count == 32) { w.result = w.bits; w.bits = 0; w.count = 0; ++w.gen; }
  while (w.gen == gen) bs_yield();
  return w.result;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }

static void bs_trampoline() {
  bs_kernel();
  bs_done[bs_cur] = 1;
  swapcontext(&bs_fib[bs_cur], &bs_sched);
}
static void bs_launch(long long nblocks, std::function<void()> k) {
  bs_kernel = k;
  gridDim.x = nblocks;
  static std::vector<std::vector<char>> stacks(256, std::vector<char>(1 << 16));
  for (long long b = 0; b < nblocks; ++b) {
    blockIdx.x = b;
    bs_bar_gen = bs_bar_count = 0;
    for (auto& w : bs_warps) w = bs_warp{0, 0, 0, 0};
    bs_fib.assign(256, ucontext_t());
    bs_done.assign(256, 0);
    for (int t = 0; t < 256; ++t) {
      getcontext(&bs_fib[t]);
      bs_fib[t].uc_stack.ss_sp = stacks[t].data();
      bs_fib[t].uc_stack.ss_size = stacks[t].size();
      bs_fib[t].uc_link = &bs_sched;
      makecontext(&bs_fib[t], bs_trampoline, 0);
    }
    int ndone = 0;
    while (ndone < 256)
      for (int t = 0; t < 256; ++t) {
        if (bs_done[t]) continue;
        bs_cur = t;
        threadIdx.x = t;
        swapcontext(&bs_sched, &bs_fib[t]);
        if (bs_done[t]) ++ndone;
      }
  }
}
"""

NP = {"int64": np.int64, "int32": np.int32, "float32": np.float32, "float64": np.float64, "bool": np.bool_,
      "int16": np.int16, "int8": np.int8, "uint8": np.uint8}


def build(tmp_path, tag, in_types, stages):
    src, out_types = jit.generate_compact(in_types, stages)
    ins = ", ".join("const %s* in%d" % (jit.CTYPE[t.name], j) for j, t in enumerate(in_types))
    in_args = ", ".join("in%d" % j for j in range(len(in_types)))
    outs = ", ".join("%s* out%d" % (jit.CTYPE[t.name], j) for j, t in enumerate(out_types))
    out_args = ", ".join("out%d" % j for j in range(len(out_types)))
    wrappers = (
        'extern "C" void run_count(long long nblocks, long long n, %s, int* counts) {\n'
        "  bs_launch(nblocks, [=] { bs_fused_count(n, %s, counts); });\n}\n"
        'extern "C" void run_write(long long nblocks, long long n, %s, const long long* offs, %s) {\n'
        "  bs_launch(nblocks, [=] { bs_fused_write(n, %s, offs, %s); });\n}\n"
        % (ins, in_args, ins, outs, in_args, out_args))
    cpp, so = tmp_path / (tag + ".cpp"), tmp_path / (tag + ".so")
    cpp.write_text(EMULATOR + src + wrappers)
    subprocess.run(["g++", "-O1", "-fwrapv", "-std=c++17", "-shared", "-fPIC", "-w", "-o", str(so), str(cpp)],
                   check=True)
    return ctypes.CDLL(str(so)), out_types


def run_emulated(lib, out_types, cols):
    n = len(cols[0])
    nblocks = (n + jit.COMPACT_CHUNK - 1) // jit.COMPACT_CHUNK
    arrs = [np.ascontiguousarray(c.numpy()) for c in cols]
    ptrs = [ctypes.c_void_p(a.ctypes.data) for a in arrs]
    counts = np.zeros(nblocks, dtype=np.int32)
    lib.run_count(ctypes.c_longlong(nblocks), ctypes.c_longlong(n), *ptrs, ctypes.c_void_p(counts.ctypes.data))
    incl = np.cumsum(counts.astype(np.int64))
    total = int(incl[-1])
    offs = np.ascontiguousarray(incl - counts)
    outs = [np.empty(total, dtype=NP[t.name]) for t in out_types]
    lib.run_write(ctypes.c_longlong(nblocks), ctypes.c_longlong(n), *ptrs, ctypes.c_void_p(offs.ctypes.data),
                  *[ctypes.c_void_p(o.ctypes.data) for o in outs])
    return [torch.from_numpy(o) for o in outs], counts


CHAINS = {
    # Map -> Filter -> Map, a pass-through column, int64 % (the benchmark's shape)
    "map_filter_map": ([bs.int64, bs.float32],
                       [jit.Stage(exprs=["c0 % 1009", "c1 * 2.0f + 1.0f", "c0"], out=[bs.int64, bs.float32, bs.int64]),
                        jit.Stage(pred="c2 % 3 != 0 && c1 > 0.0f"),
                        jit.Stage(exprs=["c0", "c1 + c2"], out=[bs.int64, bs.float32])]),
    # a bare Filter: every output is an input column
    "filter_only": ([bs.int32, bs.int64], [jit.Stage(pred="(c0 & 1) == 0")]),
    # two predicates, a dtype change and a bool output
    "two_filters": ([bs.int64, bs.int32, bs.float64],
                    [jit.Stage(pred="c0 > -50"),
                     jit.Stage(exprs=["c1 + 1", "c2 * 0.5", "c0 < 0"], out=[bs.int64, bs.float32, bs.bool_]),
                     jit.Stage(pred="c0 != 7")]),
    # nothing survives / everything survives
    "keep_none": ([bs.int64], [jit.Stage(pred="c0 > 1000000")]),
    "keep_all": ([bs.int64], [jit.Stage(exprs=["c0 * 3"], out=[bs.int64]), jit.Stage(pred="c0 == c0")]),
}


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
@pytest.mark.parametrize("name", sorted(CHAINS))
@pytest.mark.parametrize("n", [1, 255, 2048, 5000])
def test_compacting_kernels_emit_exactly_the_kept_rows_in_order(name, n, tmp_path):
    in_types, stages = CHAINS[name]
    g = torch.Generator().manual_seed(n)
    cols = []
    for t in in_types:
        if t.torch_dtype.is_floating_point:
            cols.append((torch.randn(n, generator=g, dtype=torch.float64) * 30).to(t.torch_dtype))
        else:
            cols.append(torch.randint(-100, 100, (n,), generator=g).to(t.torch_dtype))
    lib, out_types = build(tmp_path, name, in_types, stages)
    got, counts = run_emulated(lib, out_types, cols)
    want, mask = jit.eval_host(cols, in_types, stages)
    want = [w[mask] for w in want]
    assert int(counts.sum()) == int(mask.sum())
    # per-CTA counts are the kept rows of each 2048-row chunk
    assert counts.tolist() == [int(mask[i:i + jit.COMPACT_CHUNK].sum()) for i in range(0, n, jit.COMPACT_CHUNK)]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.dtype == b.dtype and a.shape == b.shape
        if b.is_floating_point():
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)
        else:
            assert torch.equal(a, b)
