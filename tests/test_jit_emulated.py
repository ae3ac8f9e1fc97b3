"""The generated compacting kernels (``jit.generate_compact``), executed on the CPU.

``bs_fused_count`` / ``bs_fused_write`` use warp ballots, a shared-memory scan and
block barriers, so defining the CUDA keywords away is not enough to run them on the
host.  This test compiles the *unmodified generated source* with g++ against the SIMT
emulator of ``tests/emu/simt.h``: every CUDA thread of a CTA is a ``ucontext`` fiber,
``__syncthreads`` and ``__ballot_sync`` are cooperative barriers over the fibers of the
block / warp, ``__shared__`` variables are statics (CTAs run one after the other).  The kept rows
must come out exactly as the host evaluation of the same chain produces them — row
order included — for chains with pass-through columns, type changes, bool outputs
and row counts that are not a multiple of the CTA chunk.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200.ops import jit

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")

EMULATOR = '#include "simt.h"\n'   # tests/emu/simt.h: fibers, warp collectives, shared memory

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
        "  bs_emu_launch(nblocks, 256, 0, [=] { bs_fused_count(n, %s, counts); });\n}\n"
        'extern "C" void run_write(long long nblocks, long long n, %s, const long long* offs, %s) {\n'
        "  bs_emu_launch(nblocks, 256, 0, [=] { bs_fused_write(n, %s, offs, %s); });\n}\n"
        % (ins, in_args, ins, outs, in_args, out_args))
    cpp, so = tmp_path / (tag + ".cpp"), tmp_path / (tag + ".so")
    cpp.write_text(EMULATOR + src + wrappers)
    subprocess.run(["g++", "-O1", "-fwrapv", "-std=c++17", "-shared", "-fPIC", "-w", "-I", EMU_DIR, "-o", str(so),
                    str(cpp)], check=True)
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
