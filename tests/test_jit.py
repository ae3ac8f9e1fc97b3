"""Expression stages: typechecks, host evaluation with C semantics, fusion of
pipelined stages, NVRTC compilation to an sm_100a cubin (no GPU needed to
compile) and, on a GPU, the fused kernel against the host evaluation."""
import os

import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200 import slicetest, typecheck
from bigslice_b200.ops import jit
from bigslice_b200.slices import _ExprReader


def chain(src):
    s = bs.MapExpr(src, ["c0 % 7", "c1 * 2 + c0"], [bs.int64, bs.float32])
    s = bs.FilterExpr(s, "c0 != 3 && c1 > 0")
    return bs.MapExpr(s, ["c0 + 100", "(int)c1 / 3"], [bs.int64, bs.int32])


def reference(k, v):
    a = torch.fmod(k, 7)
    b = v * 2 + k.float()
    m = (a != 3) & (b > 0)
    return (a + 100)[m], torch.div(b.to(torch.int32), 3, rounding_mode="trunc")[m]


def test_host_eval_c_semantics_and_results():
    k = torch.arange(-50, 50)
    v = torch.linspace(-3, 3, 100)
    res = slicetest.run(chain(bs.Const(3, k, v)))
    wa, wb = reference(k, v)
    rows = res.rows()
    assert [r[0] for r in rows] == wa.tolist() and [r[1] for r in rows] == wb.tolist()


def test_stages_fuse_into_one_reader():
    src = bs.Const(1, torch.arange(10), torch.ones(10))
    s = chain(src)
    from bigslice_b200.exec.compile import ExecInvocation, compile_slice
    from bigslice_b200.func import Invocation
    tasks = compile_slice(ExecInvocation(Invocation(1, 0, (), False, "t")), s)
    assert "mapexpr_filterexpr_mapexpr" in str(tasks[0].name)
    r = tasks[0].do([])
    inner = r.reader
    assert isinstance(inner, _ExprReader) and len(inner.stages) == 3  # one fused evaluation


def test_generated_source_and_typechecks():
    src, out, has_pred = jit.generate([bs.int64, bs.float32],
                                      [jit.Stage(exprs=["c0 % 7"], out=[bs.int64]), jit.Stage(pred="c0 > 2")])
    assert "const long long v1_0 = (long long)(v0_0 % 7);" in src and "keep[i]" in src and has_pred
    s = bs.Const(1, torch.arange(3))
    with pytest.raises(typecheck.TypecheckError) as ei:
        bs.MapExpr(s, ["c5 + 1"], [bs.int64])
    assert "refers to c5" in ei.value.err
    with pytest.raises(typecheck.TypecheckError):
        bs.MapExpr(s, ["c0"], [bs.int64, bs.int64])
    import numpy as np
    with pytest.raises(typecheck.TypecheckError):
        bs.FilterExpr(bs.Const(1, np.array(["a"], dtype=object)), "c0 == 1")


def test_nvrtc_compiles_cubin_for_sm_100a():
    try:
        jit._load_nvrtc()
    except RuntimeError:
        pytest.skip("libnvrtc not available")
    src, _, _ = jit.generate([bs.int64, bs.float64],
                             [jit.Stage(exprs=["c0 * 3", "c1 / 2"], out=[bs.int64, bs.float64]),
                              jit.Stage(pred="c0 % 2 == 0")])
    cubin = jit.compile_cubin(src)
    assert cubin[:4] == b"\x7fELF" and len(cubin) > 1000
    with pytest.raises(RuntimeError) as ei:
        jit.compile_cubin(src.replace("v0_0 * 3", "v0_0 * nonsense"))
    assert "nonsense" in str(ei.value)


@pytest.mark.gpu
def test_fused_kernel_matches_host_on_gpu():
    from bigslice_b200.ops import gpu
    assert gpu.available()
    g = torch.Generator().manual_seed(0)
    k = torch.randint(-1000, 1000, (200003,), generator=g)
    v = torch.randn(200003, generator=g)
    before = jit.launch_count
    with bs.start(bs.LocalDevice("cuda")) as sess:
        res = slicetest.run(chain(bs.Const(2, k, v)), session=sess)
        rows = res.rows()
    wa, wb = reference(k, v)
    assert [r[0] for r in rows] == wa.tolist() and [r[1] for r in rows] == wb.tolist()
    # per shard: the whole three-stage chain is one count kernel + one write kernel
    # (only kept rows are written; no mask, no gathers)
    assert jit.launch_count == before + 4
