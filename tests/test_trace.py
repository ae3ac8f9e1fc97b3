"""Tracing python row functions into fused-kernel stages (ops/trace.py).

The traced expression is checked three ways: interpreted on the host under C
semantics (``jit.eval_host``) against the function itself on random data,
compiled with NVRTC for sm_100a, and — on a GPU — executed as the fused kernel
against the eager result.
"""
import numpy as np
import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200 import slicefunc, slicetest
from bigslice_b200.ops import jit, trace
from bigslice_b200.slices import _ExprReader

G = torch.Generator().manual_seed(5)
N = 4099


def col(dt):
    if dt == torch.bool:
        return torch.randint(0, 2, (N,), generator=G).bool()
    if dt.is_floating_point:
        return (torch.randn(N, generator=G, dtype=torch.float64) * 50).to(dt)
    lo, hi = (0, 200) if dt == torch.uint8 else (-100, 100)
    c = torch.randint(lo, hi, (N,), generator=G).to(dt)
    return c


# (function, input dtypes) — every one must trace
CASES = [
    (lambda k, v: (k % 13, v * 2), [torch.int64, torch.int64]),
    (lambda k, v: (k // 7, v - k), [torch.int64, torch.int32]),
    (lambda k, v: (k % -5, -v), [torch.int32, torch.int16]),
    (lambda k: (k * k + 3, k ** 3, abs(k)), [torch.int64]),
    (lambda k, x: (k / 4, x / 3, k * x), [torch.int64, torch.float32]),
    (lambda k, x: (k + 0.5, x % 7.0, x % -2.5), [torch.int32, torch.float64]),
    (lambda a, b: (a & b, a | 3, a ^ b, ~a, a << 2, b >> 1), [torch.int64, torch.int64]),
    (lambda a, b: (a & b, a | b, ~a, a ^ b), [torch.bool, torch.bool]),
    (lambda k, x: (torch.where(k > 0, k, -k), torch.where(x < 0, 0.0, 1.0)), [torch.int64, torch.float32]),
    (lambda k, x: (torch.minimum(k, k * 2), torch.maximum(x, -x), x.clamp(-1.0, 1.0), k.clamp(min=0)),
     [torch.int64, torch.float64]),
    (lambda x: (torch.sqrt(x * x + 1), torch.exp(-x.abs() / 50), torch.log(x * x + 1), x.floor(), torch.ceil(x)),
     [torch.float32]),
    (lambda k: (k.to(torch.float32) * 0.5, k.float().long(), k.bool(), (k > 3).to(torch.int64), k.double()),
     [torch.int32]),
    (lambda k, v: ((k % 3 == 0) & (v > 0), (k < v) | (k == 7), k != v, k >= 0, k <= v), [torch.int64, torch.int64]),
    (lambda k: (torch.zeros_like(k), torch.ones_like(k, dtype=torch.float32), torch.full_like(k, 7), 5), [torch.int64]),
    (lambda u, k: (u + 1, u * 2, u // 3, u % 7, u + k), [torch.uint8, torch.int8]),
    (lambda k: (2 - k, 100 // (k * k + 1), 100 % (k * k + 1), 3 * k, 1 / (k * k + 1)), [torch.int64]),
    (lambda k, x: (torch.div(k, 7, rounding_mode="trunc"), torch.div(k, 7, rounding_mode="floor"),
                   torch.fmod(k, 7), torch.remainder(k, 7), torch.square(x), torch.neg(k)),
     [torch.int64, torch.float32]),
    (lambda x: (x ** 2, x ** 0.5, torch.pow(x.abs() + 1, 1.5)), [torch.float64]),
]


def _types(dts):
    return [bs.slicetype.dtype_of(d) for d in dts]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_traced_expression_equals_the_function(case):
    fn, dts = CASES[case]
    cols = [col(d) for d in dts]
    want = fn(*cols)
    want = list(want) if isinstance(want, tuple) else [want]
    want = [w if isinstance(w, torch.Tensor) else torch.full((N,), w) for w in want]
    outs = _types([w.dtype for w in want])
    st = trace.trace_map(slicefunc.Func(fn), _types(dts), outs)
    assert st is not None, "not traced"
    got, mask = jit.eval_host(cols, _types(dts), [st])
    assert mask is None
    for j, (g, w) in enumerate(zip(got, want)):
        assert g.dtype == w.dtype
        if w.is_floating_point():
            assert torch.allclose(g, w, rtol=1e-5, atol=1e-6, equal_nan=True), (case, j, st.exprs[j])
        else:
            assert torch.equal(g, w), (case, j, st.exprs[j])


def test_every_traced_kernel_compiles_for_sm_100a():
    try:
        jit._load_nvrtc()
    except RuntimeError:
        pytest.skip("libnvrtc not available")
    for fn, dts in CASES:
        cols = [col(d)[:4] for d in dts]
        want = fn(*cols)
        want = list(want) if isinstance(want, tuple) else [want]
        outs = _types([w.dtype if isinstance(w, torch.Tensor) else torch.int64 for w in want])
        st = trace.trace_map(slicefunc.Func(fn), _types(dts), outs)
        pred = trace.trace_filter(slicefunc.Func(lambda *c: c[0] != 0), _types(dts))
        src = jit.generate(_types(dts), [st, pred])[0]
        assert jit.compile_cubin(src)[:4] == b"\x7fELF"
        # and as the count/write pair that emits only the kept rows
        csrc, couts = jit.generate_compact(_types(dts), [st, pred])
        assert "bs_fused_count" in csrc and "bs_fused_write" in csrc and len(couts) == len(outs)
        assert jit.compile_cubin(csrc)[:4] == b"\x7fELF"


def test_what_is_not_traceable_runs_eagerly():
    table = torch.arange(100)
    I64 = _types([torch.int64, torch.int64])

    def branchy(k, v):
        if (k > 0).all():
            return k, v
        return v, k

    def ctxfn(ctx, k, v):
        return k, v
    for fn in (branchy, ctxfn,
               lambda k, v: (table[k], v),                 # lookup table
               lambda k, v: (torch.cumsum(k, 0), v),       # not elementwise
               lambda k, v: (k[:len(k)], v),               # needs the length
               lambda k, v: (k.cpu(), v),
               lambda k, v: (torch.rand(3), v),
               lambda k, v: (k + 2 ** 70, v),
               bs.rowwise(lambda k, v: (k, v))):
        assert trace.trace_map(slicefunc.Func(fn), I64, I64) is None
    # strings, lists
    assert trace.trace_map(slicefunc.Func(lambda s: s), [bs.string], [bs.string]) is None
    assert trace.trace_filter(slicefunc.Func(lambda k, v: True), I64) is None
    assert trace.trace_filter(slicefunc.Func(lambda k, v: k > v), I64).pred == "(c0 > c1)"
    # pure projections are left to the eager path? no: they become pass-through outputs
    st = trace.trace_map(slicefunc.Func(lambda k, v: (v, k)), I64, I64)
    assert st.exprs == ["c1", "c0"]


def test_pass_through_outputs_are_not_rewritten():
    I = _types([torch.int64, torch.int64, torch.float32])
    st = jit.Stage(exprs=["c1", "c0 * 2", "c2"], out=[bs.int64, bs.int64, bs.float32])
    src, outs, has_pred, alias = jit.generate_ex(I, [st, jit.Stage(exprs=["c0", "c1 + 1", "c2"],
                                                                   out=[bs.int64, bs.int64, bs.float32])])
    assert alias == [1, None, 2] and not has_pred
    assert "out0" not in src and "out2" not in src and "out1[i]" in src
    assert "v0_1 = in1" not in src          # the passed-through column is not even loaded
    # a dtype change is a real output
    assert jit.generate_ex(I, [jit.Stage(exprs=["c0"], out=[bs.int32])])[3] == [None]
    # nothing to launch when every output is an input
    k = torch.arange(5)
    outs, mask = jit.run_fused([k, k + 1], _types([torch.int64, torch.int64]),
                               [jit.Stage(exprs=["c1", "c0"], out=[bs.int64, bs.int64])])
    assert mask is None and outs[0].data_ptr() == (k + 1).data_ptr() or torch.equal(outs[0], k + 1)


def test_combiners_written_as_lambdas_are_recognised():
    i64, f32, b = bs.int64, bs.float32, bs.bool_
    assert trace.recognise_combiner(lambda a, b: a + b, i64) == "add"
    assert trace.recognise_combiner(lambda a, b: b + a, f32) == "add"
    assert trace.recognise_combiner(lambda x, y: x * y, i64) == "mul"
    assert trace.recognise_combiner(lambda a, b: torch.minimum(a, b), i64) == "min"
    assert trace.recognise_combiner(lambda a, b: torch.maximum(b, a), f32) == "max"
    assert trace.recognise_combiner(lambda a, b: torch.where(a < b, a, b), i64) == "min"
    assert trace.recognise_combiner(lambda a, b: torch.where(a >= b, a, b), i64) == "max"
    assert trace.recognise_combiner(lambda a, b: torch.where(a < b, a, b), f32) is None   # NaN order
    assert trace.recognise_combiner(lambda a, b: a | b, i64) == "or"
    assert trace.recognise_combiner(lambda a, b: a & b, b) == "and"
    assert trace.recognise_combiner(lambda a, b: a ^ b, i64) == "xor"
    assert trace.recognise_combiner(lambda a, b: a, i64) == "first"
    assert trace.recognise_combiner(lambda a, b: a + b + 1, i64) is None
    assert trace.recognise_combiner(lambda a, b: a - b, i64) is None
    assert trace.recognise_combiner(lambda a, b: a + b, bs.string) is None
    assert trace.recognise_combiner(lambda a, b: (a + b).to(torch.int32), i64) is None

    # and Reduce uses it: the lambda takes the builtin lowering
    r = bs.Reduce(bs.Const(2, torch.arange(10) % 3, torch.ones(10, dtype=torch.int64)), lambda a, b: a + b)
    assert r.combiner().builtin == "add"
    assert sorted(slicetest.run(r).rows()) == [(0, 4), (1, 3), (2, 3)]


def chain(s):
    s = bs.Map(s, lambda k, v: (k % 97, v * 2.0, k))
    s = bs.Filter(s, lambda k, v, o: (o % 3 != 0) & (v > -1.0))
    return bs.Map(s, lambda k, v, o: (k, v + o))


def chain_reference(k, v):
    k2, v2 = k % 97, v * 2.0
    m = (k % 3 != 0) & (v2 > -1.0)
    return k2[m], (v2 + k)[m]


def test_python_pipeline_becomes_one_fused_chain():
    k = torch.randint(-1000, 1000, (20011,), generator=G)
    v = torch.randn(20011, generator=G)
    s = chain(bs.Const(3, k, v))
    from bigslice_b200.exec.compile import ExecInvocation, compile_slice
    from bigslice_b200.func import Invocation
    tasks = compile_slice(ExecInvocation(Invocation(1, 0, (), False, "t")), s)
    r = tasks[0].do([])
    inner = getattr(r, "reader", r)
    assert isinstance(inner, _ExprReader) and len(inner.stages) == 3
    assert all(st.eager is not None for st in inner.stages)
    rows = slicetest.run(s).rows()
    wk, wv = chain_reference(k, v)
    assert [a for a, _ in rows] == wk.tolist()
    assert [b for _, b in rows] == pytest.approx(wv.tolist())


def test_a_filtered_row_never_reaches_a_later_eager_stage():
    # on the host the chain runs stage by stage: the division below would raise on the
    # rows the filter removes
    k = torch.arange(-5, 6)
    s = bs.Map(bs.Filter(bs.Const(1, k), lambda k: k != 0), lambda k: 100 // k)
    assert [r[0] for r in slicetest.run(s).rows()] == [100 // i for i in range(-5, 6) if i]


def test_first_tile_cross_check_rejects_a_function_whose_fused_form_differs(monkeypatch):
    """Drive the device branch of the chain reader on CPU tensors (the fused
    kernel is stood in for by the host interpreter) with a stage whose expression
    is not what its function computes."""
    from bigslice_b200.ops import gpu
    monkeypatch.setattr(jit, "run_fused", lambda cols, tys, stages: jit.eval_host(cols, tys, stages))
    monkeypatch.setattr(gpu, "compact_indices", lambda m: torch.nonzero(m).flatten())
    monkeypatch.setattr(gpu, "gather_rows", lambda c, i: c[i])
    calls = []

    def impure(k):
        calls.append(1)
        return (k + len(calls),)

    I = _types([torch.int64])
    bad = jit.Stage(exprs=["c0 + 1"], out=I, eager=lambda cols, n, dev: list(impure(*cols)), fn=impure)
    good_fn = lambda k: (k * 2,)   # noqa: E731
    good = jit.Stage(exprs=["c0 * 2"], out=I, eager=lambda cols, n, dev: list(good_fn(*cols)), fn=good_fn)
    k = torch.arange(10)

    class Src:
        def close(self):
            pass
    rd = _ExprReader(Src(), bs.slicetype.Type(I), [good], bs.slicetype.Type(I))
    assert torch.equal(rd._device([k], 10, k.device)[0], k * 2)
    assert trace.verified(good_fn) and not rd.eager_only

    rd = _ExprReader(Src(), bs.slicetype.Type(I), [bad], bs.slicetype.Type(I))
    rd._device([k], 10, k.device)           # first check passes by luck (calls == 1)...
    trace._verified.discard(trace._key(impure))
    out = rd._device([k], 10, k.device)     # ...the second exposes the impurity
    assert rd.eager_only and trace.rejected(impure)
    assert torch.equal(out[0], k + len(calls))
    # a rejected function is not traced again
    m = bs.Map(bs.Const(1, k), impure, out=[bs.int64])
    assert m.traced() is None


@pytest.mark.gpu
def test_traced_python_pipeline_runs_as_one_kernel_on_gpu():
    from bigslice_b200.ops import gpu
    assert gpu.available()
    k = torch.randint(-1000, 1000, (300007,), generator=G)
    v = torch.randn(300007, generator=G)
    before = jit.launch_count
    with bs.start(bs.LocalDevice("cuda")) as sess:
        rows = slicetest.run(chain(bs.Const(2, k, v)), session=sess).rows()
    wk, wv = chain_reference(k, v)
    assert [a for a, _ in rows] == wk.tolist()
    assert [b for _, b in rows] == pytest.approx(wv.tolist(), rel=1e-5)
    # per shard: count kernel + write kernel for the three python stages together
    assert jit.launch_count == before + 4


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(CASES)))
def test_fused_kernel_equals_the_function_on_gpu(case):
    fn, dts = CASES[case]
    cols = [col(d).cuda() for d in dts]
    want = fn(*cols)
    want = list(want) if isinstance(want, tuple) else [want]
    want = [w if isinstance(w, torch.Tensor) else torch.full((N,), w, device="cuda") for w in want]
    st = trace.trace_map(slicefunc.Func(fn), _types(dts), _types([w.dtype for w in want]))
    got, _ = jit.run_fused(cols, _types(dts), [st])
    for j, (g, w) in enumerate(zip(got, want)):
        if w.is_floating_point():
            assert torch.allclose(g, w, rtol=1e-4, atol=1e-5, equal_nan=True), (case, j, st.exprs[j])
        else:
            assert torch.equal(g, w), (case, j, st.exprs[j])


def test_python_fold_functions_trace_to_the_fold_kernel_expression():
    i64, f32 = bs.int64, bs.float32
    fn = lambda acc, v, w: (acc * 31 + v) % 1009 + (w > 0)   # noqa: E731
    e = trace.trace_fold(slicefunc.Func(fn), i64, [i64, f32])
    assert e is not None and "acc" in e and "c1" in e and "c2" in e
    # host interpretation of the expression == the python function folded row by row
    keys = torch.randint(0, 50, (3001,), generator=G)
    v = torch.randint(-20, 20, (3001,), generator=G)
    w = torch.randn(3001, generator=G)
    order = torch.sort(keys, stable=True).indices
    sk, sv, sw = keys[order], v[order], w[order]
    starts = torch.cat([torch.tensor([True]), sk[1:] != sk[:-1]])
    off = torch.cat([torch.nonzero(starts).flatten(), torch.tensor([len(sk)])])
    got = jit.fold_host(i64, [sv, sw], off, e)
    want = {}
    for k, a, b in zip(sk.tolist(), sv.tolist(), sw.tolist()):
        want[k] = (want.get(k, 0) * 31 + a) % 1009 + (1 if b > 0 else 0)
    assert got.tolist() == [want[k] for k in sorted(want)]
    try:
        jit._load_nvrtc()
        assert jit.compile_cubin(jit.generate_fold(i64, [i64, f32], e))[:4] == b"\x7fELF"
    except RuntimeError:
        pass
    # the slice carries the traced expression; on the host the function itself still runs
    s = bs.Fold(bs.Const(2, keys, v, w), fn, acc=bs.int64)
    assert s.traced_expr() == e
    assert dict(slicetest.run(s).rows()) == want
    # lambda acc, v: acc + v is the builtin add
    s = bs.Fold(bs.Const(2, keys, v), lambda acc, v: acc + v)
    assert s.func.builtin == "add"
    assert bs.Fold(bs.Const(2, keys, v), lambda acc, v: acc).func.builtin is None
    # not traceable: stays a python fold
    assert bs.Fold(bs.Const(2, keys, v), lambda acc, v: acc + int(v.sum())).traced_expr() is None


@pytest.mark.gpu
def test_traced_python_fold_runs_as_the_fold_kernel_on_gpu():
    keys = torch.cat([torch.zeros(50000, dtype=torch.int64), torch.randint(1, 500, (150000,), generator=G)])
    v = torch.randint(0, 1000, (200000,), generator=G)
    fn = lambda acc, v: (acc * 31 + v) % 1000003   # noqa: E731
    before = jit.launch_count
    with bs.start(bs.LocalDevice("cuda")) as sess:
        # the head key has 50 000 rows: far beyond MAX_FOLD_ROUNDS, fine for the kernel
        rows = dict(slicetest.run(bs.Fold(bs.Const(1, keys, v), fn), session=sess).rows())
    assert jit.launch_count > before
    want = {}
    for k, a in zip(keys.tolist(), v.tolist()):
        want[k] = (want.get(k, 0) * 31 + a) % 1000003
    assert rows == want
