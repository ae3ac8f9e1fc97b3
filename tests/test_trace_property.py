"""Property tests of the python-function tracer (ops/trace.py).

Random expression trees over int64 / int32 / float32 columns are turned into
python functions, traced, and checked against what the function itself computes
(eager PyTorch):

* through the host interpreter of the expression language (``jit.eval_host``) —
  hypothesis-driven;
* through **the generated C source itself**, compiled for the host with g++ (the
  CUDA keywords are defined away and the grid-stride loop runs as one thread), so
  literal formatting, casts, promotion and the floor-mod/floor-div helpers are
  checked under real C semantics without a GPU.
"""
import ctypes
import random
import shutil
import subprocess

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import bigslice_b200 as bs
from bigslice_b200 import slicefunc
from bigslice_b200.ops import jit, trace

DTS = [torch.int64, torch.int32, torch.float32, torch.int8, torch.uint8, torch.int16, torch.float64]   # c0 .. c6
N = 513
G = torch.Generator().manual_seed(11)
COLS = [torch.randint(-100, 100, (N,), generator=G),
        torch.randint(-100, 100, (N,), generator=G).to(torch.int32),
        (torch.randn(N, generator=G) * 20).to(torch.float32),
        torch.randint(-100, 100, (N,), generator=G).to(torch.int8),
        torch.randint(0, 250, (N,), generator=G).to(torch.uint8),
        torch.randint(-3000, 3000, (N,), generator=G).to(torch.int16),
        torch.randn(N, generator=G, dtype=torch.float64) * 20]


def is_float(x):
    return (x.dtype if isinstance(x, (torch.Tensor, trace.Sym)) else torch.result_type(x, x)).is_floating_point


def ev(t, cols):
    """Evaluate a tree on tensors *or* symbolic columns (the same code path the
    user's function would take)."""
    op = t[0]
    if op == "col":
        return cols[t[1]]
    if op == "const":
        return t[1]
    a = ev(t[1], cols)
    if op in ("neg", "abs", "inv", "f32", "i64", "i32", "sqrtabs", "floor"):
        if not isinstance(a, (torch.Tensor, trace.Sym)):
            a = cols[0] * 0 + a
        if op == "neg":
            return -a
        if op == "abs":
            return abs(a)
        if op == "inv":
            return ~a if not is_float(a) else -a
        if op == "f32":
            return a.to(torch.float32)
        if op == "i64":
            return a.to(torch.int64)
        if op == "i32":
            return a.to(torch.int32)
        if op == "sqrtabs":
            return torch.sqrt(abs(a) + 1)
        return a.floor() if is_float(a) else a
    b = ev(t[2], cols)
    if not isinstance(a, (torch.Tensor, trace.Sym)) and not isinstance(b, (torch.Tensor, trace.Sym)):
        a = cols[0] * 0 + a
    if op == "add":
        return a + b
    if op == "sub":
        return a - b
    if op == "mul":
        return a * b
    if op == "truediv":
        return a / (abs(b) + 1)
    if op in ("floordiv", "mod"):
        if is_float(a) or is_float(b):
            return a % 7.5 if op == "mod" else a - b
        d = abs(b) + 3 if t[3] else -(abs(b) + 3)
        return a // d if op == "floordiv" else a % d
    if op in ("and", "or", "xor"):
        if is_float(a) or is_float(b):
            return a + b
        return (a & b) if op == "and" else (a | b) if op == "or" else (a ^ b)
    if op in ("shl", "shr"):
        if is_float(a):
            return a * 2
        return (a << t[3]) if op == "shl" else (a >> t[3])
    if op in ("min", "max"):
        if not isinstance(b, (torch.Tensor, trace.Sym)):
            b = a * 0 + b
        if not isinstance(a, (torch.Tensor, trace.Sym)):
            a = b * 0 + a
        return torch.minimum(a, b) if op == "min" else torch.maximum(a, b)
    if op == "where":
        c = ev(t[3], cols)
        cmp = [a < b, a <= b, a == b, a != b][t[4]]
        return torch.where(cmp, c, a)
    raise AssertionError(op)


def trees(depth):
    leaf = st.one_of(st.tuples(st.just("col"), st.integers(0, 6)),
                     st.tuples(st.just("const"), st.one_of(st.integers(-9, 9), st.sampled_from([0.5, -2.25, 3.0]))))
    if depth == 0:
        return leaf
    sub = trees(depth - 1)
    un = st.tuples(st.sampled_from(["neg", "abs", "inv", "f32", "i64", "i32", "sqrtabs", "floor"]), sub)
    bi = st.tuples(st.sampled_from(["add", "sub", "mul", "truediv", "and", "or", "xor", "min", "max"]), sub, sub)
    dv = st.tuples(st.sampled_from(["floordiv", "mod"]), sub, sub, st.booleans())
    sh = st.tuples(st.sampled_from(["shl", "shr"]), sub, sub, st.integers(0, 3))
    wh = st.tuples(st.just("where"), sub, sub, sub, st.integers(0, 3))
    return st.one_of(leaf, un, bi, dv, sh, wh)


OUT_OK = (torch.int64, torch.int32, torch.float32, torch.bool, torch.float64, torch.int8, torch.uint8, torch.int16)


def _types(dts):
    return [bs.slicetype.dtype_of(d) for d in dts]


def eager(tree):
    r = ev(tree, COLS)
    if not isinstance(r, torch.Tensor):
        r = COLS[0] * 0 + r
    return r


def traced_stage(tree, want):
    fn = lambda *cols: (ev(tree, cols),)   # noqa: E731
    return trace.trace_map(slicefunc.Func(fn), _types(DTS), _types([want.dtype]))


def same(got, want):
    if want.is_floating_point():
        return torch.allclose(got, want, rtol=1e-5, atol=1e-5, equal_nan=True)
    return torch.equal(got, want)


@settings(max_examples=250, deadline=None, suppress_health_check=list(HealthCheck))
@given(trees(3))
def test_traced_random_expressions_equal_the_function(tree):
    try:
        want = eager(tree)
    except (RuntimeError, TypeError, ZeroDivisionError):
        return   # the function itself is not valid torch (e.g. bitwise ops on mixed bool/float)
    if want.dtype not in OUT_OK:
        return
    stg = traced_stage(tree, want)
    if stg is None:
        return   # declined (e.g. a negative scalar against a uint8 column): the function runs eagerly
    (got,), _ = jit.eval_host(COLS, _types(DTS), [stg])
    assert got.dtype == want.dtype
    assert same(got, want), (tree, stg.exprs[0])


HOST_SHIM = r"""
#include <cmath>
#include <cstdlib>
using std::fabs; using std::floor; using std::ceil; using std::sqrt; using std::exp; using std::log;
using std::pow; using std::fmod;
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(x)
struct bs_dim3 { long long x; };
static const bs_dim3 blockIdx{0}, blockDim{1}, threadIdx{0}, gridDim{1};
"""


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_generated_c_source_equals_the_function_under_real_c_semantics(tmp_path):
    rnd = random.Random(7)
    cases = []
    tries = 0
    while len(cases) < 150 and tries < 4000:
        tries += 1
        tree = _random_tree(rnd, 3)
        try:
            want = eager(tree)
        except (RuntimeError, TypeError, ZeroDivisionError):
            continue
        if want.dtype not in OUT_OK:
            continue
        stg = traced_stage(tree, want)
        if stg is None:
            continue
        cases.append((tree, want, stg))
    assert len(cases) >= 100
    src = [HOST_SHIM]
    for k, (_, _, stg) in enumerate(cases):
        s = jit.generate(_types(DTS), [stg])[0]
        s = s[len(jit.PRELUDE):] if k else s          # the helpers once
        src.append(s.replace("bs_fused(", "bs_case%d(" % k))
    cpp = tmp_path / "cases.cpp"
    cpp.write_text("\n".join(src))
    so = tmp_path / "cases.so"
    subprocess.run(["g++", "-O1", "-fwrapv", "-fno-fast-math", "-shared", "-fPIC", "-w", "-o", str(so), str(cpp)],
                   check=True)
    lib = ctypes.CDLL(str(so))
    ins = [c.numpy() for c in COLS]
    for k, (tree, want, stg) in enumerate(cases):
        _, _, _, alias = jit.generate_ex(_types(DTS), [stg])
        if alias[0] is not None:
            got = COLS[alias[0]]
        else:
            out = np.empty(N, dtype={torch.int64: np.int64, torch.int32: np.int32, torch.float32: np.float32,
                                     torch.bool: np.bool_, torch.float64: np.float64, torch.int8: np.int8,
                                     torch.uint8: np.uint8, torch.int16: np.int16}[want.dtype])
            fn = getattr(lib, "bs_case%d" % k)
            fn.restype = None
            fn(ctypes.c_longlong(N), *[ctypes.c_void_p(a.ctypes.data) for a in ins], ctypes.c_void_p(out.ctypes.data))
            got = torch.from_numpy(out)
        assert same(got, want), (tree, stg.exprs[0])


def _random_tree(rnd, depth):
    if depth == 0 or rnd.random() < 0.2:
        if rnd.random() < 0.6:
            return ("col", rnd.randrange(7))
        return ("const", rnd.choice([rnd.randint(-9, 9), 0.5, -2.25, 3.0]))
    kind = rnd.random()
    sub = lambda: _random_tree(rnd, depth - 1)   # noqa: E731
    if kind < 0.25:
        return (rnd.choice(["neg", "abs", "inv", "f32", "i64", "i32", "sqrtabs", "floor"]), sub())
    if kind < 0.6:
        return (rnd.choice(["add", "sub", "mul", "truediv", "and", "or", "xor", "min", "max"]), sub(), sub())
    if kind < 0.75:
        return (rnd.choice(["floordiv", "mod"]), sub(), sub(), rnd.random() < 0.5)
    if kind < 0.85:
        return (rnd.choice(["shl", "shr"]), sub(), sub(), rnd.randrange(4))
    return ("where", sub(), sub(), sub(), rnd.randrange(4))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_generated_fold_kernel_source_under_real_c_semantics(tmp_path):
    """The segmented fold kernel (one thread per key) compiled for the host: the
    traced python fold must give what folding the function row by row gives."""
    i64, f32 = bs.int64, bs.float32
    fn = lambda acc, v, w: (acc * 31 + v) % 1009 - (w > 0) * 2 + abs(v) // 3   # noqa: E731
    expr = trace.trace_fold(slicefunc.Func(fn), i64, [i64, f32])
    assert expr is not None
    src = HOST_SHIM + jit.generate_fold(i64, [i64, f32], expr)
    cpp, so = tmp_path / "fold.cpp", tmp_path / "fold.so"
    cpp.write_text(src)
    subprocess.run(["g++", "-O1", "-fwrapv", "-shared", "-fPIC", "-w", "-o", str(so), str(cpp)], check=True)
    lib = ctypes.CDLL(str(so))
    rnd = np.random.default_rng(5)
    lens = rnd.integers(0, 40, 200)
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = int(seg[-1])
    v = rnd.integers(-50, 50, n).astype(np.int64)
    w = rnd.standard_normal(n).astype(np.float32)
    out = np.empty(len(lens), dtype=np.int64)
    lib.bs_fused.restype = None
    lib.bs_fused(ctypes.c_longlong(len(lens)), ctypes.c_void_p(seg.ctypes.data), ctypes.c_void_p(v.ctypes.data),
                 ctypes.c_void_p(w.ctypes.data), ctypes.c_void_p(out.ctypes.data))
    want = []
    for s in range(len(lens)):
        acc = 0
        for i in range(seg[s], seg[s + 1]):
            acc = (acc * 31 + int(v[i])) % 1009 - (2 if w[i] > 0 else 0) + abs(int(v[i])) // 3
        want.append(acc)
    assert out.tolist() == want
    # and the host interpreter used on CPU executors agrees
    got = jit.fold_host(i64, [torch.from_numpy(v), torch.from_numpy(w)], torch.from_numpy(seg), expr)
    assert got.tolist() == want
