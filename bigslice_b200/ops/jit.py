"""NVRTC-compiled fused elementwise kernels for expression Maps and Filters.

``MapExpr(slice, ["c0 % 1000", "c1 * 2"], out=[int64, int64])`` and
``FilterExpr(slice, "c0 % 2 == 0")`` describe row functions as C expressions
over the input columns ``c0, c1, ...``.  Consecutive expression stages that the
compiler pipelines into one task are **fused into a single sm_100a kernel**:
the chain is composed symbolically, compiled once with NVRTC
(``--gpu-architecture=sm_100a``) into a cubin, cached by source, loaded through
the driver API and launched on the current stream.  One pass over the tile
evaluates every stage.  A chain with Filters compiles to a *count* kernel and a
*write* kernel (``generate_compact``): per-CTA kept-row counts, a prefix sum,
then every kept row is written straight to its final position — in row order,
with no mask, no full-size temporaries and no gathers.

On executors without a GPU the same expressions are evaluated with torch ops
under C semantics (truncating integer division/modulo), so results agree.

This is tier (b) of the user-function model in SURVEY §7.3 (tier (a): builtin
functors; tier (c): vectorised python callables) and replaces the per-row
reflective calls of ``mapReader``/``filterReader`` (reference
``slice.go:621-632,700-720``).
"""
from __future__ import annotations

import ctypes
import hashlib
import re
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .. import slicetype
from ..slicetype import DType

CTYPE = {"int8": "signed char", "int16": "short", "int32": "int", "int64": "long long",
         "uint8": "unsigned char", "uint16": "unsigned short", "uint32": "unsigned int",
         "uint64": "unsigned long long", "float32": "float", "float64": "double",
         "bool": "bool"}

_VAR = re.compile(r"\bc(\d+)\b")


class Stage:
    """One Map (``exprs``/``out``) or Filter (``pred``) step."""

    def __init__(self, exprs: Optional[Sequence[str]] = None, out: Optional[Sequence[DType]] = None,
                 pred: Optional[str] = None, eager=None, fn=None):
        self.exprs = list(exprs) if exprs is not None else None
        self.out = [slicetype.dtype_of(o) for o in out] if out is not None else None
        self.pred = pred
        #: for stages traced from a python function (ops/trace.py): ``eager(cols, n, device)``
        #: runs the original function (host frames, first-tile cross-check); ``fn`` is that function
        self.eager = eager
        self.fn = fn


#: helpers every generated kernel may call: python-semantics ``%`` and ``//`` (floor, sign of
#: the divisor), NaN-propagating min/max, abs — what traced python functions lower to.
PRELUDE = r"""
template <class T> __device__ __forceinline__ T bs_abs(T a) { return a < (T)0 ? (T)(-a) : a; }
__device__ __forceinline__ float bs_abs(float a) { return fabsf(a); }
__device__ __forceinline__ double bs_abs(double a) { return fabs(a); }
template <class T> __device__ __forceinline__ T bs_min(T a, T b) {
  return (a != a) ? a : ((b != b) ? b : (b < a ? b : a)); }
template <class T> __device__ __forceinline__ T bs_max(T a, T b) {
  return (a != a) ? a : ((b != b) ? b : (a < b ? b : a)); }
template <class T> __device__ __forceinline__ T bs_floordiv(T a, T b) {
  const T q = (T)(a / b);
  return ((a % b) != 0 && ((a < (T)0) != (b < (T)0))) ? (T)(q - 1) : q; }
template <class T> __device__ __forceinline__ T bs_pymod(T a, T b) {
  const T r = (T)(a % b);
  return (r != (T)0 && ((r < (T)0) != (b < (T)0))) ? (T)(r + b) : r; }
__device__ __forceinline__ float bs_pymod(float a, float b) {
  const float r = fmodf(a, b);
  return (r != 0.f && ((r < 0.f) != (b < 0.f))) ? r + b : r; }
__device__ __forceinline__ double bs_pymod(double a, double b) {
  const double r = fmod(a, b);
  return (r != 0.0 && ((r < 0.0) != (b < 0.0))) ? r + b : r; }
"""

_PASS = re.compile(r"^c(\d+)$")


def check_types(dts: Sequence[DType]) -> None:
    for d in dts:
        if d.name not in CTYPE or d.kind == "vec":
            raise TypeError("expression stages support scalar numeric/bool columns, not %s" % d)


def _definitions(in_types: Sequence[DType], stages: Sequence[Stage], index: str):
    """The chain as C definitions over row ``index``: (lines, output value names,
    output dtypes, predicate names, number of leading load lines)."""
    check_types(in_types)
    cur = list(in_types)
    body = []
    names = ["v0_%d" % j for j in range(len(cur))]
    for j, t in enumerate(cur):
        body.append("    const %s v0_%d = in%d[%s];" % (CTYPE[t.name], j, j, index))
    preds = []
    for s, st in enumerate(stages, 1):
        def sub(expr, names=names):
            def rep(m):
                k = int(m.group(1))
                if k >= len(names):
                    raise ValueError("expression refers to c%d but the stage has %d columns" % (k, len(names)))
                return names[k]
            return _VAR.sub(rep, expr)
        if st.pred is not None:
            body.append("    const bool p%d = (%s);" % (s, sub(st.pred)))
            preds.append("p%d" % s)
        else:
            check_types(st.out)
            new = []
            for j, (e, t) in enumerate(zip(st.exprs, st.out)):
                m = _PASS.match(e.strip())
                if m and int(m.group(1)) < len(names) and cur[int(m.group(1))].name == t.name:
                    new.append(names[int(m.group(1))])   # pass-through: no new value
                    continue
                body.append("    const %s v%d_%d = (%s)(%s);" % (CTYPE[t.name], s, j, CTYPE[t.name], sub(e)))
                new.append("v%d_%d" % (s, j))
            names, cur = new, list(st.out)
    return body, names, cur, preds, len(in_types)


def _drop_unused_loads(body: List[str], nloads: int, extra: Sequence[str]) -> List[str]:
    used = set(int(m.group(1)) for m in re.finditer(r"\bv0_(\d+)\b", "\n".join(body[nloads:] + list(extra))))
    return [ln for j, ln in enumerate(body) if j >= nloads or j in used]


def generate_ex(in_types: Sequence[DType], stages: Sequence[Stage]
                ) -> Tuple[str, List[DType], bool, List[Optional[int]]]:
    """CUDA source of the fused kernel, its output dtypes, whether it has a
    predicate, and for every output the input column it merely passes through
    (such outputs get no kernel parameter: the caller reuses the input column)."""
    body, names, cur, preds, nloads = _definitions(in_types, stages, "i")
    alias: List[Optional[int]] = [int(nm[3:]) if nm.startswith("v0_") else None for nm in names]
    body = _drop_unused_loads(body, nloads, [nm for nm, a in zip(names, alias) if a is None])
    params = ["long long n"]
    params += ["const %s* __restrict__ in%d" % (CTYPE[t.name], j) for j, t in enumerate(in_types)]
    params += ["%s* __restrict__ out%d" % (CTYPE[t.name], j) for j, t in enumerate(cur) if alias[j] is None]
    if preds:
        params.append("unsigned char* __restrict__ keep")
    for j, nm in enumerate(names):
        if alias[j] is None:
            body.append("    out%d[i] = %s;" % (j, nm))
    if preds:
        body.append("    keep[i] = (%s) ? 1 : 0;" % " && ".join(preds))
    src = (PRELUDE + "extern \"C\" __global__ void __launch_bounds__(256) bs_fused(%s) {\n"
           "  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;\n"
           "       i += (long long)gridDim.x * blockDim.x) {\n%s\n  }\n}\n"
           % (", ".join(params), "\n".join(body)))
    return src, cur, bool(preds), alias


#: rows per CTA of the compacting kernels: COMPACT_ITEMS iterations of 256 consecutive rows
COMPACT_ITEMS = 8
COMPACT_CHUNK = 256 * COMPACT_ITEMS


def generate_compact(in_types: Sequence[DType], stages: Sequence[Stage]) -> Tuple[str, List[DType]]:
    """Source of the two kernels that evaluate a chain **with predicates** and
    write only the kept rows, in row order, without materialising the unfiltered
    outputs or a mask:

    * ``bs_fused_count``  one CTA per chunk of %d rows: the number of rows the chunk keeps;
    * ``bs_fused_write``  given the exclusive prefix sum of those counts, the chain is
      evaluated again and every kept row is written straight to its final position
      (rank inside the CTA from warp ballots + a shared-memory scan of the 8 warp counts).

    Rows past ``n`` re-read row ``n-1`` (all threads stay converged for the ballots)
    and are never kept.""" % COMPACT_CHUNK
    body, names, cur, preds, nloads = _definitions(in_types, stages, "ii")
    if not preds:
        raise ValueError("generate_compact needs at least one predicate stage")
    keep = "(i < n) && (%s)" % " && ".join(preds)
    ins = ["const %s* __restrict__ in%d" % (CTYPE[t.name], j) for j, t in enumerate(in_types)]
    head = ("    const long long i = base + (long long)it * 256 + tid;\n"
            "    const long long ii = i < n ? i : n - 1;\n")
    count_body = "\n".join(_drop_unused_loads(body, nloads, []))
    write_body = "\n".join(_drop_unused_loads(body, nloads, names))
    stores = "\n".join("      out%d[o] = %s;" % (j, nm) for j, nm in enumerate(names))
    outs = ["%s* __restrict__ out%d" % (CTYPE[t.name], j) for j, t in enumerate(cur)]
    src = PRELUDE + (
        "#define BS_ITEMS %d\n"
        "extern \"C\" __global__ void __launch_bounds__(256) bs_fused_count(%s) {\n"
        "  __shared__ int wsum[8];\n"
        "  const int tid = threadIdx.x;\n"
        "  const long long base = (long long)blockIdx.x * (256 * BS_ITEMS);\n"
        "  int kept = 0;\n"
        "#pragma unroll\n"
        "  for (int it = 0; it < BS_ITEMS; ++it) {\n%s%s\n"
        "    const bool keep = %s;\n"
        "    kept += __popc(__ballot_sync(0xffffffffu, keep));\n"
        "  }\n"
        "  if ((tid & 31) == 0) wsum[tid >> 5] = kept;\n"
        "  __syncthreads();\n"
        "  if (tid == 0) {\n"
        "    int t = 0;\n"
        "    for (int w = 0; w < 8; ++w) t += wsum[w];\n"
        "    counts[blockIdx.x] = t;\n"
        "  }\n"
        "}\n"
        "extern \"C\" __global__ void __launch_bounds__(256) bs_fused_write(%s) {\n"
        "  __shared__ int wcnt[8];\n"
        "  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;\n"
        "  const long long base = (long long)blockIdx.x * (256 * BS_ITEMS);\n"
        "  const long long pos0 = offs[blockIdx.x];\n"
        "  int run = 0;\n"
        "#pragma unroll\n"
        "  for (int it = 0; it < BS_ITEMS; ++it) {\n%s%s\n"
        "    const bool keep = %s;\n"
        "    const unsigned b = __ballot_sync(0xffffffffu, keep);\n"
        "    if (lane == 0) wcnt[warp] = __popc(b);\n"
        "    __syncthreads();\n"
        "    int before = 0, tot = 0;\n"
        "#pragma unroll\n"
        "    for (int w = 0; w < 8; ++w) {\n"
        "      const int c = wcnt[w];\n"
        "      if (w < warp) before += c;\n"
        "      tot += c;\n"
        "    }\n"
        "    if (keep) {\n"
        "      const long long o = pos0 + run + before + __popc(b & ((1u << lane) - 1u));\n%s\n"
        "    }\n"
        "    run += tot;\n"
        "    __syncthreads();\n"
        "  }\n"
        "}\n"
        % (COMPACT_ITEMS,
           ", ".join(["long long n"] + ins + ["int* __restrict__ counts"]), head, count_body, keep,
           ", ".join(["long long n"] + ins + ["const long long* __restrict__ offs"] + outs), head, write_body,
           keep, stores))
    return src, cur


def generate(in_types: Sequence[DType], stages: Sequence[Stage]) -> Tuple[str, List[DType], bool]:
    """``generate_ex`` without the pass-through map."""
    return generate_ex(in_types, stages)[:3]


# ---------------------------------------------------------------------------
# NVRTC (compile) and driver API (load/launch) through ctypes

_nvrtc = None
_cuda = None
_cache: Dict[str, bytes] = {}
_funcs: Dict[Tuple[int, str], ctypes.c_void_p] = {}
_lock = threading.Lock()


def _load_nvrtc():
    global _nvrtc
    if _nvrtc is None:
        for name in ("libnvrtc.so", "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"):
            try:
                _nvrtc = ctypes.CDLL(name)
                break
            except OSError:
                continue
        if _nvrtc is None:
            raise RuntimeError("NVRTC (libnvrtc) not found")
    return _nvrtc


def compile_cubin(src: str, arch: str = "sm_100a") -> bytes:
    """Compile CUDA source to a cubin for ``arch`` with NVRTC (cached)."""
    key = hashlib.sha1((arch + src).encode()).hexdigest()
    with _lock:
        if key in _cache:
            return _cache[key]
    rt = _load_nvrtc()
    prog = ctypes.c_void_p()
    rc = rt.nvrtcCreateProgram(ctypes.byref(prog), src.encode(), b"bs_fused.cu", 0, None, None)
    if rc != 0:
        raise RuntimeError("nvrtcCreateProgram failed: %d" % rc)
    opts = [b"--gpu-architecture=" + arch.encode(), b"-lineinfo", b"--std=c++17"]
    arr = (ctypes.c_char_p * len(opts))(*opts)
    rc = rt.nvrtcCompileProgram(prog, len(opts), arr)
    if rc != 0:
        n = ctypes.c_size_t()
        rt.nvrtcGetProgramLogSize(prog, ctypes.byref(n))
        log = ctypes.create_string_buffer(n.value)
        rt.nvrtcGetProgramLog(prog, log)
        rt.nvrtcDestroyProgram(ctypes.byref(prog))
        raise RuntimeError("expression does not compile:\n%s\n--- source ---\n%s"
                           % (log.value.decode(errors="replace"), src))
    n = ctypes.c_size_t()
    rt.nvrtcGetCUBINSize(prog, ctypes.byref(n))
    buf = ctypes.create_string_buffer(n.value)
    rt.nvrtcGetCUBIN(prog, buf)
    rt.nvrtcDestroyProgram(ctypes.byref(prog))
    cubin = buf.raw
    with _lock:
        _cache[key] = cubin
    return cubin


def _driver():
    global _cuda
    if _cuda is None:
        _cuda = ctypes.CDLL("libcuda.so.1")
    return _cuda


def _function(src: str, entry: str = "bs_fused") -> ctypes.c_void_p:
    dev = torch.cuda.current_device()
    key = (dev, hashlib.sha1(src.encode()).hexdigest() + ":" + entry)
    with _lock:
        if key in _funcs:
            return _funcs[key]
    cubin = compile_cubin(src)
    cu = _driver()
    torch.cuda.current_stream()  # make sure the primary context exists and is current
    mod = ctypes.c_void_p()
    rc = cu.cuModuleLoadData(ctypes.byref(mod), cubin)
    if rc != 0:
        raise RuntimeError("cuModuleLoadData failed: %d" % rc)
    fn = ctypes.c_void_p()
    rc = cu.cuModuleGetFunction(ctypes.byref(fn), mod, entry.encode())
    if rc != 0:
        raise RuntimeError("cuModuleGetFunction failed: %d" % rc)
    with _lock:
        _funcs[key] = fn
    return fn


launch_count = 0


def run_fused(cols: Sequence[torch.Tensor], in_types: Sequence[DType], stages: Sequence[Stage]):
    """Launch the fused kernel over device columns; returns (outputs, mask)."""
    global launch_count
    src, out_types, has_pred, alias = generate_ex(in_types, stages)
    n = int(cols[0].shape[0]) if cols else 0
    dev = cols[0].device
    cols = [c.contiguous() for c in cols]
    outs = [cols[a] if a is not None else torch.empty(n, dtype=t.torch_dtype, device=dev)
            for t, a in zip(out_types, alias)]
    mask = torch.empty(n, dtype=torch.uint8, device=dev) if has_pred else None
    if n == 0 or (not has_pred and all(a is not None for a in alias)):
        return outs, mask
    fn = _function(src)
    args = [ctypes.c_longlong(n)] + [ctypes.c_void_p(c.data_ptr()) for c in cols] + \
           [ctypes.c_void_p(o.data_ptr()) for o, a in zip(outs, alias) if a is None]
    if mask is not None:
        args.append(ctypes.c_void_p(mask.data_ptr()))
    argv = (ctypes.c_void_p * len(args))(*[ctypes.cast(ctypes.pointer(a), ctypes.c_void_p) for a in args])
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    blocks = min((n + 255) // 256, sms * 16)
    rc = _driver().cuLaunchKernel(fn, blocks, 1, 1, 256, 1, 1, 0,
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), argv, None)
    if rc != 0:
        raise RuntimeError("cuLaunchKernel failed: %d" % rc)
    launch_count += 1
    return outs, mask


def _launch(fn, blocks: int, threads: int, args) -> None:
    global launch_count
    argv = (ctypes.c_void_p * len(args))(*[ctypes.cast(ctypes.pointer(a), ctypes.c_void_p) for a in args])
    rc = _driver().cuLaunchKernel(fn, blocks, 1, 1, threads, 1, 1, 0,
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), argv, None)
    if rc != 0:
        raise RuntimeError("cuLaunchKernel failed: %d" % rc)
    launch_count += 1


def run_fused_compact(cols: Sequence[torch.Tensor], in_types: Sequence[DType], stages: Sequence[Stage]
                      ) -> List[torch.Tensor]:
    """Evaluate a chain that contains Filters and return only the kept rows, in
    row order: count kernel -> prefix sum of the per-CTA counts -> write kernel
    (``generate_compact``).  One host read (the number of kept rows, to size the
    outputs); no mask, no full-size temporaries, no gathers."""
    src, out_types = generate_compact(in_types, stages)
    n = int(cols[0].shape[0]) if cols else 0
    dev = cols[0].device
    if n == 0:
        return [torch.empty(0, dtype=t.torch_dtype, device=dev) for t in out_types]
    cols = [c.contiguous() for c in cols]
    nblocks = (n + COMPACT_CHUNK - 1) // COMPACT_CHUNK
    counts = torch.empty(nblocks, dtype=torch.int32, device=dev)
    ins = [ctypes.c_void_p(c.data_ptr()) for c in cols]
    _launch(_function(src, "bs_fused_count"), nblocks, 256,
            [ctypes.c_longlong(n)] + ins + [ctypes.c_void_p(counts.data_ptr())])
    incl = torch.cumsum(counts, 0, dtype=torch.int64)
    total = int(incl[-1])
    outs = [torch.empty(total, dtype=t.torch_dtype, device=dev) for t in out_types]
    if total == 0:
        return outs
    offs = (incl - counts).contiguous()
    _launch(_function(src, "bs_fused_write"), nblocks, 256,
            [ctypes.c_longlong(n)] + ins + [ctypes.c_void_p(offs.data_ptr())] +
            [ctypes.c_void_p(o.data_ptr()) for o in outs])
    return outs


# ---------------------------------------------------------------------------
# ordered per-key folds: ``acc = expr(acc, c1, ..., cn)`` in row order

def generate_fold(acc_type: DType, val_types: Sequence[DType], expr: str) -> str:
    """CUDA source of the segmented fold kernel: rows are sorted by key (stable),
    ``seg[s] .. seg[s+1]`` delimits key ``s``; one thread folds one key's rows in
    order, starting from the zero accumulator.  Arbitrary (non-associative)
    expressions are allowed, so a key's rows are inherently sequential — the
    parallelism is across keys."""
    check_types([acc_type] + list(val_types))
    A = CTYPE[acc_type.name]
    names = {0: None}

    def rep(m):
        k = int(m.group(1))
        if k == 0 or k > len(val_types):
            raise ValueError("fold expression may use acc and c1..c%d (c0 is the key), not c%d"
                             % (len(val_types), k))
        return "v%d" % k
    body = _VAR.sub(rep, expr)
    loads = "\n".join("      const %s v%d = in%d[i];" % (CTYPE[t.name], j + 1, j + 1)
                      for j, t in enumerate(val_types))
    params = ["long long nseg", "const long long* __restrict__ seg"]
    params += ["const %s* __restrict__ in%d" % (CTYPE[t.name], j + 1) for j, t in enumerate(val_types)]
    params.append("%s* __restrict__ out" % A)
    return (PRELUDE + "extern \"C\" __global__ void __launch_bounds__(128) bs_fused(%s) {\n"
            "  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < nseg;\n"
            "       s += (long long)gridDim.x * blockDim.x) {\n"
            "    %s acc = (%s)0;\n"
            "    const long long hi = seg[s + 1];\n"
            "    for (long long i = seg[s]; i < hi; ++i) {\n%s\n"
            "      acc = (%s)(%s);\n"
            "    }\n"
            "    out[s] = acc;\n  }\n}\n" % (", ".join(params), A, A, loads, A, body))


def run_fold(acc_type: DType, val_cols: Sequence[torch.Tensor], val_types: Sequence[DType],
             seg_offsets: torch.Tensor, expr: str) -> torch.Tensor:
    """Launch the fold kernel over sorted device columns; one output per segment."""
    global launch_count
    fn = _function(generate_fold(acc_type, val_types, expr))
    nseg = int(seg_offsets.numel()) - 1
    dev = val_cols[0].device
    out = torch.empty(max(nseg, 0), dtype=acc_type.torch_dtype, device=dev)
    if nseg <= 0:
        return out
    cols = [c.contiguous() for c in val_cols]
    seg = seg_offsets.to(device=dev, dtype=torch.int64).contiguous()
    args = [ctypes.c_longlong(nseg), ctypes.c_void_p(seg.data_ptr())] + \
           [ctypes.c_void_p(c.data_ptr()) for c in cols] + [ctypes.c_void_p(out.data_ptr())]
    argv = (ctypes.c_void_p * len(args))(*[ctypes.cast(ctypes.pointer(a), ctypes.c_void_p) for a in args])
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    blocks = min((nseg + 127) // 128, sms * 16)
    rc = _driver().cuLaunchKernel(fn, blocks, 1, 1, 128, 1, 1, 0,
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), argv, None)
    if rc != 0:
        raise RuntimeError("cuLaunchKernel failed: %d" % rc)
    launch_count += 1
    return out


def fold_host(acc_type: DType, val_cols: Sequence[torch.Tensor], seg_offsets: torch.Tensor,
              expr: str) -> torch.Tensor:
    """The same fold with torch ops under C semantics: round ``r`` folds the r-th
    row of every key at once (CPU executors and the reference for the kernel)."""
    nseg = int(seg_offsets.numel()) - 1
    dev = val_cols[0].device if val_cols else None
    acc = torch.zeros(max(nseg, 0), dtype=acc_type.torch_dtype, device=dev)
    if nseg <= 0:
        return acc
    off = seg_offsets.to(dev)
    lens = off[1:] - off[:-1]
    rewritten = re.sub(r"\bacc\b", "c0", expr)
    for r in range(int(lens.max())):
        live = torch.nonzero(lens > r).flatten()
        rows = off[:-1][live] + r
        v = eval_expr(rewritten, [acc[live]] + [c[rows] for c in val_cols])
        if v.is_floating_point() and not acc.dtype.is_floating_point:
            v = torch.trunc(v)
        acc[live] = v.to(acc.dtype)
    return acc


# ---------------------------------------------------------------------------
# host evaluation with C semantics: a small expression parser (C precedence,
# casts, && || ! ?:, truncating integer / and %) evaluated with torch ops.

_TOK = re.compile(r"\s*(?:(\d+\.\d*(?:[eE][-+]?\d+)?[fF]?|\.\d+[fF]?|\d+[eE][-+]?\d+[fF]?|\d+[uUlL]*)"
                  r"|([A-Za-z_][A-Za-z_0-9]*)|(<<|>>|<=|>=|==|!=|&&|\|\||[-+*/%<>&|^~!?:(),]))")
_CAST = {"int": torch.int32, "long": torch.int64, "float": torch.float32, "double": torch.float64,
         "bool": torch.bool, "short": torch.int16, "char": torch.int8, "unsigned": torch.int64,
         "signed": torch.int64}
_BIN = [("||",), ("&&",), ("|",), ("^",), ("&",), ("==", "!="), ("<", "<=", ">", ">="),
        ("<<", ">>"), ("+", "-"), ("*", "/", "%")]
_FUNCS = {"abs": torch.abs, "fabs": torch.abs, "sqrt": torch.sqrt, "floor": torch.floor,
          "ceil": torch.ceil, "exp": torch.exp, "log": torch.log,
          "min": torch.minimum, "max": torch.maximum, "fmin": torch.minimum, "fmax": torch.maximum,
          "bs_abs": torch.abs, "bs_min": torch.minimum, "bs_max": torch.maximum,
          "bs_pymod": torch.remainder, "bs_floordiv": lambda a, b: torch.div(a, b, rounding_mode="floor"),
          "fmod": torch.fmod, "pow": torch.pow}


def _tokens(expr: str):
    out, pos = [], 0
    expr = expr.strip()
    while pos < len(expr):
        m = _TOK.match(expr, pos)
        if not m or m.end() == pos:
            raise ValueError("cannot parse expression %r at %r" % (expr, expr[pos:]))
        num, ident, op = m.groups()
        out.append(("num", num) if num else ("id", ident) if ident else ("op", op))
        pos = m.end()
    return out


class _Parser:
    def __init__(self, expr: str, env: Dict[str, torch.Tensor], n: int, device):
        self.toks, self.i, self.env, self.n, self.device = _tokens(expr), 0, env, n, device

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else (None, None)

    def take(self, val=None):
        t = self.peek()
        if val is not None and t[1] != val:
            raise ValueError("expected %r, got %r" % (val, t[1]))
        self.i += 1
        return t

    def tensor(self, v):
        if isinstance(v, torch.Tensor):
            return v
        return torch.full((self.n,), v, device=self.device)

    def parse(self):
        v = self.ternary()
        if self.i != len(self.toks):
            raise ValueError("unexpected %r" % (self.peek()[1],))
        return v

    def ternary(self):
        c = self.binary(0)
        if self.peek()[1] == "?":
            self.take()
            a = self.ternary()
            self.take(":")
            b = self.ternary()
            return torch.where(self.tensor(c).bool(), self.tensor(a), self.tensor(b))
        return c

    def binary(self, level):
        if level == len(_BIN):
            return self.unary()
        left = self.binary(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in _BIN[level]:
            op = self.take()[1]
            right = self.binary(level + 1)
            left = self.apply(op, left, right)
        return left

    def apply(self, op, a, b):
        ta, tb = isinstance(a, torch.Tensor), isinstance(b, torch.Tensor)
        if not ta and not tb:
            a = self.tensor(a)
            ta = True
        isint = lambda x: (isinstance(x, torch.Tensor) and not x.is_floating_point()) or \
            (not isinstance(x, torch.Tensor) and isinstance(x, (int, bool)))  # noqa: E731
        if op == "||":
            return self.tensor(a).bool() | self.tensor(b).bool()
        if op == "&&":
            return self.tensor(a).bool() & self.tensor(b).bool()
        if op == "/":
            if isint(a) and isint(b):
                return torch.div(self.tensor(a), b, rounding_mode="trunc")
            return self.tensor(a) / b if ta else a / b
        if op == "%":
            return torch.fmod(self.tensor(a), b)
        if ta and a.dtype == torch.bool and op in "+-*<<>>":
            a = a.to(torch.int32)
        if tb and b.dtype == torch.bool and op in "+-*<<>>":
            b = b.to(torch.int32)
        import operator as o
        f = {"|": o.or_, "^": o.xor, "&": o.and_, "==": o.eq, "!=": o.ne, "<": o.lt, "<=": o.le,
             ">": o.gt, ">=": o.ge, "<<": o.lshift, ">>": o.rshift, "+": o.add, "-": o.sub,
             "*": o.mul}[op]
        return f(a, b)

    def unary(self):
        kind, val = self.peek()
        if kind == "op" and val in ("-", "+", "!", "~"):
            self.take()
            v = self.unary()
            if val == "-":
                return -v
            if val == "+":
                return v
            if val == "!":
                return ~self.tensor(v).bool()
            return ~self.tensor(v)
        if kind == "op" and val == "(":
            # cast or parenthesised expression
            j = self.i + 1
            words = []
            while j < len(self.toks) and self.toks[j][0] == "id" and self.toks[j][1] in _CAST:
                words.append(self.toks[j][1])
                j += 1
            if words and j < len(self.toks) and self.toks[j][1] == ")":
                self.i = j + 1
                v = self.unary()
                if words[:1] == ["unsigned"] and words[-1] == "char":
                    dt = torch.uint8
                else:
                    dt = torch.int64 if "long" in words else _CAST[words[-1] if words[-1] != "unsigned" else "long"]
                v = self.tensor(v)
                if not dt.is_floating_point and dt != torch.bool and v.is_floating_point():
                    v = torch.trunc(v)
                return v.to(dt)
            self.take("(")
            v = self.ternary()
            self.take(")")
            return v
        if kind == "num":
            self.take()
            txt = val.rstrip("uUlLfF")
            return float(txt) if any(c in txt for c in ".eE") else int(txt)
        if kind == "id":
            self.take()
            if self.peek()[1] == "(" and val in _FUNCS:
                self.take("(")
                args = [self.ternary()]
                while self.peek()[1] == ",":
                    self.take()
                    args.append(self.ternary())
                self.take(")")
                args = [self.tensor(a) for a in args]
                if val in ("sqrt", "exp", "log", "floor", "ceil") and not args[0].is_floating_point():
                    args[0] = args[0].to(torch.float64)
                return _FUNCS[val](*args)
            if val in ("true", "false"):
                return val == "true"
            if val not in self.env:
                raise ValueError("unknown identifier %r" % val)
            return self.env[val]
        raise ValueError("unexpected token %r" % (val,))


def eval_expr(expr: str, cols: Sequence[torch.Tensor]) -> torch.Tensor:
    n = int(cols[0].shape[0]) if cols else 0
    dev = cols[0].device if cols else None
    env = {"c%d" % j: c for j, c in enumerate(cols)}
    p = _Parser(expr, env, n, dev)
    return p.tensor(p.parse())


def eval_host(cols: Sequence[torch.Tensor], in_types: Sequence[DType], stages: Sequence[Stage]):
    """Evaluate the stage chain with torch ops (CPU or any device)."""
    cur = list(cols)
    mask = None
    for st in stages:
        if st.pred is not None:
            p = eval_expr(st.pred, cur).to(torch.bool)
            mask = p if mask is None else (mask & p)
        else:
            new = []
            for e, t in zip(st.exprs, st.out):
                v = eval_expr(e, cur)
                if v.is_floating_point() and not t.torch_dtype.is_floating_point and t.torch_dtype != torch.bool:
                    v = torch.trunc(v)
                new.append(v.to(t.torch_dtype))
            cur = new
    return cur, mask
