"""Slice combinators: the user-facing dataflow DSL.

A :class:`Slice` is a typed, sharded, lazily evaluated dataset.  Its kernel is
``reader(shard, deps) -> Reader``; slices with a single non-shuffle dependency
are pipelined into one task per shard by the compiler, shuffle dependencies
cut the pipeline (``exec/compile``).

All combinators work on whole column batches: ``Map``/``Filter``/``Flatmap``
call the user function once per batch with column tensors (see
:mod:`bigslice_b200.slicefunc`; per-row python functions via ``rowwise``),
which on a B200 means one fused elementwise/compaction pass per tile rather
than one reflective call per row.

Capability parity: reference ``slice.go:40-1173`` — Slice/Dep/ShardType/
Pragma(s)/Procs/Exclusive/ExperimentalMaterialize, Const, ReaderFunc,
WriterFunc, Map, Filter, Flatmap, Fold, Head, Scan, Prefixed, Unwrap, String,
Name/MakeName/Helper.
"""
from __future__ import annotations

import logging
import threading
import types as _pytypes
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import context as _context
from . import errors, slicefunc, slicetype, typecheck
from .frame import Frame
from .frame import columns as C
from .internal import defaultsize
from .ops.segmented import group_starts, segmented_fold
from .sliceio import EOF, FrameReader, FuncReader, Reader, Scanner, read_all
from .slicetype import DType, Type

log = logging.getLogger("bigslice")

# --------------------------------------------------------------------------
# naming


class Name:
    """Unique name of a slice: op, creation site and per-site index."""

    __slots__ = ("op", "file", "line", "index")

    def __init__(self, op: str, file: str, line: int, index: int):
        self.op, self.file, self.line, self.index = op, file, line, index

    def __str__(self) -> str:
        import os
        s = "%s@%s:%d" % (self.op, os.path.basename(self.file), self.line)
        if self.index:
            s += "(%d)" % self.index
        return s

    __repr__ = __str__
    Op = property(lambda self: self.op)


_name_lock = threading.Lock()
_name_counts: dict = {}


def make_name(op: str) -> Name:
    """Name for a new slice created at the user's call site (``MakeName``)."""
    file, line = typecheck.caller_location()
    with _name_lock:
        key = (op, file, line)
        idx = _name_counts.get(key, 0)
        _name_counts[key] = idx + 1
    return Name(op, file, line, idx)


MakeName = make_name


def helper(fn: Callable) -> Callable:
    """Decorator marking ``fn`` as a helper that constructs slices: names and
    type errors are attributed to its caller (``bigslice.Helper``)."""
    return typecheck.mark_helper(fn)


Helper = helper

# --------------------------------------------------------------------------
# pragmas


class Pragma:
    """Runtime directives attached to a slice (``bigslice.Pragma``)."""

    def procs(self) -> int:
        return 1

    def exclusive(self) -> bool:
        return False

    def materialize(self) -> bool:
        return False


class _Procs(Pragma):
    def __init__(self, n: int):
        self.n = n

    def procs(self):
        return self.n


class _Exclusive(Pragma):
    def exclusive(self):
        return True


class _Materialize(Pragma):
    def materialize(self):
        return True


def Procs(n: int) -> Pragma:
    """Reserve ``n`` procs for each task of the slice (``bigslice.Procs``)."""
    return _Procs(n)


Exclusive: Pragma = _Exclusive()
ExperimentalMaterialize: Pragma = _Materialize()


class Pragmas(Pragma, list):
    """Composition: max procs, any exclusive, any materialize."""

    def procs(self):
        return max([p.procs() for p in self] + [1])

    def exclusive(self):
        return any(p.exclusive() for p in self)

    def materialize(self):
        return any(p.materialize() for p in self)


# --------------------------------------------------------------------------
# core types

HashShard, RangeShard = 0, 1


class Dep:
    """A dependency edge (``bigslice.Dep``)."""

    __slots__ = ("slice", "shuffle", "partitioner", "expand")

    def __init__(self, slice_: "Slice", shuffle: bool = False,
                 partitioner: Optional[Callable] = None, expand: bool = False):
        self.slice = slice_
        self.shuffle = shuffle
        self.partitioner = partitioner
        self.expand = expand


class Slice(Pragma):
    """Base class of all slices.  Subclass and override ``reader`` (and
    ``num_dep``/``dep``) to define custom slices — the documented extension
    protocol."""

    def __init__(self, name: Name, typ: Type, num_shard: int,
                 deps: Sequence[Dep] = (), pragmas: Sequence[Pragma] = ()):
        self._name = name
        self.typ = typ
        self._num_shard = num_shard
        self._deps = list(deps)
        self._pragma = Pragmas(pragmas)

    # slicetype.Type
    def num_out(self) -> int:
        return self.typ.num_out()

    def out(self, i: int) -> DType:
        return self.typ.out(i)

    def prefix(self) -> int:
        return self.typ.prefix()

    # Slice
    def name(self) -> Name:
        return self._name

    def num_shard(self) -> int:
        return self._num_shard

    def shard_type(self) -> int:
        return HashShard

    def num_dep(self) -> int:
        return len(self._deps)

    def dep(self, i: int) -> Dep:
        return self._deps[i]

    def combiner(self) -> Optional[slicefunc.Func]:
        return None

    def reader(self, shard: int, deps: List[Reader]) -> Reader:
        raise NotImplementedError

    # Pragma
    def procs(self):
        return self._pragma.procs()

    def exclusive(self):
        return self._pragma.exclusive()

    def materialize(self):
        return self._pragma.materialize()

    def __str__(self):
        return string(self)

    NumOut, Out, Prefix, NumShard, ShardType, NumDep = \
        num_out, out, prefix, num_shard, shard_type, num_dep
    Combiner, Reader = combiner, reader


def string(s: Slice) -> str:
    """``op@file:line<types>`` (``bigslice.String``)."""
    return "%s%s" % (s.name(), slicetype.string_of(s.typ).replace("slice", "", 1))


String = string


def _fn_name(fn) -> str:
    return getattr(fn, "__qualname__", None) or type(fn).__name__


def _is_slice(s) -> bool:
    return isinstance(s, Slice)


def _check_slice(op: str, s):
    if not _is_slice(s):
        typecheck.panicf("%s: not a slice: %s", op, type(s).__name__)


# --------------------------------------------------------------------------
# Const


def const_shard(n: int, nshard: int, shard: int):
    """Row range [beg, end) of ``shard``: even contiguous split, sizes differ
    by at most one (``slice.go:263-277``)."""
    base, rem = divmod(n, nshard)
    beg = shard * base + min(shard, rem)
    end = beg + base + (1 if shard < rem else 0)
    return beg, end


class _Const(Slice):
    def __init__(self, nshard, frame: Frame):
        super().__init__(make_name("const"), frame.typ, nshard)
        self.frame = frame

    def reader(self, shard, deps):
        beg, end = const_shard(len(self.frame), self._num_shard, shard)
        f = self.frame.slice(beg, end)
        dev = _context.current_device()
        if dev is not None:
            f = f.to(dev)
        return FrameReader(f)


def Const(nshard: int, *columns) -> Slice:
    """A slice over in-memory columns split evenly across ``nshard`` shards."""
    if len(columns) == 0:
        typecheck.panic("const: must have at least one column")
    if nshard < 1:
        typecheck.panic("const: shard must be >= 1")
    try:
        frame = Frame.of(*columns)
    except (TypeError, ValueError):
        typecheck.panic("const: invalid slice inputs")
    return _Const(nshard, frame)


# --------------------------------------------------------------------------
# ReaderFunc


class _ReaderFunc(Slice):
    def __init__(self, nshard, fn, typ, pragmas, state_factory, batch):
        super().__init__(make_name("reader"), typ, nshard, (), pragmas)
        self.fn = fn
        self.func = slicefunc.Func(fn, None, typ)
        self.state_factory = state_factory
        self.batch = batch

    def reader(self, shard, deps):
        state = self.state_factory() if self.state_factory is not None \
            else _pytypes.SimpleNamespace()
        done = [False]
        consecutive_empty = [0]
        func, typ, batch_default = self.func, self.typ, self.batch

        def read(n):
            if done[0]:
                return None
            rows = n or batch_default or defaultsize.Chunk
            ctx = _context.current()
            buf = Frame.make(typ, rows)  # zeroed before every call
            cols = buf.cols()
            try:
                if func.takes_ctx:
                    r = func.fn(ctx, shard, state, *cols)
                else:
                    r = func.fn(shard, state, *cols)
            except (EOFError, StopIteration):
                done[0] = True
                return None
            except errors.Error as e:
                if errors.is_temporary(e):
                    raise
                raise errors.fatal(e)
            err = None
            if isinstance(r, tuple):
                cnt, err = r
            else:
                cnt = r
            cnt = int(cnt or 0)
            if err is not None and err is not EOF and not isinstance(err, type(EOF)):
                if isinstance(err, BaseException):
                    if errors.is_temporary(err):
                        raise err
                    raise errors.fatal(err)
            if err is not None:
                done[0] = True
            if cnt == 0 and err is None:
                consecutive_empty[0] += 1
                c = consecutive_empty[0]
                if c >= 8 and (c & (c - 1)) == 0:
                    log.warning("%s: %d consecutive empty reads; did you forget to "
                                "return sliceio.EOF?", self.name(), c)
            else:
                consecutive_empty[0] = 0
            out = buf.slice(0, cnt)
            dev = ctx.device
            if dev is not None:
                out = out.to(dev)
            if cnt == 0 and done[0]:
                return None
            return out

        return FuncReader(read)


def ReaderFunc(nshard: int, read: Callable, *pragmas, out=None, state=None,
               batch: Optional[int] = None) -> Slice:
    """A source slice produced by a user function.

    ``read(shard, state, *cols) -> n | (n, EOF)`` fills the zeroed column
    buffers ``cols`` (tensors / object arrays) with up to ``len(cols[0])``
    rows and returns how many it wrote; it signals the end of the shard by
    returning ``(n, EOF)`` or raising ``EOFError``.  ``state`` is a fresh
    per-shard object (``state=`` factory, default a ``SimpleNamespace``).
    Column types come from ``out=`` or from the annotations of the column
    parameters.
    """
    if not callable(read):
        typecheck.panicf("readerfunc: invalid reader function type %s", type(read).__name__)
    if nshard < 1:
        typecheck.panic("readerfunc: shard must be >= 1")
    if out is None:
        import inspect
        import typing
        try:
            params = list(inspect.signature(read).parameters.values())
            hints = typing.get_type_hints(read)
        except (TypeError, ValueError, NameError):
            params, hints = [], {}
        if params and (params[0].name == "ctx" or params[0].annotation is slicefunc.Context):
            params = params[1:]
        out = []
        for p in params[2:]:
            if p.default is not p.empty or p.kind not in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD):
                continue  # bound extras (closures via defaults) are not columns
            ann = hints.get(p.name, p.annotation)
            try:
                out.append(slicetype.dtype_of(ann))
            except TypeError:
                out = []
                break
        if not out:
            typecheck.panicf("readerfunc: function %s is not vectorized", _fn_name(read))
    typ = Type([slicetype.dtype_of(t) for t in out])
    if typ.num_out() == 0:
        typecheck.panicf("readerfunc: function %s is not vectorized", _fn_name(read))
    return _ReaderFunc(nshard, read, typ, pragmas, state, batch)


class _ReaderGen(Slice):
    def __init__(self, nshard, gen, typ, pragmas):
        super().__init__(make_name("reader"), typ, nshard, (), pragmas)
        self.gen = gen

    def reader(self, shard, deps):
        it = [None]
        typ = self.typ

        def read(n):
            if it[0] is None:
                it[0] = iter(self.gen(shard))
            try:
                item = next(it[0])
            except StopIteration:
                return None
            f = item if isinstance(item, Frame) else Frame.of(*item, typ=typ)
            dev = _context.current_device()
            return f.to(dev) if dev is not None else f

        from .sliceio import Rebatch
        return Rebatch(FuncReader(read))


def ReaderGen(nshard: int, gen: Callable, out, *pragmas) -> Slice:
    """Source slice from a generator: ``gen(shard)`` yields column tuples (or
    Frames); the idiomatic python form of ``ReaderFunc``."""
    if not callable(gen):
        typecheck.panicf("readerfunc: invalid reader function type %s", type(gen).__name__)
    typ = Type([slicetype.dtype_of(t) for t in out])
    return _ReaderGen(nshard, gen, typ, pragmas)


# --------------------------------------------------------------------------
# WriterFunc


class _WriterFunc(Slice):
    def __init__(self, s: Slice, fn, state_factory):
        super().__init__(make_name("writer"), s.typ, s.num_shard(), [Dep(s)], ())
        self.fn = fn
        self.func = slicefunc.Func(fn)
        self.state_factory = state_factory

    def reader(self, shard, deps):
        src = deps[0]
        state = self.state_factory() if self.state_factory is not None \
            else _pytypes.SimpleNamespace()
        func = self.func
        finished = [False]

        def call(err, cols):
            ctx = _context.current()
            try:
                if func.takes_ctx:
                    r = func.fn(ctx, shard, state, err, *cols)
                else:
                    r = func.fn(shard, state, err, *cols)
            except errors.Error as e:
                if errors.is_temporary(e):
                    raise
                raise errors.fatal(e)
            if isinstance(r, BaseException):
                if errors.is_temporary(r):
                    raise r
                raise errors.fatal(r)

        def read(n):
            if finished[0]:
                return None
            try:
                f = src.read(n)
            except BaseException as e:
                finished[0] = True
                call(e, Frame.empty(self.typ).cols())
                raise
            if f is None:
                finished[0] = True
                call(EOF, Frame.empty(self.typ).cols())
                return None
            call(None, f.cols())
            return f

        return FuncReader(read, src.close)


def WriterFunc(slice_: Slice, write: Callable, state=None) -> Slice:
    """Pass-through slice calling ``write(shard, state, err, *cols)`` for every
    batch read (``err`` is None, ``EOF`` at end of shard, or the read error).
    A returned/raised error fails the task (fatal unless temporary)."""
    _check_slice("writerfunc", slice_)
    if not callable(write):
        typecheck.panicf("writerfunc: invalid writer function type %s; %s",
                         type(write).__name__, "must be a function")
    f = slicefunc.Func(write)
    want = slice_.num_out() + 3
    if not f.is_variadic and f.nparams != want:
        typecheck.panicf("writerfunc: invalid writer function type %s; %s", _fn_name(write),
                         "must have %d arguments (shard, state, err, cols...)" % want)
    return _WriterFunc(slice_, write, state)


# --------------------------------------------------------------------------
# Map / Filter / Flatmap


def _resolve_out(op: str, func: slicefunc.Func, arg: Type, out, fn, flat=False) -> Type:
    if out is not None:
        if not isinstance(out, (list, tuple, Type)):
            out = [out]
        t = Type([slicetype.dtype_of(o) for o in out])
        func.Out = t
        return t
    if func.Out is not None:
        return func.Out
    t = func.infer_out(arg, flat)
    if t is None and not flat:
        from .ops import trace   # the zero-row probe failed: the symbolic trace also knows the types
        tys = trace.infer_out(func, arg.cols)
        if tys is not None:
            t = func.Out = Type(tys)
    if t is None:
        typecheck.panicf("%s: cannot infer output types of %s; annotate it or pass out=",
                         op, _fn_name(fn))
    return t


class _Map(Slice):
    def __init__(self, s, func, typ, pragmas):
        super().__init__(make_name("map"), typ, s.num_shard(), [Dep(s)], pragmas)
        self.func = func
        self.in_typ = s.typ
        self._stage = False   # not traced yet (None: not traceable)

    def _apply(self, cols, n, device):
        func, typ = self.func, self.typ
        out = func.call(_context.current(), cols)
        if len(out) != typ.num_out():
            raise errors.fatal("map: function returned %d columns, want %d"
                               % (len(out), typ.num_out()))
        return [_coerce(c, t, n, device) for c, t in zip(out, typ.cols)]

    def traced(self):
        """The function as a fused-kernel stage (``ops/trace.py``), or None."""
        if self._stage is False:
            from .ops import trace
            st = None
            if not trace.rejected(self.func.fn):
                st = trace.trace_map(self.func, self.in_typ.cols, self.typ.cols)
            if st is not None:
                st.eager, st.fn = self._apply, self.func.fn
            self._stage = st
        return self._stage

    def reader(self, shard, deps):
        src, typ = deps[0], self.typ
        st = self.traced()
        if st is not None:
            return _chain_stage(src, self.in_typ, st, typ)

        def read(n):
            f = src.read(n)
            if f is None:
                return None
            return Frame(self._apply(f.cols(), len(f), f.device), typ)

        return FuncReader(read, src.close)


def _coerce(c, t: DType, n: int, device):
    """Make a user-returned column conform to dtype ``t`` and length ``n``."""
    if isinstance(c, (C.ListColumn, C.StringColumn)):
        return c
    if t.is_tensor:
        if not isinstance(c, torch.Tensor):
            c = torch.as_tensor(np.asarray(c)) if not np.isscalar(c) else torch.tensor(c)
        if c.dim() == 0:
            c = c.expand(n)
        if c.dtype != t.torch_dtype:
            c = c.to(t.torch_dtype)
        if device is not None and c.device != device:
            c = c.to(device)
        return c
    if isinstance(c, np.ndarray) and c.dtype == object:
        return c
    a = np.empty(len(c), dtype=object)
    for i, v in enumerate(c):
        a[i] = v
    return a


def Map(slice_: Slice, fn: Callable, *pragmas, out=None) -> Slice:
    """Transform rows with ``fn`` (batch-vectorised; ``rowwise(fn)`` for a
    per-row function).  The key prefix of the result resets to 1."""
    _check_slice("map", slice_)
    if not callable(fn):
        typecheck.panicf("map: invalid map function %s", type(fn).__name__)
    func = slicefunc.Func(fn)
    if not func.accepts(slice_.num_out()) or \
            (func.In is not None and not typecheck.can_apply(func.In.cols, slice_.typ, func.is_variadic)):
        typecheck.panicf("map: function %s does not match input slice type %s",
                         _fn_name(fn), slicetype.string_of(slice_.typ))
    typ = _resolve_out("map", func, slice_.typ, out, fn)
    if typ.num_out() == 0:
        typecheck.panicf("map: need at least one output column")
    return _Map(slice_, func, Type(typ.cols, 1), pragmas)


class _Filter(Slice):
    def __init__(self, s, func, pragmas):
        super().__init__(make_name("filter"), s.typ, s.num_shard(), [Dep(s)], pragmas)
        self.func = func
        self._stage = False

    def _mask(self, cols, n, device):
        (m,) = self.func.call(_context.current(), cols)
        if not isinstance(m, torch.Tensor):
            m = torch.as_tensor(np.asarray(m, dtype=bool))
        return m.to(torch.bool)

    def traced(self):
        if self._stage is False:
            from .ops import trace
            st = None
            if not trace.rejected(self.func.fn):
                st = trace.trace_filter(self.func, self.typ.cols)
            if st is not None:
                st.eager, st.fn = self._mask, self.func.fn
            self._stage = st
        return self._stage

    def reader(self, shard, deps):
        src = deps[0]
        st = self.traced()
        if st is not None:
            return _chain_stage(src, self.typ, st, self.typ)

        def read(n):
            while True:
                f = src.read(n)
                if f is None:
                    return None
                m = self._mask(f.cols(), len(f), f.device)
                if m.is_cuda:
                    from .ops import gpu
                    if gpu.available():
                        idx = gpu.compact_indices(m)
                        out = f.take(idx)
                    else:
                        out = f.mask(m)
                else:
                    out = f.mask(m)
                if len(out) or n is None:
                    return out

        return FuncReader(read, src.close)


def Filter(slice_: Slice, pred: Callable, *pragmas) -> Slice:
    """Keep rows for which ``pred`` (returning one boolean column) is true."""
    _check_slice("filter", slice_)
    if not callable(pred):
        typecheck.panicf("filter: invalid predicate function %s", type(pred).__name__)
    func = slicefunc.Func(pred)
    if not func.accepts(slice_.num_out()) or \
            (func.In is not None and not typecheck.can_apply(func.In.cols, slice_.typ, func.is_variadic)):
        typecheck.panicf("filter: function %s does not match input slice type %s",
                         _fn_name(pred), slicetype.string_of(slice_.typ))
    out = func.Out or func.infer_out(slice_.typ)
    if out is None or out.num_out() != 1 or out.out(0) != slicetype.bool_:
        typecheck.panic("filter: predicate must return a single boolean value")
    return _Filter(slice_, func, pragmas)


class _Flatmap(Slice):
    def __init__(self, s, func, typ, pragmas):
        super().__init__(make_name("flatmap"), typ, s.num_shard(), [Dep(s)], pragmas)
        self.func = func

    def _apply(self, f: Frame) -> Frame:
        func, typ = self.func, self.typ
        ctx = _context.current()
        if func.rowwise:
            lists = [C.col_tolist(c) for c in f.cols()]
            outs: List[list] = [[] for _ in typ.cols]
            for i in range(len(f)):
                r = func.call_row(ctx, [l[i] for l in lists])
                if not isinstance(r, tuple):
                    r = (r,)
                k = None
                for o, v in zip(outs, r):
                    v = C.col_tolist(v) if C.is_column(v) else list(v)
                    if k is None:
                        k = len(v)
                    elif k != len(v):
                        raise errors.fatal("flatmap: function returned columns of unequal length")
                    o.extend(v)
            dev = f.device if f.device.type != "cpu" else None
            cols = [C.col_from_list(o, t, dev if t.is_tensor else None)
                    for o, t in zip(outs, typ.cols)]
            return Frame(cols, typ)
        cols = func.call(ctx, f.cols())
        n = C.col_len(cols[0]) if cols else 0
        cols = [_coerce(c, t, n, None) for c, t in zip(cols, typ.cols)]
        return Frame(cols, typ)

    def reader(self, shard, deps):
        src = deps[0]
        pending: List[Optional[Frame]] = [None]

        def read(n):
            # Output may exceed the requested batch; the remainder is carried
            # (reference carries overflow in f.out, slice.go:797-829).
            while True:
                f = pending[0]
                pending[0] = None
                if f is None:
                    g = src.read(n)
                    if g is None:
                        return None
                    f = self._apply(g)
                if n is not None and len(f) > n:
                    pending[0] = f.slice(n, len(f))
                    f = f.slice(0, n)
                if len(f) or n is None:
                    return f

        return FuncReader(read, src.close)


def Flatmap(slice_: Slice, fn: Callable, *pragmas, out=None) -> Slice:
    """Map each row to zero or more rows.  A vectorised ``fn`` returns output
    columns of any (equal) length for the batch; a ``rowwise`` ``fn`` returns
    one list per output column for each row."""
    _check_slice("flatmap", slice_)
    if not callable(fn):
        typecheck.panicf("flatmap: invalid flatmap function %s", type(fn).__name__)
    func = slicefunc.Func(fn)
    if not func.accepts(slice_.num_out()) or \
            (func.In is not None and not typecheck.can_apply(func.In.cols, slice_.typ, func.is_variadic)):
        typecheck.panicf("flatmap: flatmap function %s does not match input slice type %s",
                         _fn_name(fn), slicetype.string_of(slice_.typ))
    if out is None and func.Out is not None and func.rowwise:
        dv = typecheck.devectorize(func.Out)
        if dv is None:
            typecheck.panicf("flatmap: flatmap function %s is not vectorized", _fn_name(fn))
        func.Out = None
        out = list(dv.cols)
    typ = _resolve_out("flatmap", func, slice_.typ, out, fn, flat=True)
    if typ.num_out() == 0:
        typecheck.panicf("flatmap: flatmap function %s is not vectorized", _fn_name(fn))
    return _Flatmap(slice_, func, Type(typ.cols, 1), pragmas)


# --------------------------------------------------------------------------
# Fold


#: A vectorised python fold is applied in rounds (round r = the r-th row of every
#: key); beyond this many rounds — one key owning that many rows — it refuses and
#: points at the builtin functors / ``FoldExpr``, which have device kernels.
MAX_FOLD_ROUNDS = 4096


class _Fold(Slice):
    def __init__(self, s, func, typ, pragmas, expr: Optional[str] = None):
        super().__init__(make_name("fold"), typ, s.num_shard(), [Dep(s, shuffle=True)], pragmas)
        self.func = func
        self.expr = expr
        self.in_typ = s.typ
        self._traced = False

    def traced_expr(self) -> Optional[str]:
        """The python fold function as a device-kernel expression (``ops/trace.py``)."""
        if self._traced is False:
            from .ops import trace
            e = None
            if self.func is not None and self.func.fn is not None and not trace.rejected(self.func.fn):
                e = trace.trace_fold(self.func, self.typ.out(1), list(self.in_typ.cols[1:]))
                if e is not None:
                    from .ops import jit
                    try:
                        jit.generate_fold(self.typ.out(1), list(self.in_typ.cols[1:]), e)
                    except (TypeError, ValueError):
                        e = None
            self._traced = e
        return self._traced

    def _traced_fold(self, acc_t, vals, starts, off) -> Optional[torch.Tensor]:
        """Device fold with the traced expression; the first use of a function is
        cross-checked on a bounded prefix against the function itself."""
        from .ops import jit, trace
        expr, fn = self.traced_expr(), self.func.fn
        vt = list(self.in_typ.cols[1:])
        if not trace.verified(fn):
            m = min(int(vals[0].shape[0]), 1024)
            soff = torch.cat([off[off < m], torch.tensor([m], dtype=torch.int64, device=off.device)])
            got = jit.run_fold(acc_t, [v[:m] for v in vals], vt, soff, expr)
            want = segmented_fold(acc_t, [v[:m] for v in vals], starts[:m], self.func, _context.current())
            if not _same_columns([got], [want]):
                trace.reject(fn)
                self._traced = None
                log.warning("fused form of %s disagrees with the function; running it eagerly", _fn_name(fn))
                return None
            trace.mark_verified(fn)
        return jit.run_fold(acc_t, vals, vt, off, expr)

    def _fold_frame(self, f: Frame) -> Frame:
        """Fold one materialised shard.  Three tiers:
        builtin functor (add/min/max) on a device frame -> the hash-aggregation
        kernels (no sort at all); expression -> stable sort + the NVRTC segmented
        fold kernel; python callable -> sort + vectorised rounds / per-row loop."""
        acc_t = self.typ.out(1)
        dev = f.device
        name = self.func.builtin if self.func is not None else None
        vals = f.cols()[1:]
        if (name in ("add", "min", "max") and len(vals) == 1 and dev.type == "cuda"
                and isinstance(vals[0], torch.Tensor) and vals[0].dim() == 1
                and vals[0].dtype == acc_t.torch_dtype):
            from .ops import gpu, keypack
            k = f.col(0)
            lay = keypack.layout_for(Type([self.typ.out(0)], 1))   # any key that fits one 64-bit word
            if gpu.available() and isinstance(k, torch.Tensor) and k.dim() == 1 and lay is not None \
                    and vals[0].dtype in gpu.VTYPE:
                uk, acc = gpu.hash_combine(k if lay.plain_int64 else keypack.pack([k], lay), vals[0], name)
                if not lay.plain_int64:
                    uk = keypack.unpack(uk, lay)[0]
                if name == "min":      # the accumulator starts at zero
                    acc = torch.clamp(acc, max=0)
                elif name == "max":
                    acc = torch.clamp(acc, min=0)
                order = Frame([uk], Type([self.typ.out(0)], 1)).sort_indices()
                return Frame([uk[order], acc[order]], self.typ)
        f = f.prefixed(1).sorted()   # stable: per-key row order is preserved
        starts = group_starts(f, 1)
        idx = torch.nonzero(starts).flatten()
        keys = C.col_take(f.col(0), idx)
        vals = f.cols()[1:]
        if self.expr is not None:
            from .ops import jit
            off = torch.cat([idx.to(torch.int64), torch.tensor([len(f)], dtype=torch.int64, device=idx.device)])
            vt = list(self.in_typ.cols[1:])
            if dev.type == "cuda":
                acc = jit.run_fold(acc_t, vals, vt, off, self.expr)
            else:
                acc = jit.fold_host(acc_t, vals, off, self.expr)
            return Frame([keys, acc], self.typ)
        if dev.type == "cuda" and _gpu_ok() and len(idx) and self.traced_expr() is not None:
            off = torch.cat([idx.to(torch.int64), torch.tensor([len(f)], dtype=torch.int64, device=idx.device)])
            acc = self._traced_fold(acc_t, vals, starts, off)
            if acc is not None:
                return Frame([keys, acc], self.typ)
        if not self.func.rowwise and isinstance(vals[0], torch.Tensor) and len(idx):
            seglen = torch.diff(torch.cat([idx, torch.tensor([len(f)], device=idx.device)]))
            longest = int(seglen.max())
            if longest > MAX_FOLD_ROUNDS and self.func.builtin not in ("add",):
                raise errors.fatal(
                    "fold: one key has %d rows; a python fold function is applied in one vectorised "
                    "round per row position (limit %d).  Use a builtin functor (ops.add/min/max) or "
                    "FoldExpr(slice, \"<C expression of acc, c1..>\"), which run as device kernels"
                    % (longest, MAX_FOLD_ROUNDS))
        acc = segmented_fold(acc_t, vals, starts, self.func, _context.current())
        return Frame([keys, acc], self.typ)

    def reader(self, shard, deps):
        src = deps[0]
        done = [False]

        def read(n):
            if done[0]:
                return None
            done[0] = True
            # Fold materialises the shard (reference slice.go:918-933).
            f = read_all(src, self.in_typ, _context.current_device())
            if len(f) == 0:
                return None
            return self._fold_frame(f)

        from .sliceio import Rebatch
        return Rebatch(FuncReader(read, src.close))


def Fold(slice_: Slice, fold: Callable, *pragmas, acc=None) -> Slice:
    """Per-key ordered fold: ``acc = fold(acc, v2, ..., vn)`` over the rows of
    each key (first column); the accumulator starts at its zero value.  The
    result is ``<key, acc>``.  ``acc=`` names the accumulator dtype when it
    cannot be inferred from annotations."""
    _check_slice("fold", slice_)
    n = slice_.num_out()
    if n < 2:
        typecheck.panicf("Fold can be applied only for slices with at least two columns; got %d", n)
    key = slice_.out(0)
    if not key.hashable:
        typecheck.panicf("fold: key type %s is not partitionable", key)
    if not key.comparable:
        typecheck.panicf("fold: key type %s cannot be accumulated", key)
    if not callable(fold):
        typecheck.panicf("fold: invalid fold function %s", type(fold).__name__)
    func = slicefunc.Func(fold)
    if not func.accepts(n):
        typecheck.panicf("fold: expected func(acc, t2, t3, ..., tn), got %s", _fn_name(fold))
    if acc is not None:
        acc_t = slicetype.dtype_of(acc)
    elif func.Out is not None:
        if func.Out.num_out() != 1:
            typecheck.panicf("fold: fold functions must return exactly one value")
        acc_t = func.Out.out(0)
    elif func.In is not None:
        acc_t = func.In.out(0)
    else:
        # default: accumulate in the type of the first value column
        acc_t = slice_.out(1)
    func.Out = Type([acc_t])
    if func.In is not None and not typecheck.can_apply(
            func.In.cols, Type([acc_t] + list(slice_.typ.cols[1:])), func.is_variadic):
        typecheck.panicf("fold: expected func(acc, t2, t3, ..., tn), got %s", _fn_name(fold))
    if func.builtin is None and n == 2 and slice_.out(1) == acc_t and not func.rowwise and not func.takes_ctx:
        # ``lambda acc, v: acc + v`` is the builtin add (ops/trace.py): hash-aggregation kernels
        from .ops import trace
        func.builtin = trace.recognise_combiner(fold, acc_t)
        if func.builtin == "first":    # fold(acc, v) = acc is not the combiner "first"
            func.builtin = None
    return _Fold(slice_, func, Type([key, acc_t], 1), pragmas)


def FoldExpr(slice_: Slice, expr: str, acc, *pragmas) -> Slice:
    """Per-key ordered fold with a C expression: ``acc = <expr>`` over the rows of
    each key in row order, ``acc`` starting at zero; the expression may use ``acc``
    and the value columns ``c1 .. cn`` (``c0`` is the key).  ``acc`` names the
    accumulator dtype.  Non-associative expressions are fine (e.g. a polynomial
    hash ``acc * 31 + c1``): rows are stably sorted by key and one device thread
    folds one key (NVRTC-compiled for sm_100a, ``ops/jit.py``)."""
    from .ops import jit
    _check_slice("foldexpr", slice_)
    n = slice_.num_out()
    if n < 2:
        typecheck.panicf("Fold can be applied only for slices with at least two columns; got %d", n)
    key = slice_.out(0)
    if not key.hashable or not key.comparable:
        typecheck.panicf("foldexpr: key type %s is not partitionable", key)
    acc_t = slicetype.dtype_of(acc)
    try:
        jit.generate_fold(acc_t, list(slice_.typ.cols[1:]), expr)
    except (TypeError, ValueError) as e:
        typecheck.panicf("foldexpr: %s", e)
    return _Fold(slice_, None, Type([key, acc_t], 1), pragmas, expr=expr)


# --------------------------------------------------------------------------
# Head / Scan / Prefixed / Unwrap


class _Head(Slice):
    def __init__(self, s, n):
        super().__init__(make_name("head(%d)" % n), s.typ, s.num_shard(), [Dep(s)], ())
        self.n = n

    def reader(self, shard, deps):
        src = deps[0]
        left = [self.n]

        def read(n):
            if left[0] <= 0:
                return None
            want = left[0] if n is None else min(n, left[0])
            f = src.read(want)
            if f is None:
                return None
            if len(f) > left[0]:
                f = f.slice(0, left[0])
            left[0] -= len(f)
            return f

        return FuncReader(read, src.close)


def Head(slice_: Slice, n: int) -> Slice:
    """The first ``n`` rows of *each shard*."""
    _check_slice("head", slice_)
    return _Head(slice_, n)


class _Scan(Slice):
    def __init__(self, s, fn):
        super().__init__(make_name("scan"), Type([], 0), s.num_shard(), [Dep(s)], ())
        self.fn = fn
        self.in_typ = s.typ

    def reader(self, shard, deps):
        src = deps[0]
        done = [False]

        def read(n):
            if done[0]:
                return None
            done[0] = True
            sc = Scanner(src, self.in_typ)
            try:
                err = self.fn(shard, sc)
                if isinstance(err, BaseException):
                    raise err
                if sc.err() is not None:
                    raise sc.err()
            finally:
                sc.close()
            return None

        return FuncReader(read)


def Scan(slice_: Slice, scan: Callable) -> Slice:
    """Sink: run ``scan(shard, scanner)`` once per shard; the result has no
    columns."""
    _check_slice("scan", slice_)
    if not callable(scan):
        typecheck.panicf("scan: invalid scan function %s", type(scan).__name__)
    return _Scan(slice_, scan)


class _Prefixed(Slice):
    """Wrapper changing the key prefix; transparent to naming/pipelining."""

    def __init__(self, s: Slice, prefix: int):
        Slice.__init__(self, s.name(), Type(s.typ.cols, prefix), s.num_shard())
        self.inner = s

    def name(self):
        return self.inner.name()

    def num_shard(self):
        return self.inner.num_shard()

    def shard_type(self):
        return self.inner.shard_type()

    def num_dep(self):
        return self.inner.num_dep()

    def dep(self, i):
        return self.inner.dep(i)

    def combiner(self):
        return self.inner.combiner()

    def reader(self, shard, deps):
        r = self.inner.reader(shard, deps)
        prefix = self.typ.prefix()

        def read(n):
            f = r.read(n)
            return None if f is None else f.prefixed(prefix)

        return FuncReader(read, r.close)

    def procs(self):
        return self.inner.procs()

    def exclusive(self):
        return self.inner.exclusive()

    def materialize(self):
        return self.inner.materialize()


def Prefixed(slice_: Slice, prefix: int) -> Slice:
    """The same slice with the first ``prefix`` columns as its key."""
    _check_slice("prefixed", slice_)
    if prefix < 1:
        typecheck.panic("prefixed: prefix must include at least one column")
    if prefix > slice_.num_out():
        typecheck.panicf("prefixed: prefix %d is greater than number of columns %d",
                         prefix, slice_.num_out())
    return _Prefixed(slice_, prefix)


def Unwrap(slice_: Slice) -> Slice:
    """Strip ``Prefixed`` wrappers (``bigslice.Unwrap``)."""
    while isinstance(slice_, _Prefixed):
        slice_ = slice_.inner
    return slice_


# --------------------------------------------------------------------------
# default partitioner


def default_partitioner(frame: Frame, nshard: int) -> torch.Tensor:
    """``shard[i] = murmur3(key prefix of row i) % nshard`` (int64 tensor).

    Bit-exact with the reference placement (``exec/compile.go:20-24``)."""
    return frame.hash(0) % nshard


# --------------------------------------------------------------------------
# Expression stages (NVRTC-fused Map/Filter)


class _ExprReader(Reader):
    """Evaluates a chain of expression stages over each tile in one fused
    kernel (device) or stage by stage on the host: traced python stages run their
    original function there, string-expression stages are interpreted with torch
    ops under C semantics."""

    def __init__(self, src: Reader, in_typ: Type, stages: list, out_typ: Type):
        self.src, self.in_typ, self.stages, self.out_typ = src, in_typ, stages, out_typ
        self.started = False
        self.eager_only = False

    def _host(self, cols, n, device):
        """Stage-by-stage evaluation; rows are dropped at every Filter, so a later
        stage never sees a row an earlier predicate rejected."""
        from .ops import jit
        cur, types = list(cols), list(self.in_typ.cols)
        for st in self.stages:
            if st.eager is not None:
                r = st.eager(cur, n, device)
                mask = r if st.pred is not None else None
                cur = cur if st.pred is not None else r
            else:
                cur, mask = jit.eval_host(cur, types, [st])
            if st.pred is None:
                types = list(st.out)
            elif mask is not None:
                if mask.is_cuda and _gpu_ok():
                    from .ops import gpu
                    idx = gpu.compact_indices(mask.view(torch.bool) if mask.dtype == torch.uint8 else mask)
                    cur = [C.col_take(c, idx) for c in cur]
                    n = int(idx.numel())
                else:
                    cur = [c[mask] for c in cur]
                    n = int(cur[0].shape[0]) if cur else 0
        return cur

    def _device(self, cols, n, device):
        from .ops import jit, trace
        filtered = any(st.pred is not None for st in self.stages)
        try:
            if filtered:
                # count -> prefix sum -> write: only the kept rows are ever written
                outs = jit.run_fused_compact(cols, self.in_typ.cols, self.stages)
            else:
                outs, _ = jit.run_fused(cols, self.in_typ.cols, self.stages)
        except RuntimeError as e:
            traced = [st for st in self.stages if st.fn is not None]
            if not traced or "does not compile" not in str(e):
                raise
            # a traced function produced source NVRTC rejects: the function still runs eagerly
            for st in traced:
                trace.reject(st.fn)
            log.warning("fused form of %s does not compile (%s); running it eagerly",
                        ", ".join(_fn_name(st.fn) for st in traced), str(e).splitlines()[0])
            self.eager_only = True
            return self._host(cols, n, device)
        unchecked = [st for st in self.stages if st.fn is not None and not trace.verified(st.fn)]
        if unchecked:
            # first tile of a traced python function: its fused result must equal what the
            # function itself computes (a bounded prefix is enough to catch an impure or
            # mis-traced function; it then runs eagerly from here on)
            m = min(n, 4096)
            ref = self._host([c[:m] for c in cols], m, device)
            k = int(ref[0].shape[0]) if filtered else m   # kept rows among the first m
            if _same_columns(ref, [o[:k] for o in outs]):
                for st in unchecked:
                    trace.mark_verified(st.fn)
            else:
                for st in unchecked:
                    trace.reject(st.fn)
                log.warning("fused form of %s disagrees with the function; running it eagerly",
                            ", ".join(_fn_name(st.fn) for st in unchecked))
                self.eager_only = True
                return self._host(cols, n, device)
        return outs

    def read(self, n=None):
        self.started = True
        while True:
            f = self.src.read(n)
            if f is None:
                return None
            cols = f.cols()
            on_gpu = f.device.type == "cuda" and _gpu_ok() and len(f) > 0 and not self.eager_only
            outs = self._device(cols, len(f), f.device) if on_gpu else self._host(cols, len(f), f.device)
            out = Frame(list(outs), self.out_typ)
            if len(out) or n is None:
                return out

    def close(self):
        self.src.close()


def _gpu_ok() -> bool:
    from .ops import gpu
    return gpu.available()


def _same_columns(a, b) -> bool:
    if len(a) != len(b):
        return False
    for x, y in zip(a, b):
        if x.shape != y.shape or x.dtype != y.dtype:
            return False
        if x.is_floating_point():
            if not torch.allclose(x, y, rtol=1e-4, atol=1e-6, equal_nan=True):
                return False
        elif not torch.equal(x, y):
            return False
    return True


def _chain_stage(src: Reader, in_typ: Type, stage, out_typ: Type) -> Reader:
    """Append ``stage`` to the fused chain feeding ``src`` (one kernel for the
    whole chain) or start a new chain."""
    from .sliceio import LabelReader
    inner = src.reader if isinstance(src, LabelReader) else src
    if isinstance(inner, _ExprReader) and not inner.started:
        return _ExprReader(inner.src, inner.in_typ, inner.stages + [stage], out_typ)
    return _ExprReader(src, in_typ, [stage], out_typ)


class _ExprSlice(Slice):
    def __init__(self, op: str, s: Slice, stage, typ: Type, pragmas):
        super().__init__(make_name(op), typ, s.num_shard(), [Dep(s)], pragmas)
        self.stage = stage
        self.in_typ = s.typ

    def reader(self, shard, deps):
        return _chain_stage(deps[0], self.in_typ, self.stage, self.typ)


def _check_expr_cols(op: str, s: Slice, exprs):
    from .ops import jit
    try:
        jit.check_types(s.typ.cols)
    except TypeError as e:
        typecheck.panicf("%s: %s", op, e)
    for e in exprs:
        if not isinstance(e, str):
            typecheck.panicf("%s: expression must be a string, got %s", op, type(e).__name__)
        for m in jit._VAR.finditer(e):
            if int(m.group(1)) >= s.num_out():
                typecheck.panicf("%s: expression %r refers to c%s but the slice has %d columns",
                                 op, e, m.group(1), s.num_out())


def MapExpr(slice_: Slice, exprs: Sequence[str], out: Sequence, *pragmas) -> Slice:
    """Map with C expressions over the input columns ``c0, c1, ...`` — compiled
    with NVRTC and fused with adjacent expression stages into one sm_100a
    kernel (see :mod:`bigslice_b200.ops.jit`)."""
    from .ops import jit
    _check_slice("mapexpr", slice_)
    exprs = list(exprs)
    outs = [slicetype.dtype_of(o) for o in out]
    if len(exprs) == 0:
        typecheck.panicf("map: need at least one output column")
    if len(exprs) != len(outs):
        typecheck.panicf("mapexpr: %d expressions but %d output types", len(exprs), len(outs))
    _check_expr_cols("mapexpr", slice_, exprs)
    try:
        jit.check_types(outs)
    except TypeError as e:
        typecheck.panicf("mapexpr: %s", e)
    return _ExprSlice("mapexpr", slice_, jit.Stage(exprs=exprs, out=outs), Type(outs, 1), pragmas)


def FilterExpr(slice_: Slice, pred: str, *pragmas) -> Slice:
    """Filter with a C boolean expression over ``c0, c1, ...`` (fused like
    :func:`MapExpr`)."""
    from .ops import jit
    _check_slice("filterexpr", slice_)
    _check_expr_cols("filterexpr", slice_, [pred])
    return _ExprSlice("filterexpr", slice_, jit.Stage(pred=pred), slice_.typ, pragmas)
