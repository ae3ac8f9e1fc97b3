"""Reduce: keyed associative+commutative reduction with map-side combining.

``Reduce(slice, fn)`` reduces the single non-key column of rows sharing a key
prefix with ``fn(a, b) -> c``.  The dependency is a combining shuffle: every
producer pre-aggregates ("map-side combine") before the exchange and the
consumer merges the per-producer results.

On the B200 executor the built-in functors (``ops.add/min/max/...``) lower to
the fused combine→partition→exchange kernels (hot keys absorbed in a shared-
memory table, the rest aggregated in an HBM hash table, distinct keys pushed
to their destination GPU over NVLink and combined on arrival); arbitrary
vectorised functions use sort + segmented tree reduction.

Capability parity: reference ``reduce.go:42-94``.
"""
from __future__ import annotations

from typing import Callable, List

from . import context as _context
from . import slicefunc, slicetype, sortio, typecheck
from .frame import ops as frame_ops
from .sliceio import Reader
from .slices import Dep, Slice, _check_slice, _fn_name, make_name
from .slicetype import Type


def can_make_combining_frame(typ: Type) -> str:
    """Empty string if key columns are hashable and comparable, else why not
    (``canMakeCombiningFrame``, reference ``exec``/``reduce.go:45``)."""
    for i in range(typ.prefix()):
        t = typ.out(i)
        if not frame_ops.can_hash(t):
            return "key column(%d) type %s cannot be hashed" % (i, t)
        if not frame_ops.can_compare(t):
            return "key column(%d) type %s cannot be sorted" % (i, t)
    return ""


class _Reduce(Slice):
    def __init__(self, s: Slice, func: slicefunc.Func):
        super().__init__(make_name("reduce"), s.typ, s.num_shard(),
                         [Dep(s, shuffle=True, expand=True)], ())
        self.func = func

    def combiner(self):
        return self.func

    def reader(self, shard, deps: List[Reader]) -> Reader:
        if len(deps) == 1 and getattr(deps[0], "combined_unique", False):
            return deps[0]
        return sortio.reduce_reader(self.typ, deps, self.func, _context.current())


def Reduce(slice_: Slice, reduce: Callable) -> Slice:
    """Reduce values by key with an associative, commutative ``reduce``."""
    _check_slice("reduce", slice_)
    res = slice_.num_out() - slice_.prefix()
    if res != 1:
        typecheck.panicf("the slice must only have one 1 residual column; has %d", res)
    why = can_make_combining_frame(slice_.typ)
    if why:
        typecheck.panic(why)
    if not callable(reduce):
        typecheck.panicf("reduce: invalid reduce function %s", type(reduce).__name__)
    func = slicefunc.Func(reduce)
    vt = slice_.out(slice_.num_out() - 1)
    want = Type([vt, vt])
    bad = not func.accepts(2)
    if func.In is not None and not typecheck.can_apply(func.In.cols, want, func.is_variadic):
        bad = True
    if func.Out is not None and not (func.Out.num_out() == 1 and func.Out.out(0) == vt):
        bad = True
    if bad:
        typecheck.panicf("reduce: invalid reduce function %s, expected func(%s, %s) %s",
                         _fn_name(reduce), vt, vt, vt)
    func.Out = Type([vt])
    if func.builtin is None and not func.rowwise and not func.takes_ctx:
        # ``lambda a, b: a + b`` is the builtin ``add``: recognise it so python combiners take
        # the fused hash-combine/exchange lowering too (ops/trace.py)
        from .ops import trace
        func.builtin = trace.recognise_combiner(reduce, vt)
    return _Reduce(slice_, func)
