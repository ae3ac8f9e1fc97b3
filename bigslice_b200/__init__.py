"""bigslice_b200 — a Blackwell-native sharded dataflow engine.

Typed, sharded ``Slice`` combinators (Const, ReaderFunc, Map, Filter, Flatmap,
Fold, Reduce, Cogroup, Reshuffle, Reshard, Scan, Cache, ...) evaluated by a
loss-tolerant task scheduler on CPU tensors, on one B200, or SPMD across the
GPUs of an NVSwitch box where shuffles are hand-written sm_100a kernels that
push rows straight into peer HBM.

    import bigslice_b200 as bs

    @bs.Func
    def wordcount(nshard, keys, ones):
        s = bs.Const(nshard, keys, ones)
        return bs.Reduce(s, bs.ops.add)

    with bs.start(bs.Local, bs.Parallelism(8)) as sess:
        for key, count in sess.run(wordcount, 8, keys, ones).scanner():
            ...
"""
from . import errors, slicetype, slicefunc, typecheck, metrics, stats  # noqa: F401
from .slicetype import (List, Type, Vec, bool_, bytes_, float16, bfloat16, float32,
                        float64, int8, int16, int32, int64, obj, string, uint8,
                        uint16, uint32, uint64)
from .slicefunc import Context, rowwise, vectorized
from .frame import Frame, ListColumn
from .sliceio import EOF, Reader, Scanner
from .slices import (Const, Dep, Exclusive, ExperimentalMaterialize, Filter, Flatmap,
                     Fold, HashShard, Head, Helper, Map, Name, Pragma, Pragmas,
                     Prefixed, Procs, RangeShard, ReaderFunc, ReaderGen, Scan, Slice,
                     String, Unwrap, WriterFunc, helper, make_name, MakeName, MapExpr, FilterExpr, FoldExpr)
from .reduce import Reduce
from .cogroup import Cogroup, BroadcastJoin
from .reshuffle import Repartition, Reshard, Reshuffle
from .scan import ScanReader
from .cache import Cache, CachePartial, ReadCache
from .func import Func, FuncValue, Invocation
from .ops import functors as ops
from . import exec  # noqa: A004 - mirrors the reference package name
from . import slicetest, sliceio, sortio, frame  # noqa: F401,E402
from .exec import (Local, LocalDevice, MachineCombiners, MaxLoad, Parallelism, Result, StoreOpt, Journal,
                   Session, Spmd, TracePath, start, Start, StatusOpt, EventerOpt, Bigmachine)
from . import sliceconfig, text  # noqa: F401,E402

__version__ = "0.1.0"
