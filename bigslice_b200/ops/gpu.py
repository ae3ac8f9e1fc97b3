"""ctypes bindings for the sm_100a kernels in ``bigslice_b200/_bscuda.so``.

Every wrapper launches on the current torch CUDA stream and takes/returns
torch tensors; nothing here falls back to PyTorch silently — on a machine with
a GPU the library must load (``require()`` raises otherwise), so a passing GPU
test always exercised the hand-written kernels.

What the kernels replace in the reference: ``frame/frame.go:389-401``
+ ``frame/ops_builtin.go:143-164`` (hashing), ``exec/local.go:217-230`` (row-wise
partition copies), ``exec/combiner.go:62-357`` + ``exec/bigmachine.go:1144-1209``
(combining frame / map-side combine), ``sortio/sort.go:52`` (in-memory sort),
``slice.go:700-720`` (filter compaction), ``sortio/reader.go:88-127`` (merge-reduce),
``exec/topn.go:39-57`` (top-n).
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import List, Optional, Tuple

import torch

from ..frame import Frame
from ..frame import columns as C

_lib = None
_lock = threading.Lock()
_load_error: Optional[str] = None

c_void_p, c_int, c_ll, c_ull = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong

VTYPE = {torch.int32: 0, torch.int64: 1, torch.float32: 2, torch.float64: 3}
OPS = {"add": 0, "min": 1, "max": 2}
HASHABLE = (torch.int8, torch.int16, torch.int32, torch.int64, torch.uint8, torch.float32,
            torch.float64, torch.uint16, torch.uint32, torch.uint64)
SEGREDUCE_DTYPES = tuple(VTYPE)


def lib_path() -> str:
    if os.environ.get("BIGSLICE_CUDA_LIB"):  # debugging aid: load an alternative build
        return os.environ["BIGSLICE_CUDA_LIB"]
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_bscuda.so")


def _load():
    global _lib, _load_error
    with _lock:
        if _lib is not None or _load_error is not None:
            return _lib
        path = lib_path()
        if not os.path.exists(path) and not os.environ.get("BIGSLICE_CUDA_LIB"):
            # a fresh checkout (the .so is not tracked): build in-tree once
            try:
                from .. import build
                build.build_cuda()
            except Exception as e:  # noqa: BLE001 - reported through _load_error
                _load_error = "%s not built and the in-tree build failed: %s" % (path, e)
                return None
        if not os.path.exists(path):
            _load_error = "%s not built (run python -m bigslice_b200.build)" % path
            return None
        try:
            lib = ctypes.CDLL(path)
        except OSError as e:
            _load_error = str(e)
            return None
        lib.bs_launches.restype = c_ll
        lib.bs_sort_scratch_bytes.restype = c_ll
        lib.bs_compact_scratch_bytes.restype = c_ll
        lib.bs_crc32_nblocks.restype = c_ll
        _lib = lib
        return _lib


def available() -> bool:
    """True when CUDA is usable and the kernel library is loaded."""
    if not torch.cuda.is_available():
        return False
    if _load() is None:
        raise RuntimeError("bigslice_b200: CUDA device present but kernels missing: %s" % _load_error)
    return True


def require():
    if not available():
        raise RuntimeError("bigslice_b200.ops.gpu: no CUDA device / kernel library")
    return _lib


def launches() -> int:
    """Number of kernels this library has launched in this process."""
    lib = _load()
    return int(lib.bs_launches()) if lib is not None else 0


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def _chk(rc: int, what: str):
    if rc != 0:
        raise RuntimeError("bigslice_b200 kernel %s failed: cuda error %d" % (what, rc))


# ---------------------------------------------------------------- hashing

def hash_column(col: torch.Tensor, seed: int = 0, out: Optional[torch.Tensor] = None,
                xor_into: bool = False) -> torch.Tensor:
    """murmur3 of every element (int64 tensor of uint32 values)."""
    lib = require()
    col = col.contiguous()
    n = col.numel()
    if out is None:
        out = torch.empty(n, dtype=torch.int64, device=col.device)
    es = col.element_size()
    signed = col.dtype in (torch.int8, torch.int16)
    _chk(lib.bs_hash_column(_ptr(col), c_int(es), c_int(1 if signed else 0), c_ll(n),
                            ctypes.c_uint(seed & 0xFFFFFFFF), _ptr(out),
                            c_int(1 if xor_into else 0), _stream()), "hash_column")
    return out


# ---------------------------------------------------------------- partition

def partition_perm(ids: Optional[torch.Tensor], keys: Optional[torch.Tensor], n: int,
                   npart: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Permutation grouping rows by destination, and per-destination counts.

    Either explicit ``ids`` (int64) or int64 ``keys`` hashed with the default
    partitioner inside the kernel.  Row order inside a partition is unspecified
    (as across any shuffle)."""
    lib = require()
    dev = (ids if ids is not None else keys).device
    counts = torch.zeros(npart, dtype=torch.int64, device=dev)
    cursors = torch.empty(npart, dtype=torch.int64, device=dev)
    perm = torch.empty(n, dtype=torch.int64, device=dev)
    _chk(lib.bs_partition_perm(_ptr(ids), _ptr(keys), c_ll(n), c_int(npart), _ptr(counts),
                               _ptr(cursors), _ptr(perm), _stream()), "partition_perm")
    return perm, counts


def gather_rows(col: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """out[i] = col[idx[i]] with the hand-written gather kernel."""
    lib = require()
    col = col.contiguous()
    n = idx.numel()
    out = torch.empty((n,) + tuple(col.shape[1:]), dtype=col.dtype, device=col.device)
    row_bytes = col.element_size() * (col.numel() // col.shape[0] if col.shape[0] else 1)
    _chk(lib.bs_gather_rows(_ptr(col), c_int(row_bytes), _ptr(idx), c_ll(n), _ptr(out), _stream()),
         "gather_rows")
    return out


def take_frame(f: Frame, idx: torch.Tensor) -> Frame:
    # col_take picks the gather kernel for wide rows and the library for scalars
    return Frame([C.col_take(c, idx) for c in f.cols()], f.typ)


def partition_frame(f: Frame, ids: torch.Tensor, npart: int) -> List[Frame]:
    """Split a device frame into ``npart`` frames by destination id."""
    n = len(f)
    ids = ids.to(device=f.device, dtype=torch.int64).contiguous()
    perm, counts = partition_perm(ids, None, n, npart)
    g = take_frame(f, perm)
    out, off = [], 0
    for c in counts.tolist():
        out.append(g.slice(off, off + c))
        off += c
    return out


# ---------------------------------------------------------------- hash combine

def quantise_capacity(n: int) -> int:
    """Round ``n`` up to five significant bits (<= 6 % above ``n``): table sizes
    follow the cardinality closely — every scan of the slot array (init,
    compaction, exchange) is proportional to it — while slightly different
    estimates still map to the same pooled table."""
    n = max(int(n), 1024)
    k = n.bit_length() - 5
    return n if k <= 0 else ((n + (1 << k) - 1) >> k) << k


def table_capacity(distinct_estimate: int) -> int:
    """Main-level slots for an expected number of distinct keys (load <= ~0.55)."""
    return quantise_capacity(int(1.8 * distinct_estimate) + 1024)


def _norm_dev(device) -> torch.device:
    d = torch.device(device)
    if d.type == "cuda" and d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    return d


class HashTable:
    """HBM open-addressing table used by the combine kernels.

    Array-of-structs: ``slots[cap, 2]`` int64 — column 0 the key, column 1 the
    value bits — so a probe touches one 32-byte sector."""

    #: Slots of the L2-resident warm level placed in front of large tables
    #: (4 Mi slots = 64 MiB, about half of the B200's L2).
    WARM_SLOTS = int(os.environ.get("BIGSLICE_WARM_SLOTS", 1 << 22))

    def __init__(self, capacity: int, vdtype: torch.dtype, op: str, device, warm: Optional[int] = None):
        require()
        cap = quantise_capacity(capacity)
        if warm is None:
            warm = self.WARM_SLOTS if cap >= 2 * self.WARM_SLOTS else 0
        self.cap, self.warm, self.total = cap, warm, cap + warm
        self.vdtype, self.op, self.device = vdtype, op, _norm_dev(device)
        self.slots = torch.empty((self.total, 2), dtype=torch.int64, device=device)
        self.status = torch.zeros(4, dtype=torch.int64, device=device)
        # destination tracking (see track()): counts of first-time insertions per shard
        self.nparts = 0
        self.dest_counts: Optional[torch.Tensor] = None
        self.ready: Optional[torch.cuda.Event] = None  # pending background re-init
        self.reset()

    def track(self, nparts: int) -> None:
        """Count, while combining, how many distinct keys fall in each of the
        ``nparts`` shards of the default partitioner (``murmur3(key) % nparts``);
        the table exchange then needs no counting pass.  0 disables."""
        if nparts and nparts > 256:
            nparts = 0
        self.nparts = int(nparts or 0)
        if self.nparts:
            if self.dest_counts is None:
                self.dest_counts = torch.zeros(256, dtype=torch.int64, device=self.device)
            else:
                self.dest_counts.zero_()

    def reset(self):
        self.nparts = 0
        _chk(_lib.bs_table_init(_ptr(self.slots), c_ll(self.total), c_int(VTYPE[self.vdtype]),
                                c_int(OPS[self.op]), _ptr(self.status), _stream()), "table_init")

    def combine(self, keys: torch.Tensor, vals: torch.Tensor, force_simple: bool = False):
        keys, vals = keys.contiguous(), vals.contiguous()
        _chk(_lib.bs_hash_combine(_ptr(keys), _ptr(vals), c_ll(keys.numel()),
                                  c_int(VTYPE[self.vdtype]), c_int(OPS[self.op]), _ptr(self.slots),
                                  c_ll(self.warm), c_ll(self.cap), _ptr(self.status),
                                  c_int(1 if force_simple else 0), c_int(self.nparts),
                                  _ptr(self.dest_counts if self.nparts else None), _stream()),
             "hash_combine")

    def column_desc(self):
        """(ptr, row_bytes, stride, dtype, tail) of the key and value fields for
        PeerGroup.exchange_raw (shuffle the table in place, skipping EMPTY)."""
        p = self.slots.data_ptr()
        vb = torch.empty(0, dtype=self.vdtype).element_size()
        return [(p, 8, 16, torch.int64, ()), (p + 8, vb, 16, self.vdtype, ())]

    def identity_bits(self) -> int:
        _lib.bs_identity_bits.restype = ctypes.c_ulonglong
        return int(_lib.bs_identity_bits(c_int(VTYPE[self.vdtype]), c_int(OPS[self.op])))

    def reset_status(self):
        """Prepare a *clean* table (every slot EMPTY, e.g. after ``compact(clear=True)``
        or a clearing exchange) for reuse: only the 32-byte status block is rewritten."""
        _chk(_lib.bs_table_status_reset(_ptr(self.status), c_int(VTYPE[self.vdtype]),
                                        c_int(OPS[self.op]), _stream()), "table_status_reset")
        self.nparts = 0

    def compact(self, max_rows: Optional[int] = None, clear: bool = False
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Dense (keys, vals) of the occupied slots; syncs once for the count.
        ``clear=True`` empties the table in the same pass (scan-and-clear)."""
        cap_out = self.total + 1 if max_rows is None else max_rows
        ok = torch.empty(cap_out, dtype=torch.int64, device=self.device)
        ov = torch.empty(cap_out, dtype=self.vdtype, device=self.device)
        _chk(_lib.bs_table_compact(_ptr(self.slots), c_ll(self.total), c_int(VTYPE[self.vdtype]),
                                   _ptr(ok), _ptr(ov), _ptr(self.status), c_int(1 if clear else 0),
                                   ctypes.c_ulonglong(self.identity_bits() if clear else 0),
                                   c_ll(cap_out), _stream()), "table_compact")
        st = self.status.tolist()
        if st[0]:
            raise TableFull(st[0])
        n = st[3]
        return ok[:n], ov[:n]

    def compact_into(self, out_keys: torch.Tensor, out_vals: torch.Tensor, clear: bool = True) -> None:
        """Asynchronous compaction into caller-provided buffers (rows beyond their
        capacity are dropped; ``status.out_count`` still reports the rows needed).
        No host synchronisation: pair with ``ReduceExchange.publish``."""
        _chk(_lib.bs_table_compact(_ptr(self.slots), c_ll(self.total), c_int(VTYPE[self.vdtype]),
                                   _ptr(out_keys), _ptr(out_vals), _ptr(self.status), c_int(1 if clear else 0),
                                   ctypes.c_ulonglong(self.identity_bits() if clear else 0),
                                   c_ll(out_keys.numel()), _stream()), "table_compact")

    def failed(self) -> int:
        return int(self.status[0])


class TableFull(RuntimeError):
    def __init__(self, n):
        super().__init__("hash table full: %d rows found no slot" % n)
        self.rows = n


def can_hash_combine(f: Frame, op: str) -> bool:
    """``<key prefix, one value>`` frames whose key fits one 64-bit word (ops/keypack.py:
    an int64/float64 column, a narrower scalar column, or two columns of <= 4 bytes)."""
    npre = f.prefix()
    if op not in OPS or f.num_out() != npre + 1:
        return False
    from . import keypack
    if keypack.layout_for(f.typ) is None:
        return False
    v = f.col(npre)
    return (all(isinstance(f.col(i), torch.Tensor) and f.col(i).dim() == 1 for i in range(npre))
            and isinstance(v, torch.Tensor) and v.dtype in VTYPE and v.dim() == 1)


# How a used combine table gets ready for its next use: "clear" = the pass that
# reads it out (compaction / table scatter) also empties it; "reinit" = it is
# re-initialised on a side stream, off the critical path of the next user;
# "init" = the next user re-initialises it on its own stream.
RECYCLE = os.environ.get("BIGSLICE_TABLE_RECYCLE", "init")


class TablePool:
    """Clean hash tables kept between uses.

    A table handed back with :meth:`put` must be *clean* (all slots EMPTY): its
    last reader was a clearing pass (``compact(clear=True)`` or the clearing
    table exchange), so :meth:`get` only resets the status block instead of
    re-initialising hundreds of megabytes of slots."""

    def __init__(self, max_bytes: int = 8 << 30):
        self._lock = threading.Lock()
        self._free: dict = {}
        self._made: dict = {}
        self.tables_per_key = int(os.environ.get("BIGSLICE_POOL_TABLES", 2))
        self._bytes = 0
        self.max_bytes = max_bytes
        self.hits = 0
        self.misses = 0

    @staticmethod
    def _key(t: "HashTable"):
        return (t.cap, t.warm, t.vdtype, t.op, str(t.device))

    def get(self, capacity: int, vdtype: torch.dtype, op: str, device, warm: Optional[int] = None
            ) -> "HashTable":
        cap = quantise_capacity(capacity)
        if warm is None:
            warm = HashTable.WARM_SLOTS if cap >= 2 * HashTable.WARM_SLOTS else 0
        key = (cap, warm, vdtype, op, str(_norm_dev(device)))
        with self._lock:
            lst = self._free.get(key) or []
            t = None
            # prefer a table whose background re-init has finished; while the only
            # candidate is still being re-initialised, grow to a second table so
            # consecutive steps ping-pong instead of waiting for the init
            for i, c in enumerate(lst):
                if c.ready is None or c.ready == "dirty" or c.ready.query():
                    t = lst.pop(i)
                    break
            if t is None and lst and self._made.get(key, 0) >= self.tables_per_key:
                t = lst.pop(0)
            if t is not None:
                self._bytes -= t.total * 16
                self.hits += 1
            else:
                self.misses += 1
                self._made[key] = self._made.get(key, 0) + 1
        if t is None:
            return HashTable(cap, vdtype, op, device, warm=warm)
        if t.ready == "dirty":
            t.ready = None
            t.reset()
        elif t.ready is not None:
            torch.cuda.current_stream().wait_event(t.ready)  # background re-init (device-side wait)
            t.ready = None
        else:
            t.reset_status()
        return t

    def put(self, t: "HashTable", clean: bool = True) -> None:
        """Return a table.  ``clean=False``: its slots still hold the last result;
        they are re-initialised on a side stream (ordered after the work queued on
        the current stream), off the critical path of the next user."""
        with self._lock:
            if self._bytes + t.total * 16 > self.max_bytes:
                return
        if not clean and RECYCLE == "init":
            t.ready = "dirty"  # re-initialised by the next user, on its own stream
        elif not clean:
            side = self._side_stream(t.device)
            done = torch.cuda.Event()
            done.record()                    # everything queued so far reads the table
            side.wait_event(done)
            t.slots.record_stream(side)
            t.status.record_stream(side)
            with torch.cuda.stream(side):
                t.reset()
                ev = torch.cuda.Event()
                ev.record()
            t.ready = ev
        with self._lock:
            self._free.setdefault(self._key(t), []).append(t)
            self._bytes += t.total * 16

    def _side_stream(self, device) -> "torch.cuda.Stream":
        with self._lock:
            st = getattr(self, "_streams", None)
            if st is None:
                st = self._streams = {}
            key = str(device)
            if key not in st:
                st[key] = torch.cuda.Stream(device=device)
            return st[key]

    def clear(self) -> None:
        with self._lock:
            self._free.clear()
            self._made.clear()
            self._bytes = 0


_pool = TablePool()


def table_pool() -> TablePool:
    return _pool


def hash_combine(keys: torch.Tensor, vals: torch.Tensor, op: str,
                 distinct_hint: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Reduce ``vals`` by ``keys`` with the hash-aggregation kernels."""
    n = keys.numel()
    cap = table_capacity(distinct_hint if distinct_hint is not None else min(n, 1 << 22))
    while True:
        t = _pool.get(max(cap, 2048), vals.dtype, op, keys.device)
        t.combine(keys, vals)
        try:
            clear = RECYCLE == "clear"
            out = t.compact(clear=clear)
            _pool.put(t, clean=clear)
            return out
        except TableFull:
            cap = t.cap * 4  # rare: grow and redo (the overflowed table is dropped)


def hash_combine_frame(f: Frame, op: str) -> Frame:
    from . import keypack
    lay, npre = keypack.layout_for(f.typ), f.prefix()
    if lay.plain_int64:
        k, v = hash_combine(f.col(0), f.col(1), op)
        return Frame([k, v], f.typ)
    word = keypack.pack([f.col(i) for i in range(npre)], lay)
    k, v = hash_combine(word, f.col(npre), op)
    return Frame(keypack.unpack(k, lay) + [v], f.typ)


# ---------------------------------------------------------------- sort

def sort_pairs_i64(keys: torch.Tensor, vals: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stable radix sort of int64 keys with a 64-bit payload (default: row
    index, i.e. the sorting permutation)."""
    lib = require()
    n = keys.numel()
    dev = keys.device
    ka = keys.contiguous().clone()
    va = torch.arange(n, dtype=torch.int64, device=dev) if vals is None else vals.contiguous().clone()
    if n <= 1:
        return ka, va
    kb, vb = torch.empty_like(ka), torch.empty_like(va)
    scratch = torch.empty(int(lib.bs_sort_scratch_bytes(c_ll(n))) // 8 + 8, dtype=torch.int64, device=dev)
    rc = lib.bs_radix_sort_pairs(_ptr(ka), _ptr(va), _ptr(kb), _ptr(vb), c_ll(n), _ptr(scratch), _stream())
    if rc < 0:
        raise RuntimeError("radix sort failed: cuda error %d" % -rc)
    return (kb, vb) if rc == 1 else (ka, va)


# ---------------------------------------------------------------- compaction / segments

def compact_indices(mask: torch.Tensor) -> torch.Tensor:
    """Indices of the true elements of a bool mask, in order."""
    lib = require()
    mask = mask.contiguous()
    n = mask.numel()
    out = torch.empty(n, dtype=torch.int64, device=mask.device)
    scratch = torch.empty(int(lib.bs_compact_scratch_bytes(c_ll(n))) // 8 + 2, dtype=torch.int64,
                          device=mask.device)
    count = torch.zeros(1, dtype=torch.int64, device=mask.device)
    _chk(lib.bs_compact_indices(_ptr(mask), c_ll(n), _ptr(out), _ptr(scratch), _ptr(count), _stream()),
         "compact_indices")
    return out[:int(count)]


def segmented_reduce(vals: torch.Tensor, starts: torch.Tensor, nseg: int, op: str) -> torch.Tensor:
    lib = require()
    vals = vals.contiguous()
    seg = (torch.cumsum(starts.to(torch.int64), 0) - 1).contiguous()
    out = torch.empty(nseg, dtype=vals.dtype, device=vals.device)
    _chk(lib.bs_segreduce(_ptr(vals), _ptr(seg), c_ll(vals.numel()), _ptr(out), c_ll(nseg),
                          c_int(VTYPE[vals.dtype]), c_int(OPS[op]), _stream()), "segreduce")
    return out


def expand_indices(offsets: torch.Tensor) -> torch.Tensor:
    """src row for every output slot given exclusive offsets (n+1)."""
    lib = require()
    offsets = offsets.contiguous()
    n = offsets.numel() - 1
    total = int(offsets[-1])
    src = torch.empty(total, dtype=torch.int64, device=offsets.device)
    _chk(lib.bs_expand_indices(_ptr(offsets), c_ll(n), c_ll(total), _ptr(src), _stream()),
         "expand_indices")
    return src


def _orderable_i64(v: torch.Tensor) -> torch.Tensor:
    """int64 tensor whose signed order equals the order of ``v`` (ints: widened;
    floats: the usual sign-magnitude -> two's-complement bit trick; NaNs are the largest)."""
    if v.dtype in (torch.int64,):
        return v.contiguous()
    if v.dtype in (torch.int8, torch.int16, torch.int32, torch.uint8, torch.bool):
        return v.to(torch.int64)
    if v.dtype in (torch.float16, torch.bfloat16):
        v = v.to(torch.float32)
    if v.dtype.is_floating_point:
        v = v + 0.0  # -0.0 -> +0.0: the two zeros are one value
    if v.dtype == torch.float32:
        b = v.contiguous().view(torch.int32).to(torch.int64)
        return torch.where(b < 0, b ^ 0x7FFFFFFF, b)
    if v.dtype == torch.float64:
        b = v.contiguous().view(torch.int64)
        return torch.where(b < 0, b ^ 0x7FFFFFFFFFFFFFFF, b)
    raise TypeError("topk: unsupported dtype %s" % v.dtype)


def topk(values: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(values, indices) of the ``k`` largest elements, descending; equal values
    are listed in index order, and when ties straddle the cut it is unspecified
    which of them are kept (as with ``torch.topk``).  Radix select on the device
    (``csrc/misc.cu: bs_topk_i64``)."""
    lib = require()
    n = values.numel()
    k = min(int(k), n)
    if k <= 0:
        e = torch.empty(0, dtype=torch.int64, device=values.device)
        return values[:0], e
    keys = _orderable_i64(values.reshape(-1))
    idx = torch.empty(k, dtype=torch.int64, device=values.device)
    scratch = torch.empty(1040 // 8 + 1, dtype=torch.int64, device=values.device)
    rc = lib.bs_topk_i64(_ptr(keys), c_ll(n), c_ll(k), _ptr(idx), _ptr(scratch), _stream())
    if rc != 0:
        raise RuntimeError("bigslice_b200 kernel topk failed: cuda error %d" % -rc)
    # k is small: order the winners (value descending, then index ascending)
    kv = keys[idx]
    order = torch.argsort(idx, stable=True)
    order = order[torch.argsort(kv[order], descending=True, stable=True)]
    idx = idx[order]
    return values.reshape(-1)[idx], idx


# ---------------------------------------------------------------- merge path

def merge_sorted_i64(ka: torch.Tensor, kb: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stable merge of two ascending int64 key columns (``csrc/merge.cu``):
    ``(keys, src)`` where ``src[i]`` is the position of output row ``i`` in the
    concatenation ``[a | b]`` (rows of ``a`` first among equal keys)."""
    lib = require()
    lib.bs_merge_scratch_elems.restype = c_ll
    ka, kb = ka.contiguous(), kb.contiguous()
    na, nb = ka.numel(), kb.numel()
    dev = ka.device
    keys = torch.empty(na + nb, dtype=torch.int64, device=dev)
    src = torch.empty(na + nb, dtype=torch.int64, device=dev)
    scratch = torch.empty(int(lib.bs_merge_scratch_elems(c_ll(na), c_ll(nb))), dtype=torch.int64, device=dev)
    _chk(lib.bs_merge_sorted_i64(_ptr(ka), c_ll(na), _ptr(kb), c_ll(nb), _ptr(keys), _ptr(src), _ptr(scratch),
                                 _stream()), "merge_sorted_i64")
    return keys, src


def merge_frames(fa: Frame, fb: Frame) -> Frame:
    """Merge two frames sorted by one int64 key column (device)."""
    keys, src = merge_sorted_i64(fa.col(0), fb.col(0))
    cols = [keys]
    for j in range(1, fa.num_out()):
        cols.append(C.col_take(C.col_concat([fa.col(j), fb.col(j)], fa.out(j)), src))
    return Frame(cols, fa.typ)


# ---------------------------------------------------------------- cogroup (sort-merge join)

def run_starts(sorted_keys: torch.Tensor) -> torch.Tensor:
    """Indices where a new key starts in a key-sorted int64 column (run detection
    kernel + stream compaction)."""
    lib = require()
    n = sorted_keys.numel()
    flags = torch.empty(n, dtype=torch.uint8, device=sorted_keys.device)
    _chk(lib.bs_cg_run_flags(_ptr(sorted_keys), c_ll(n), _ptr(flags), _stream()), "cg_run_flags")
    return compact_indices(flags.view(torch.bool))


def cogroup_search(mine: torch.Tensor, other: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """For every key of the ascending list ``mine``: its lower-bound rank in the
    ascending list ``other`` and whether it occurs there (0/1)."""
    lib = require()
    n = mine.numel()
    r = torch.empty(n, dtype=torch.int64, device=mine.device)
    m = torch.empty(n, dtype=torch.int64, device=mine.device)
    _chk(lib.bs_cg_search(_ptr(mine), c_ll(n), _ptr(other), c_ll(other.numel()), _ptr(r), _ptr(m), _stream()),
         "cg_search")
    return r, m


def cogroup_union2(ka: torch.Tensor, sa: torch.Tensor, na: int, kb: torch.Tensor, sb: torch.Tensor, nb: int
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Union of two sorted *distinct* key lists with CSR group sizes.

    ``ka``/``kb``: distinct keys of each input (ascending); ``sa``/``sb``: the row
    index where each key's run starts in its sorted input of ``na``/``nb`` rows.
    Returns ``(union_keys, sizes_a, sizes_b)``: for every key of the union how many
    rows each input holds (0 when absent).  Binary-search merge on the device
    (``csrc/cogroup.cu``)."""
    lib = require()
    dev = ka.device
    a, b = ka.numel(), kb.numel()
    ra, ma = torch.empty(a, dtype=torch.int64, device=dev), torch.empty(a, dtype=torch.int64, device=dev)
    rb, mb = torch.empty(b, dtype=torch.int64, device=dev), torch.empty(b, dtype=torch.int64, device=dev)
    _chk(lib.bs_cg_search(_ptr(ka), c_ll(a), _ptr(kb), c_ll(b), _ptr(ra), _ptr(ma), _stream()), "cg_search")
    _chk(lib.bs_cg_search(_ptr(kb), c_ll(b), _ptr(ka), c_ll(a), _ptr(rb), _ptr(mb), _stream()), "cg_search")
    ca = torch.cumsum(ma, 0) - ma   # common keys below (exclusive scan)
    cb = torch.cumsum(mb, 0) - mb
    common = int(ma.sum()) if a else 0
    nu = a + b - common
    uk = torch.empty(nu, dtype=torch.int64, device=dev)
    za = torch.zeros(nu, dtype=torch.int64, device=dev)
    zb = torch.zeros(nu, dtype=torch.int64, device=dev)
    _chk(lib.bs_cg_place(_ptr(ka), c_ll(a), _ptr(ra), _ptr(ma), _ptr(ca), _ptr(sa), c_ll(na), _ptr(uk),
                         _ptr(za), c_int(1), _stream()), "cg_place")
    _chk(lib.bs_cg_place(_ptr(kb), c_ll(b), _ptr(rb), _ptr(mb), _ptr(cb), _ptr(sb), c_ll(nb), _ptr(uk),
                         _ptr(zb), c_int(0), _stream()), "cg_place")
    return uk, za, zb


def crc32(t: torch.Tensor) -> int:
    """CRC32-IEEE of a device tensor's bytes (device per-block CRCs + host
    GF(2) combine)."""
    lib = require()
    from .. import _native
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes == 0:
        return 0
    nb = int(lib.bs_crc32_nblocks(c_ll(nbytes)))
    out = torch.empty(nb, dtype=torch.int32, device=t.device)
    _chk(lib.bs_crc32_blocks(_ptr(t), c_ll(nbytes), _ptr(out), _stream()), "crc32")
    crcs = out.cpu().numpy().view("uint32")
    nat = _native.lib()
    block = int(lib.bs_crc32_block_bytes())
    if nat is not None:
        return int(nat.crc32_combine_blocks(crcs, nbytes, block))
    raise RuntimeError("crc32 combine needs the native module")
