"""Invocation transport: serialisable invocations, references, disk cache.

Reference: ``exec/invocation.go:16-142`` (``execInvocation`` with a custom gob
codec in which ``*Result`` arguments travel as ``invocationRef{Index}`` and are
turned back into Results in the receiver's address space),
``exec/bigmachine_inv_disk_cache.go:18-108`` (``invDiskCache``: compressed,
serialised-access on-disk cache of encoded invocations so that large argument
lists are encoded once and streamed to every worker) and the invocation
dependency bookkeeping of ``exec/bigmachine.go:79-145``.

Design here: every SPMD rank runs the same driver program, so invocations are
normally *re-created* on each rank rather than shipped.  The codec is used

* to ship invocations to executors that do not run the driver program
  (``MultiDeviceExecutor`` helpers, tooling),
* for the cross-rank *agreement check*: rank 0 broadcasts the fingerprint of
  each invocation and the other ranks refuse to continue when theirs differs
  (the SPMD analogue of the reference's worker-side ``Compile`` of exactly the
  driver's invocation).

The wire format is ``pickle`` of a plain dict (framework-internal, same trust
domain as ``torch.distributed`` object collectives), zlib-compressed on disk.
"""
from __future__ import annotations

import hashlib
import io
import os
import pickle
import shutil
import tempfile
import threading
import zlib
from typing import Any, Callable, Dict, Iterable, List, Optional, Set

from .. import errors
from ..func import Invocation, func_by_index

__all__ = ["InvocationRef", "encode_invocation", "decode_invocation", "fingerprint",
           "InvDiskCache", "InvocationGraph"]


class InvocationRef:
    """Stand-in for a ``Result`` argument while an invocation is in transit
    (``invocationRef``, invocation.go:24-29)."""

    __slots__ = ("index",)

    def __init__(self, index: int):
        self.index = int(index)

    def __eq__(self, other):
        return isinstance(other, InvocationRef) and other.index == self.index

    def __hash__(self):
        return hash(("invref", self.index))

    def __repr__(self):
        return "InvocationRef(%d)" % self.index


def _is_result(x) -> bool:
    from .session import Result
    return isinstance(x, Result)


def _swap_results(args: Iterable[Any]) -> List[Any]:
    out = []
    for a in args:
        if _is_result(a):
            out.append(InvocationRef(a.inv_index))
        else:
            out.append(a)
    return out


def encode_invocation(einv) -> bytes:
    """Serialise an ``ExecInvocation`` (Results become ``InvocationRef``s)."""
    inv = einv.inv
    fv = func_by_index(inv.func)
    args = _swap_results(inv.args)
    enc_args = []
    for i, a in enumerate(args):
        try:
            enc_args.append(pickle.dumps(a, protocol=pickle.HIGHEST_PROTOCOL))
        except Exception as e:  # noqa: BLE001 - reported with the argument position
            raise errors.fatal("encoding arg %d of %s: %s" % (i, fv.__name__, e)) from e
    doc = {"index": inv.index, "func": inv.func, "exclusive": inv.exclusive,
           "location": inv.location, "env": pickle.dumps(einv.env), "args": enc_args}
    return pickle.dumps(doc, protocol=pickle.HIGHEST_PROTOCOL)


def decode_invocation(data: bytes, resolve: Optional[Callable[[int], Any]] = None):
    """Inverse of :func:`encode_invocation`.  ``resolve(index)`` maps an
    ``InvocationRef`` back to a Result of this process; without it the refs are
    left in place (the caller substitutes them later, bigmachine.go:617-640)."""
    from .compile import ExecInvocation
    doc = pickle.loads(data)
    fv = func_by_index(doc["func"])
    args = []
    for i, raw in enumerate(doc["args"]):
        try:
            a = pickle.loads(raw)
        except Exception as e:  # noqa: BLE001
            raise errors.fatal("decoding arg %d of %s: %s" % (i, fv.__name__, e)) from e
        if isinstance(a, InvocationRef) and resolve is not None:
            a = resolve(a.index)
        args.append(a)
    inv = Invocation(doc["index"], doc["func"], args, doc["exclusive"], doc["location"])
    return ExecInvocation(inv, pickle.loads(doc["env"]))


def _fp_value(h, a, depth=0) -> None:
    """Feed a stable description of ``a`` into the hash.  Tensors and arrays
    contribute dtype/shape only (their contents are rank-local shards in SPMD
    programs); opaque objects contribute their type name."""
    import numpy as np
    import torch
    if _is_result(a):
        h.update(b"R%d;" % a.inv_index)
    elif isinstance(a, InvocationRef):
        h.update(b"R%d;" % a.index)
    elif isinstance(a, torch.Tensor):
        h.update(("T%s%s;" % (a.dtype, tuple(a.shape))).encode())
    elif isinstance(a, np.ndarray):
        h.update(("N%s%s;" % (a.dtype, a.shape)).encode())
    elif isinstance(a, (bool, int, float, str, bytes, type(None))):
        h.update(("V%s:%r;" % (type(a).__name__, a)).encode())
    elif isinstance(a, (list, tuple)) and depth < 3 and len(a) <= 64:
        h.update(("L%d[" % len(a)).encode())
        for x in a:
            _fp_value(h, x, depth + 1)
        h.update(b"];")
    elif isinstance(a, dict) and depth < 3 and len(a) <= 64:
        h.update(("D%d{" % len(a)).encode())
        for k in sorted(a, key=repr):
            _fp_value(h, k, depth + 1)
            _fp_value(h, a[k], depth + 1)
        h.update(b"};")
    else:
        h.update(("O%s;" % type(a).__qualname__).encode())


def fingerprint(inv: Invocation) -> str:
    """Digest identifying an invocation across ranks: Func index, call site
    independent argument structure.  Two ranks executing the same driver
    program produce equal fingerprints for equal invocations."""
    h = hashlib.sha1()
    h.update(("inv%d:f%d:x%d;" % (inv.index, inv.func, int(inv.exclusive))).encode())
    for a in inv.args:
        _fp_value(h, a)
    return h.hexdigest()


class InvDiskCache:
    """On-disk cache of encoded invocations keyed by invocation index
    (``invDiskCache``).  Access is serialised; entries are zlib streams."""

    def __init__(self, directory: Optional[str] = None):
        self._lock = threading.Lock()
        self._dir = directory
        self._own_dir = directory is None
        self._paths: Optional[Dict[int, str]] = {}

    def _init(self) -> None:
        if self._dir is None:
            try:
                self._dir = tempfile.mkdtemp(prefix="bigslice-inv-cache")
            except OSError as e:
                raise errors.Error("could not create invocation disk cache", cause=e) from e
        else:
            os.makedirs(self._dir, exist_ok=True)

    def get_or_create(self, index: int, create: Callable[[io.BufferedIOBase], None]) -> io.BufferedIOBase:
        """Reader of entry ``index``; ``create(writer)`` is called (once) to
        produce it when absent.  A failing ``create`` leaves no entry behind."""
        with self._lock:
            if self._paths is None:
                raise RuntimeError("InvDiskCache used after close")
            self._init()
            if index not in self._paths:
                path = os.path.join(self._dir, str(int(index)))
                try:
                    with open(path, "wb") as f:
                        w = _ZWriter(f)
                        create(w)
                        w.close()
                except BaseException as e:
                    try:
                        os.unlink(path)
                    except OSError:
                        pass
                    if isinstance(e, Exception):
                        raise errors.Error("could not create invocation disk cache entry", cause=e) from e
                    raise
                self._paths[index] = path
            try:
                f = open(self._paths[index], "rb")
            except OSError as e:
                raise errors.Error("could not open invocation disk cache entry", cause=e) from e
            return _ZReader(f)

    def __contains__(self, index: int) -> bool:
        with self._lock:
            return self._paths is not None and index in self._paths

    def close(self) -> None:
        with self._lock:
            if self._paths is None:
                raise RuntimeError("InvDiskCache closed twice")
            self._paths = None
            if self._dir is not None and self._own_dir:
                shutil.rmtree(self._dir, ignore_errors=True)


class _ZWriter(io.RawIOBase):
    def __init__(self, f):
        self._f = f
        self._z = zlib.compressobj(3)

    def writable(self):
        return True

    def write(self, b) -> int:
        self._f.write(self._z.compress(bytes(b)))
        return len(b)

    def close(self):
        if not self.closed:
            self._f.write(self._z.flush())
            super().close()


class _ZReader(io.RawIOBase):
    def __init__(self, f):
        self._f = f
        self._z = zlib.decompressobj()
        self._buf = b""
        self._eof = False

    def readable(self):
        return True

    def _fill(self, want: int) -> None:
        while not self._eof and (want < 0 or len(self._buf) < want):
            chunk = self._f.read(1 << 16)
            if not chunk:
                self._buf += self._z.flush()
                self._eof = True
                break
            self._buf += self._z.decompress(chunk)

    def read(self, n: int = -1) -> bytes:
        self._fill(n)
        if n < 0:
            out, self._buf = self._buf, b""
        else:
            out, self._buf = self._buf[:n], self._buf[n:]
        return out

    def readinto(self, b) -> int:
        data = self.read(len(b))
        b[:len(data)] = data
        return len(data)

    def close(self):
        if not self.closed:
            self._f.close()
            super().close()


class InvocationGraph:
    """Which invocations an invocation needs (through ``Result`` arguments), so
    a receiver can be sent — in order — everything required to compile it
    (``bigmachineExecutor.invocationDeps``, bigmachine.go:98-107, 196-238)."""

    def __init__(self):
        self._lock = threading.Lock()
        self._invs: Dict[int, Any] = {}
        self._deps: Dict[int, Set[int]] = {}

    def add(self, einv) -> None:
        deps = set()
        for a in einv.inv.args:
            if _is_result(a):
                deps.add(a.inv_index)
            elif isinstance(a, InvocationRef):
                deps.add(a.index)
        with self._lock:
            self._invs[einv.index] = einv
            self._deps[einv.index] = deps

    def get(self, index: int):
        with self._lock:
            return self._invs[index]

    def deps(self, index: int) -> Set[int]:
        with self._lock:
            return set(self._deps.get(index, ()))

    def closure(self, index: int) -> List[int]:
        """Indices required to compile ``index``, dependencies first
        (topological; the invocation itself is last)."""
        order: List[int] = []
        seen: Set[int] = set()
        with self._lock:
            def visit(i: int) -> None:
                if i in seen:
                    return
                seen.add(i)
                for d in sorted(self._deps.get(i, ())):
                    visit(d)
                order.append(i)
            visit(index)
        return order

    def discard(self, index: int) -> None:
        with self._lock:
            self._invs.pop(index, None)
            self._deps.pop(index, None)
