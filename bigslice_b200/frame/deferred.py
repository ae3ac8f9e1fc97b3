"""Frames whose row count is still on the device.

The fused Reduce exchange and the asynchronous compaction write their results
into capacity-sized buffers; the number of valid rows lands in pinned host
memory when the kernels finish.  A :class:`DeferredFrame` lets the task graph
keep flowing (readers pass it along, results hold it) without a host
synchronisation: the first access that needs the length — ``len()``, a column,
a slice — waits for the producing stream's event and reads the count.

Resolution can fail: a device-detected overflow or peer failure surfaces here as
``TaskLostError`` (the producing exchange is re-run by the evaluator) or as a
fatal error (a rank's producers failed), mirroring how the reference surfaces
read errors on a task's output stream (``exec/bigmachine.go:1321-1402``).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

from ..slicetype import Type
from . import columns as C
from .frame import Frame


class DeferredFrame(Frame):
    """A :class:`Frame` over capacity buffers with a lazily known length."""

    __slots__ = ("_n", "_resolve", "ready_event", "_finalize", "_raw", "__weakref__")

    def __init__(self, data: Sequence, typ: Type, resolve: Callable[[], int], ready_event=None,
                 finalize: Optional[Callable[[list, int], list]] = None):
        # Frame.__init__ is bypassed: it would need the length
        self._finalize = None
        self._raw = list(data)
        self.typ = typ
        self._off = 0
        self._cap = C.col_len(self._raw[0]) if self._raw else 0
        self._n: Optional[int] = None
        self._resolve = resolve
        #: optional ``finalize(buffers, n) -> columns``: turns the raw result buffers
        #: into the frame's columns once the length is known (e.g. unpacking a packed
        #: key word into its key columns)
        self._finalize = finalize
        #: CUDA event after which the buffers are valid (consumers on other
        #: streams wait on it on the *device*)
        self.ready_event = ready_event

    @property
    def _len(self) -> int:  # shadows the base class slot
        n = self._n
        if n is None:
            n = int(self._resolve())
            if n > self._cap:
                raise RuntimeError("deferred frame: %d rows exceed the buffer capacity %d" % (n, self._cap))
            fin = self._finalize
            if fin is not None:
                self._finalize = None
                self.wait_ready()
                self._raw = list(fin(self._raw, n))
                self._cap = n
            self._n = n
            self._resolve = None
        return n

    @_len.setter
    def _len(self, v: int) -> None:
        self._n = v

    @property
    def _data(self):  # shadows the base class slot: the columns exist once finalized
        if self._finalize is not None:
            self._len
        return self._raw

    @_data.setter
    def _data(self, v) -> None:
        self._raw = v

    @property
    def device(self):
        for c in self._raw:
            return C.col_device(c)
        import torch
        return torch.device("cpu")

    def num_out(self) -> int:
        return len(self.typ.cols)

    def resolved(self) -> bool:
        return self._n is not None

    def wait_ready(self) -> None:
        """Make the current CUDA stream wait (on the device) for the producer."""
        ev = self.ready_event
        if ev is not None:
            import torch
            torch.cuda.current_stream().wait_event(ev)
            for c in self._raw:
                if isinstance(c, torch.Tensor) and c.is_cuda:
                    c.record_stream(torch.cuda.current_stream())
