"""Frame sinks (reference ``sliceio/writer.go:14-18``)."""
from __future__ import annotations

from typing import Callable

from ..frame import Frame


class Writer:
    """Accepts frames; ``write`` must consume the whole frame or raise."""

    def write(self, frame: Frame) -> None:
        raise NotImplementedError


class FuncWriter(Writer):
    def __init__(self, fn: Callable[[Frame], None]):
        self.fn = fn

    def write(self, frame: Frame) -> None:
        self.fn(frame)


class DiscardWriter(Writer):
    def write(self, frame: Frame) -> None:
        return None
