"""Per-stream codec session state for custom column codecs.

Custom ``Ops.encode``/``Ops.decode`` functions (see :mod:`.ops`) receive a
:class:`Session` that lives as long as one encoded stream, so a codec can keep
dictionaries / schemas across batches.

Capability parity: reference ``frame/codec.go:11-50`` (``Key``, ``Session.
State``, ``Encoder``/``Decoder`` interfaces).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Hashable


class Key:
    """An opaque, unique key for codec session state (``frame.Key``)."""

    __slots__ = ("name",)

    def __init__(self, name: str = ""):
        self.name = name

    def __repr__(self):
        return "frame.Key(%s)" % self.name


class Session:
    """Holds per-stream state for custom encoders and decoders."""

    def __init__(self):
        self._state: Dict[Hashable, Any] = {}

    def state(self, key: Hashable, init: Callable[[], Any]) -> Any:
        """Return the state for ``key``, creating it with ``init`` once."""
        if key not in self._state:
            self._state[key] = init()
        return self._state[key]

    State = state
