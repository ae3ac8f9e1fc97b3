"""Packing key prefixes into one 64-bit word for the device hash tables.

The combine/exchange kernels key their tables by an ``int64``.  Any key prefix
that fits in 8 bytes rides the same kernels: one 8-byte column (int64, uint64,
float64), one narrower scalar column, or two columns of at most 4 bytes each —
``(int32, int32)``, ``(float32, int16)``, … — packed as ``lo | hi << 32``.
Narrow integers are widened to 32 bits the way the reference hashes them
(``frame/ops_builtin.go:31-37,143-150``: the value as a uint32, 4 LE bytes),
floats contribute their IEEE bit patterns, so the partition function computed by
the kernels from the packed word (``key mode``) is bit-exact with
``Frame.hash``:

* mode 0: ``murmur3_u64(word)``                       (one 8-byte column)
* mode 1: ``murmur3_u32(lo32)``                       (one column of <= 4 bytes)
* mode 2: ``murmur3_u32(lo32) ^ murmur3_u32(hi32)``   (two columns of <= 4 bytes)

Equality of packed words is equality of the key tuples (for floats: of their
bit patterns, which is also what the reference's hash placement distinguishes).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

_NARROW = (torch.int8, torch.int16, torch.int32, torch.uint8, torch.float32)
_WIDE = (torch.int64, torch.float64)
for _n in ("uint16", "uint32"):
    if hasattr(torch, _n):
        _NARROW = _NARROW + (getattr(torch, _n),)
if hasattr(torch, "uint64"):
    _WIDE = _WIDE + (torch.uint64,)


class KeyLayout:
    __slots__ = ("mode", "dtypes")

    def __init__(self, mode: int, dtypes: Sequence[torch.dtype]):
        self.mode, self.dtypes = mode, tuple(dtypes)

    @property
    def plain_int64(self) -> bool:
        return self.mode == 0 and self.dtypes[0] == torch.int64

    def __repr__(self):
        return "KeyLayout(mode=%d, %s)" % (self.mode, ", ".join(str(d) for d in self.dtypes))


def layout_for(typ) -> Optional[KeyLayout]:
    """The packing of ``typ``'s key prefix, or None when it does not fit."""
    p = typ.prefix()
    if p not in (1, 2):
        return None
    dts = []
    for i in range(p):
        c = typ.out(i)
        if not getattr(c, "is_tensor", False) or getattr(c, "kind", "") in ("vec", "list", "bool"):
            return None
        dts.append(c.torch_dtype)
    if p == 1:
        if dts[0] in _WIDE:
            return KeyLayout(0, dts)
        if dts[0] in _NARROW:
            return KeyLayout(1, dts)
        return None
    if all(d in _NARROW for d in dts):
        return KeyLayout(2, dts)
    return None


def _word32(c: torch.Tensor) -> torch.Tensor:
    """The column's elements as non-negative int64 holding their 32-bit images."""
    if c.dtype == torch.float32:
        return c.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    return c.to(torch.int64) & 0xFFFFFFFF   # sign-extended to 64, truncated to 32: "widened to uint32"


def pack(cols: Sequence[torch.Tensor], lay: KeyLayout) -> torch.Tensor:
    if lay.mode == 0:
        c = cols[0].contiguous()
        return c if c.dtype == torch.int64 else c.view(torch.int64)
    if lay.mode == 1:
        return _word32(cols[0])
    return _word32(cols[0]) | (_word32(cols[1]) << 32)


def _unword32(w: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
    w32 = (w & 0xFFFFFFFF)
    if dt == torch.float32:
        return torch.where(w32 >= 0x80000000, w32 - (1 << 32), w32).to(torch.int32).view(torch.float32)
    if dt in (torch.int8, torch.int16, torch.int32):
        return torch.where(w32 >= 0x80000000, w32 - (1 << 32), w32).to(dt)
    return w32.to(dt)


def unpack(packed: torch.Tensor, lay: KeyLayout) -> List[torch.Tensor]:
    if lay.mode == 0:
        return [packed if lay.dtypes[0] == torch.int64 else packed.view(lay.dtypes[0])]
    if lay.mode == 1:
        return [_unword32(packed, lay.dtypes[0])]
    return [_unword32(packed, lay.dtypes[0]), _unword32(packed >> 32, lay.dtypes[1])]
