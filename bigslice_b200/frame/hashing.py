"""Bit-exact seeded MurmurHash3_x86_32 over column values.

The default shuffle partitioner is ``murmur3_32(LE bytes of key, seed) %
nshard`` so placement is deterministic and identical to the reference
(``frame/ops_builtin.go:143-164``: 4 LE bytes for <=32-bit ints and float32
bit patterns, 8 LE bytes for 64-bit ints / float64 bit patterns, raw bytes for
string/[]byte; bool -> seed+1/seed, ``frame/ops.go:112-131``).

Three implementations share these semantics: this vectorised torch version
(any device, used as the fp/CPU reference in tests), the native C++ host path
(``csrc/native.cpp``) and the sm_100a kernel (``csrc/hash_partition.cu``).
"""
from __future__ import annotations

import numpy as np
from typing import Optional

import torch

_M = 0xFFFFFFFF
_C1 = 0xCC9E2D51
_C2 = 0x1B873593


def _rotl(x: torch.Tensor, r: int) -> torch.Tensor:
    return ((x << r) | (x >> (32 - r))) & _M


def _mix_k(k: torch.Tensor) -> torch.Tensor:
    k = (k * _C1) & _M
    k = _rotl(k, 15)
    return (k * _C2) & _M


def _mix_h(h: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    h = h ^ _mix_k(k)
    h = _rotl(h, 13)
    return (h * 5 + 0xE6546B64) & _M


def _fmix(h: torch.Tensor) -> torch.Tensor:
    h = h ^ (h >> 16)
    h = (h * 0x85EBCA6B) & _M
    h = h ^ (h >> 13)
    h = (h * 0xC2B2AE35) & _M
    return h ^ (h >> 16)


_NATIVE_MIN = 2048  # below this the torch ops are as fast as the call overhead


def _native_hash(v: torch.Tensor, seed: int, width: int) -> Optional[torch.Tensor]:
    """C++ host path (``csrc/native.cpp: host_hash_i64``, GIL released) for CPU
    int64 tensors; None when it does not apply."""
    if v.device.type != "cpu" or v.dtype != torch.int64 or v.dim() != 1 or v.numel() < _NATIVE_MIN:
        return None
    from .. import _native
    nat = _native.lib()
    if nat is None or not hasattr(nat, "host_hash_i64"):
        return None
    v = v.contiguous()
    out = torch.empty_like(v)
    nat.host_hash_i64(v.data_ptr(), out.data_ptr(), v.numel(), width, seed & _M, False)
    return out


def hash32(v: torch.Tensor, seed: int) -> torch.Tensor:
    """murmur3 of 4 LE bytes; ``v`` holds uint32 values in int64."""
    out = _native_hash(v, seed, 4)
    if out is not None:
        return out
    h = torch.full_like(v, seed & _M)
    h = _mix_h(h, v & _M)
    return _fmix(h ^ 4)


def hash64(v: torch.Tensor, seed: int) -> torch.Tensor:
    """murmur3 of 8 LE bytes; ``v`` is int64 (bit pattern)."""
    out = _native_hash(v, seed, 8)
    if out is not None:
        return out
    lo = v & _M
    hi = (v >> 32) & _M
    h = torch.full_like(v, seed & _M)
    h = _mix_h(h, lo)
    h = _mix_h(h, hi)
    return _fmix(h ^ 8)


def murmur3_bytes(data: bytes, seed: int = 0) -> int:
    """Scalar murmur3_x86_32 over a byte string (host reference)."""
    h = seed & _M
    n = len(data)
    nblocks = n // 4
    for i in range(nblocks):
        k = int.from_bytes(data[4 * i:4 * i + 4], "little")
        k = (k * _C1) & _M
        k = ((k << 15) | (k >> 17)) & _M
        k = (k * _C2) & _M
        h ^= k
        h = ((h << 13) | (h >> 19)) & _M
        h = (h * 5 + 0xE6546B64) & _M
    tail = data[4 * nblocks:]
    k = 0
    if len(tail) >= 3:
        k ^= tail[2] << 16
    if len(tail) >= 2:
        k ^= tail[1] << 8
    if len(tail) >= 1:
        k ^= tail[0]
        k = (k * _C1) & _M
        k = ((k << 15) | (k >> 17)) & _M
        k = (k * _C2) & _M
        h ^= k
    h ^= n
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & _M
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & _M
    h ^= h >> 16
    return h


def hash_tensor(col: torch.Tensor, seed: int = 0) -> torch.Tensor:
    """Per-element hash of a tensor column -> int64 tensor of uint32 values."""
    dt = col.dtype
    if col.dim() == 2:  # vec columns hash as XOR-free sequential mix of lanes
        h = torch.full((col.shape[0],), seed & _M, dtype=torch.int64, device=col.device)
        for j in range(col.shape[1]):
            h = hash_tensor(col[:, j], 0) ^ _rotl(h, 5)
        return h
    if dt == torch.bool:
        return col.to(torch.int64) + (seed & _M) & _M
    if dt in (torch.int64,):
        return hash64(col, seed)
    if dt == torch.uint64:
        return hash64(col.view(torch.int64), seed)
    if dt == torch.float64:
        return hash64(col.contiguous().view(torch.int64), seed)
    if dt == torch.float32:
        return hash32(col.contiguous().view(torch.int32).to(torch.int64) & _M, seed)
    if dt in (torch.float16, torch.bfloat16):
        return hash32(col.contiguous().view(torch.int16).to(torch.int64) & 0xFFFF, seed)
    if dt in (torch.int8, torch.int16, torch.int32):
        # Go widens the signed value to uint32 (two's complement, 4 bytes).
        return hash32(col.to(torch.int64) & _M, seed)
    if dt in (torch.uint8,):
        return hash32(col.to(torch.int64), seed)
    if dt in (torch.uint16, torch.uint32):
        return hash32(col.to(torch.int64) & _M, seed)
    raise TypeError("unhashable tensor dtype %r" % (dt,))


def hash_objects(col: np.ndarray, seed: int = 0) -> torch.Tensor:
    """Per-element hash of a host object column (str/bytes)."""
    from .. import _native
    nat = _native.lib()
    if nat is not None and hasattr(nat, "murmur3_objects") and len(col):
        try:  # no Python-level loop: UTF-8 views of the str objects are hashed in C++
            return torch.from_numpy(nat.murmur3_objects(col if isinstance(col, list) else col.tolist(),
                                                        seed & _M))
        except TypeError:
            pass  # an unusual element: the loop below names its type
    enc = []
    for v in col:
        if isinstance(v, str):
            enc.append(v.encode("utf-8"))
        elif isinstance(v, (bytes, bytearray)):
            enc.append(bytes(v))
        elif v is None:
            enc.append(b"")
        else:
            raise TypeError("unhashable object column element %r" % (type(v),))
    nat = _native.lib()
    if nat is not None and enc:
        return torch.from_numpy(nat.murmur3_strings(enc, seed & _M).astype(np.int64))
    out = np.fromiter((murmur3_bytes(b, seed) for b in enc), dtype=np.int64, count=len(enc))
    return torch.from_numpy(out)
