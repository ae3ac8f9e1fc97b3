"""Top-n selection helper.

``top_n(values, n)`` returns the indices of the ``n`` largest values,
descending (host: ties broken towards the lower index; device: which ties make
the cut is unspecified).  On device tensors it uses the
radix-select kernel (``ops.gpu.topk`` -> ``csrc/misc.cu: bs_topk_i64``); parity with the reference heap
helper ``exec/topn.go:39-57``.
"""
from __future__ import annotations

import heapq
from typing import List

import torch


def top_n(values, n: int) -> List[int]:
    """Indices of the ``n`` largest elements of ``values``."""
    if n <= 0:
        return []
    if isinstance(values, torch.Tensor):
        k = min(n, values.numel())
        if k == 0:
            return []
        if values.is_cuda and values.dim() == 1:
            from ..ops import gpu
            if gpu.available():
                return gpu.topk(values, k)[1].tolist()
        return torch.topk(values, k, sorted=True).indices.tolist()
    idx = heapq.nlargest(min(n, len(values)), range(len(values)),
                         key=lambda i: (values[i], -i))
    return list(idx)


topN = top_n
