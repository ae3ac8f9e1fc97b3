"""Replicating a (small) relation on every GPU through the NVSwitch.

``Replicator.all_gather(columns)``: every rank contributes its rows; afterwards
every rank holds all rows of all ranks.  The data path is one kernel per rank
(``csrc/cogroup.cu``):

* **multicast** (NVLS): the symmetric buffer has a multicast mapping
  (``torch.distributed._symmetric_memory`` does the allocation and the handle
  exchange — plumbing); ``cg_multicast_rows_kernel`` issues ``multimem.st`` so each
  16-byte store is replicated by the switch into every GPU's copy: a rank sends
  its rows *once*, not once per peer.
* **peer stores**: without multicast support the same kernel shape writes the
  rows to every peer's buffer with plain stores (W x the egress).
* **library**: no symmetric memory at all (CPU ranks, tests): ``all_gather``.

Used by :func:`bigslice_b200.cogroup.BroadcastJoin` — the broadcast-shaped join of
BASELINE config 3 that avoids shuffling the large side.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

c_void_p, c_ll, c_int = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int


def _align16(n: int) -> int:
    return (n + 15) & ~15


class Replicator:
    def __init__(self, group, nbytes: int, device, ctrl_group=None):
        self.group = group
        self.ctrl_group = ctrl_group if ctrl_group is not None else group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.nbytes = _align16(int(nbytes))
        self.mode = "library"
        self.buf: Optional[torch.Tensor] = None
        self.handle = None
        self.mc_ptr = 0
        self.peer_ptrs: Optional[torch.Tensor] = None
        self.why = ""
        if self.device.type == "cuda" and self.world > 1:
            try:
                import torch.distributed._symmetric_memory as symm
                self.buf = symm.empty(self.nbytes, dtype=torch.uint8, device=self.device)
                self.handle = symm.rendezvous(self.buf, group)
                self.mc_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)
                ptrs = list(self.handle.buffer_ptrs)
                self.peer_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
                self.mode = "multicast" if self.mc_ptr else "peer"
            except Exception as e:  # noqa: BLE001 - any failure means: use the library path
                self.why = "%s: %s" % (type(e).__name__, e)
                self.buf = None
                self.handle = None
                self.mode = "library"

    def all_gather(self, cols: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """All rows of every rank's ``cols`` (equal-typed columns), rank order."""
        n = int(cols[0].shape[0]) if cols else 0
        counts = [None] * self.world
        dist.all_gather_object(counts, n, group=self.ctrl_group)
        total = sum(counts)
        if self.mode == "library" or total == 0:
            return self._library(cols, counts)
        from ..ops import gpu
        lib = gpu.require()
        need = sum(_align16(total * c.element_size() * (c.numel() // max(c.shape[0], 1) if c.shape[0] else
                                                         int(torch.tensor(c.shape[1:]).prod()) if c.dim() > 1 else 1))
                   for c in cols)
        if need > self.nbytes:
            return self._library(cols, counts)
        out, base = [], 0
        before = sum(counts[:self.rank])
        self.handle.barrier()  # everybody is done reading the previous contents
        for c in cols:
            c = c.contiguous()
            rb = c.element_size() * (int(torch.tensor(c.shape[1:]).prod()) if c.dim() > 1 else 1)
            region = _align16(total * rb)
            nb = n * rb
            if nb % 16 or (before * rb) % 16:
                # 16-byte granularity of the vector stores: pad through a staging copy
                return self._library(cols, counts)
            if nb:
                if self.mode == "multicast":
                    rc = lib.bs_cg_multicast_rows(c_void_p(c.data_ptr()), c_ll(nb), c_void_p(self.mc_ptr),
                                                  c_ll(base + before * rb), gpu._stream())
                else:
                    rc = lib.bs_cg_peer_rows(c_void_p(c.data_ptr()), c_ll(nb),
                                             c_void_p(self.peer_ptrs.data_ptr()), c_int(self.world),
                                             c_ll(base + before * rb), gpu._stream())
                if rc != 0:
                    raise RuntimeError("row broadcast kernel failed: %d" % rc)
            view = self.buf[base:base + total * rb].view(c.dtype).reshape((total,) + tuple(c.shape[1:]))
            out.append(view)
            base += region
        self.handle.barrier()  # all ranks' rows have landed everywhere
        return [o.clone() for o in out]

    def _library(self, cols, counts) -> List[torch.Tensor]:
        if self.world == 1:
            return [c for c in cols]
        out = []
        kmax = max(counts)
        for c in cols:
            # equal-sized contributions (the library all_gather requires them): pad to
            # the largest shard, trim on arrival
            pad = torch.zeros((kmax,) + tuple(c.shape[1:]), dtype=c.dtype, device=c.device)
            pad[:c.shape[0]] = c
            via_host = c.is_cuda and dist.get_backend(self.group) == "gloo"
            src = pad.cpu() if via_host else pad
            parts = [torch.empty_like(src) for _ in counts]
            dist.all_gather(parts, src.contiguous(), group=self.group)
            got = torch.cat([p[:k] for p, k in zip(parts, counts)], 0)
            out.append(got.to(c.device) if via_host else got)
        return out
