"""NVLink peer memory and the fused exchange (host side).

A :class:`PeerGroup` gives every rank (one process per GPU) direct pointers to
every peer's *control block* (flags + count matrix) and *receive arena*:
memory is ``cudaMalloc``'d by ``_bscuda.so``, exported with CUDA IPC handles
that are all-gathered once over the bootstrap process group, and opened with
``cudaIpcOpenMemHandle`` (peer access enabled lazily).  After that a shuffle is
:meth:`PeerGroup.exchange`: the four kernels of ``csrc/exchange.cu`` on the
current stream — no NCCL, no host round trips except one small D2H read of the
plan to learn the received row counts.

NCCL (or gloo on CPU) all-to-all of the same data is available as
:func:`collective_exchange` — the measured baseline, and the path for columns
the kernel does not handle (strings, lists, objects).

Takes the place of the reference's pull-based shuffle transport: ``Worker.Read``
streams dialled per producer machine (``exec/bigmachine.go:857-1005``,
``:1306-1319``) with retry/reopen semantics (``:1351-1393``); here rows are pushed
once, from inside the partition kernel, and failures surface as an error on every
rank (bounded spins, abort flags, the vote carried with the counts).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ..ops import gpu
from ..utils.spans import span

c_void_p, c_int, c_ll, c_ull = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong

MAXW, MAXP, MAXC, MAXK = 16, 256, 8, 4


class XcPeers(ctypes.Structure):
    _fields_ = [("world", c_int), ("rank", c_int), ("ctrl", c_void_p * MAXW),
                ("arena", c_void_p * MAXW), ("arena_bytes", c_ull)]


class XcSpec(ctypes.Structure):
    _fields_ = [("n", c_ll), ("nparts", c_int), ("ncols", c_int), ("nkeys", c_int),
                ("skip_empty", c_int), ("ids", c_void_p), ("cols", c_void_p * MAXC),
                ("es", c_int * MAXC), ("stride", c_int * MAXC)]


class _DevMem:
    """Exposes raw device memory to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int, owner=None):
        self.owner = owner
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
            "strides": None}


def device_view(ptr: int, nbytes: int, device, owner=None) -> torch.Tensor:
    """uint8 tensor aliasing ``nbytes`` of device memory at ``ptr`` (zero copy)."""
    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevMem(ptr, nbytes, owner), device=device)


class ExchangeVoteError(RuntimeError):
    """The vote carried by an exchange failed on some rank (raised on all ranks):
    kind 3 = a rank's producers failed, 4 = ranks issued different invocations."""

    def __init__(self, kind: int):
        super().__init__("a rank reported failed producers" if kind == 3
                         else "ranks issued different invocations")
        self.kind = kind


class ExchangeError(RuntimeError):
    pass


class PeerGroup:
    """Symmetric NVLink-mapped buffers shared by the ranks of one box."""

    def __init__(self, group=None, arena_bytes: int = 1 << 30, device=None):
        lib = gpu.require()
        self.lib = lib
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > MAXW:
            raise ExchangeError("fused exchange supports at most %d ranks" % MAXW)
        self.device = torch.device(device if device is not None else
                                   "cuda:%d" % torch.cuda.current_device())
        sizes = [c_int() for _ in range(4)]
        lib.bs_xc_sizes(*[ctypes.byref(s) for s in sizes])
        self.ctrl_bytes, self.plan_bytes, peers_b, spec_b = [s.value for s in sizes]
        if peers_b != ctypes.sizeof(XcPeers) or spec_b != ctypes.sizeof(XcSpec):
            raise ExchangeError("ctypes/CUDA struct layout mismatch (%d/%d, %d/%d)" % (
                peers_b, ctypes.sizeof(XcPeers), spec_b, ctypes.sizeof(XcSpec)))
        offs = [c_ll() for _ in range(5)]
        lib.bs_xc_plan_offsets(*[ctypes.byref(o) for o in offs])
        (self.off_recv_total, self.off_recv_off, self.off_arena_end, self.off_send_count,
         self.off_error) = [o.value for o in offs]
        self.arena_bytes = int(arena_bytes)
        # local allocations + IPC handles
        self._ctrl = c_void_p()
        self._arena = c_void_p()
        h_ctrl = (ctypes.c_ubyte * 64)()
        h_arena = (ctypes.c_ubyte * 64)()
        self._chk(lib.bs_ipc_alloc(c_ll(self.ctrl_bytes), ctypes.byref(self._ctrl), h_ctrl), "ipc_alloc ctrl")
        self._chk(lib.bs_ipc_alloc(c_ll(self.arena_bytes), ctypes.byref(self._arena), h_arena), "ipc_alloc arena")
        mine = (bytes(h_ctrl), bytes(h_arena), torch.cuda.current_device(), os.getpid())
        everyone: List = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self.peers = XcPeers()
        self.peers.world, self.peers.rank = self.world, self.rank
        self.peers.arena_bytes = self.arena_bytes
        self._opened: List[int] = []
        for r, (hc, ha, devidx, pid) in enumerate(everyone):
            if r == self.rank:
                self.peers.ctrl[r], self.peers.arena[r] = self._ctrl.value, self._arena.value
                continue
            if devidx != torch.cuda.current_device():
                lib.bs_enable_peer(c_int(devidx))  # best effort; IPC open enables lazily too
            pc, pa = c_void_p(), c_void_p()
            self._chk(lib.bs_ipc_open((ctypes.c_ubyte * 64).from_buffer_copy(hc), ctypes.byref(pc)),
                      "ipc_open ctrl of rank %d" % r)
            self._chk(lib.bs_ipc_open((ctypes.c_ubyte * 64).from_buffer_copy(ha), ctypes.byref(pa)),
                      "ipc_open arena of rank %d" % r)
            self.peers.ctrl[r], self.peers.arena[r] = pc.value, pa.value
            self._opened += [pc.value, pa.value]
        self.epoch = 0
        self.last_sent = 0
        self.barrier_epoch = 0
        self.plan = torch.zeros(self.plan_bytes, dtype=torch.uint8, device=self.device)
        self.err = torch.zeros(2, dtype=torch.int32, device=self.device)
        self.arena_view = device_view(self._arena.value, self.arena_bytes, self.device, self)
        self.stats = {"exchanges": 0, "bytes_sent": 0, "rows_sent": 0}
        dist.barrier(group=group)

    @staticmethod
    def _chk(rc: int, what: str):
        if rc != 0:
            raise ExchangeError("%s failed: cuda error %d" % (what, rc))

    def barrier(self) -> None:
        """Device-side barrier across ranks (flag all-to-all kernel)."""
        self.barrier_epoch += 1
        self._chk(self.lib.bs_xc_barrier(ctypes.byref(self.peers), c_ull(self.barrier_epoch),
                                         c_void_p(self.err.data_ptr()), gpu._stream()), "barrier")

    # ------------------------------------------------------------------
    def exchange(self, cols: Sequence[torch.Tensor], nparts: int, nkeys: int = 1,
                 ids: Optional[torch.Tensor] = None, skip_empty: bool = False,
                 zero_copy: bool = False) -> Dict[int, List[torch.Tensor]]:
        """Shuffle rows (given as equal-length column tensors) to the owners of
        ``nparts`` destination shards (shard d lives on rank d % world).

        Returns ``{local shard: [column tensors]}`` holding every row of that
        shard from all ranks.  Collective: every rank must call it in the same
        order."""
        if nparts > MAXP or len(cols) > MAXC or nkeys > MAXK:
            raise ExchangeError("exchange shape outside fused-kernel limits")
        if self.ROW_RECORDS and not skip_empty and not zero_copy:
            got = self.exchange_rows(cols, nparts, nkeys, ids)
            if got is not None:
                return got
        cols = [c.contiguous() for c in cols]
        n = int(cols[0].shape[0]) if cols else 0
        desc = []
        for c in cols:
            rb = c.element_size() * (c.numel() // c.shape[0] if c.shape[0] else
                                     int(np.prod(c.shape[1:])) if c.dim() > 1 else 1)
            desc.append((c.data_ptr(), rb, rb, c.dtype, tuple(c.shape[1:])))
        return self.exchange_raw(desc, n, nparts, nkeys, ids, skip_empty, zero_copy, keep=cols)

    #: generic frames move column by column by default; ``BIGSLICE_XCHG=rows`` selects
    #: the row-record kernel (TMA bulk stores), which measured slower at N=2
    #: (profiles/r2_reshard_bw_n2.jsonl) and stays opt-in until it wins
    ROW_RECORDS = os.environ.get("BIGSLICE_XCHG", "cols") == "rows"

    def exchange_rows(self, cols: Sequence[torch.Tensor], nparts: int, nkeys: int = 1,
                      ids: Optional[torch.Tensor] = None, vote: Tuple[int, int] = (0, 0)
                      ) -> Optional[Dict[int, List[torch.Tensor]]]:
        """Row-record exchange (``xchg_scatter_rows``: rows packed into 16-byte
        aligned records in shared memory, per-destination write combining, TMA bulk
        stores over NVLink; ``xchg_unpack_rows`` on arrival).  Returns None when the
        shape is outside the kernel's limits (the column mover handles it)."""
        cols = [c.contiguous() for c in cols]
        n = int(cols[0].shape[0]) if cols else 0
        spec = XcSpec()
        spec.n, spec.nparts, spec.ncols, spec.nkeys = n, nparts, len(cols), nkeys
        if ids is not None:
            ids = ids.to(torch.int64).contiguous()
            spec.ids = ids.data_ptr()
        rbs = []
        for i, c in enumerate(cols):
            rb = c.element_size() * (int(np.prod(c.shape[1:])) if c.dim() > 1 else 1)
            spec.cols[i], spec.es[i], spec.stride[i] = c.data_ptr(), rb, rb
            rbs.append(rb)
        rec = int(self.lib.bs_xc_rec_bytes(ctypes.byref(spec)))
        if nparts > 64 or nparts * rec * 8 > 128 * 1024:
            return None
        self.epoch += 1
        rc = self.lib.bs_xc_exchange_rows(ctypes.byref(self.peers), ctypes.byref(spec),
                                          c_void_p(self.plan.data_ptr()), c_ull(self.epoch),
                                          c_ull(vote[0]), c_ull(vote[1]), gpu._stream())
        if rc == -3:
            self.epoch -= 1
            return None
        self._chk(rc, "exchange_rows")
        with span("xchg/plan_d2h_sync"):
            plan = self.plan.cpu().numpy()
        err = int(plan[self.off_error:self.off_error + 4].view(np.int32)[0])
        if err == 1:
            raise ExchangeError("receive arena overflow (%d bytes): raise arena_bytes" % self.arena_bytes)
        if err in (3, 4):
            raise ExchangeVoteError(err)
        if err:
            raise ExchangeError("peer exchange aborted (timeout or peer failure)")
        recv_total = plan[self.off_recv_total:self.off_recv_total + 8 * MAXP].view(np.int64)
        recv_off = plan[self.off_recv_off:self.off_recv_off + 8 * MAXP * MAXC].view(np.uint64).reshape(MAXP, MAXC)
        send = plan[self.off_send_count:self.off_send_count + 8 * MAXP].view(np.int64)
        self.last_sent = int(send[:nparts].sum())
        self.stats["exchanges"] += 1
        self.stats["rows_sent"] += self.last_sent
        self.stats["bytes_sent"] += self.last_sent * rec
        out: Dict[int, List[torch.Tensor]] = {}
        for d in range(self.rank, nparts, self.world):
            rows = int(recv_total[d])
            got = [torch.empty((rows,) + tuple(c.shape[1:]), dtype=c.dtype, device=self.device) for c in cols]
            if rows:
                ospec = XcSpec()
                ospec.ncols = len(cols)
                for i, (g, rb) in enumerate(zip(got, rbs)):
                    ospec.cols[i], ospec.es[i], ospec.stride[i] = g.data_ptr(), rb, rb
                self._chk(self.lib.bs_xc_unpack_rows(c_void_p(self._arena.value + int(recv_off[d, 0])),
                                                     c_ll(rows), ctypes.byref(ospec), gpu._stream()),
                          "unpack_rows")
            out[d] = got
        return out

    def exchange_raw(self, desc, n: int, nparts: int, nkeys: int = 1,
                     ids: Optional[torch.Tensor] = None, skip_empty: bool = False,
                     zero_copy: bool = False, keep=None, vote: Tuple[int, int] = (0, 0)
                     ) -> Dict[int, List[torch.Tensor]]:
        """``vote = (status, fingerprint)`` travels with the counts: if any rank
        reports a non-zero status, or fingerprints differ, *every* rank's call
        raises :class:`ExchangeVoteError` (no host collective needed).

        Exchange columns described as ``(ptr, row_bytes, stride_bytes, dtype,
        tail_shape)`` — lets strided views (the key/value fields of a hash-table
        slot array) be shuffled without materialising them."""
        spec = XcSpec()
        spec.n, spec.nparts, spec.ncols, spec.nkeys = n, nparts, len(desc), nkeys
        spec.skip_empty = 1 if skip_empty else 0
        if ids is not None:
            ids = ids.to(torch.int64).contiguous()
            spec.ids = ids.data_ptr()
        row_bytes = []
        for i, (ptr, rb, stride, dt, tail) in enumerate(desc):
            spec.cols[i], spec.es[i], spec.stride[i] = ptr, rb, stride
            row_bytes.append(rb)
        if ids is None:
            for k in range(nkeys):
                if spec.es[k] not in (4, 8):
                    raise ExchangeError("fused exchange hashes 4- or 8-byte key columns only")
        self.epoch += 1
        self._chk(self.lib.bs_xc_exchange(ctypes.byref(self.peers), ctypes.byref(spec),
                                          c_void_p(self.plan.data_ptr()), c_ull(self.epoch),
                                          c_ull(0), c_ull(vote[0]), c_ull(vote[1]), gpu._stream()),
                  "exchange")
        return self._collect(desc, nparts, zero_copy)

    def exchange_table(self, table, nparts: int, clear: bool = True, vote: Tuple[int, int] = (0, 0)
                       ) -> Dict[int, List[torch.Tensor]]:
        """Shuffle a :class:`~bigslice_b200.ops.gpu.HashTable` in place with the
        specialised table kernels: rows of shard ``murmur3(key) % nparts`` arrive
        as ``[keys, values]`` views of this rank's arena (zero copy; valid until
        the next exchange).  With ``clear`` the table is left empty — ready for
        reuse after ``reset_status()`` — so no separate init pass is needed.  The
        table's reserved EMPTY key is *not* exchanged: its combined value stays in the
        status block for the caller to route (the Reduce path proper uses
        ``parallel/rx.py``, whose scatter kernel routes it like any other row)."""
        vb = torch.empty(0, dtype=table.vdtype).element_size()
        self.epoch += 1
        self._chk(self.lib.bs_xc_exchange_table(
            ctypes.byref(self.peers), c_void_p(table.slots.data_ptr()), ctypes.c_longlong(table.total),
            ctypes.c_int(vb), ctypes.c_int(nparts), c_void_p(self.plan.data_ptr()), c_ull(self.epoch),
            c_ull(0), ctypes.c_int(1 if clear else 0),
            c_ull(table.identity_bits() if clear else 0),
            c_void_p(table.dest_counts.data_ptr() if table.nparts == nparts and nparts else 0),
            c_ull(vote[0]), c_ull(vote[1]), gpu._stream()), "exchange_table")
        desc = [(0, 8, 16, torch.int64, ()), (0, vb, 16, table.vdtype, ())]
        return self._collect(desc, nparts, zero_copy=True)

    def _collect(self, desc, nparts: int, zero_copy: bool) -> Dict[int, List[torch.Tensor]]:
        """Read the plan back (one small D2H, syncs the stream) and build the
        per-shard column views of the receive arena."""
        with span("xchg/plan_d2h_sync"):
            plan = self.plan.cpu().numpy()
        err = int(plan[self.off_error:self.off_error + 4].view(np.int32)[0])
        if err == 1:
            raise ExchangeError("receive arena overflow (%d bytes): raise arena_bytes" % self.arena_bytes)
        if err in (3, 4):
            raise ExchangeVoteError(err)
        if err:
            raise ExchangeError("peer exchange aborted (timeout or peer failure)")
        recv_total = plan[self.off_recv_total:self.off_recv_total + 8 * MAXP].view(np.int64)
        recv_off = plan[self.off_recv_off:self.off_recv_off + 8 * MAXP * MAXC].view(np.uint64).reshape(MAXP, MAXC)
        send = plan[self.off_send_count:self.off_send_count + 8 * MAXP].view(np.int64)
        row_bytes = sum(d[1] for d in desc)
        self.last_sent = int(send[:nparts].sum())
        self.stats["exchanges"] += 1
        self.stats["rows_sent"] += self.last_sent
        self.stats["bytes_sent"] += self.last_sent * row_bytes
        out: Dict[int, List[torch.Tensor]] = {}
        for d in range(self.rank, nparts, self.world):
            rows = int(recv_total[d])
            got = []
            for i, (ptr, rb, stride, dt, tail) in enumerate(desc):
                nb = rows * rb
                view = self.arena_view[int(recv_off[d, i]):int(recv_off[d, i]) + nb]
                t = view.view(dt).reshape((rows,) + tuple(tail))
                got.append(t if zero_copy else t.clone())
            out[d] = got
        return out

    def close(self) -> None:
        for p in self._opened:
            self.lib.bs_ipc_close(c_void_p(p))
        self._opened = []


# ----------------------------------------------------------------------------
# Baseline / fallback: partition locally, then a library all-to-all.

def collective_exchange_tensors(cols: Sequence[torch.Tensor], dest_rank: torch.Tensor,
                                group=None) -> List[torch.Tensor]:
    """All-to-all of rows by destination *rank* using torch.distributed
    (NCCL on GPU, gloo on CPU).  Returns the received columns."""
    world = dist.get_world_size(group)
    dest_rank = dest_rank.to(torch.int64)
    order = torch.sort(dest_rank, stable=True)[1]
    counts = torch.bincount(dest_rank, minlength=world)
    send_counts = counts.cpu()
    recv_counts = torch.empty_like(send_counts)
    sc = send_counts.to(dest_rank.device) if dest_rank.is_cuda else send_counts
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = rc.cpu()
    out = []
    in_splits, out_splits = send_counts.tolist(), recv_counts.tolist()
    for c in cols:
        s = c.index_select(0, order.to(c.device)).contiguous()
        r = torch.empty((int(recv_counts.sum()),) + tuple(c.shape[1:]), dtype=c.dtype, device=c.device)
        dist.all_to_all_single(r, s, out_splits, in_splits, group=group)
        out.append(r)
    return out
