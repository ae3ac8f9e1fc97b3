"""The Reduce exchange (host side of ``csrc/reduce_exchange.cu``).

A :class:`ReduceExchange` owns, per rank, an IPC-mapped control block and
landing arena, a private device state block and two rings of *pinned* report
slots.  One exchange = ``set_vote`` + one CUDA-graph launch of

    rx_begin -> rx_scatter -> rx_wait -> (rx_combine -> table_compact)* -> rx_end

on the exchange stream: no host round trip, no count all-gather, no NCCL.  The
host learns row counts and errors from the pinned report only when somebody
needs them (:class:`Pending.resolve`), so consecutive ``Session.run`` steps
pipeline on the device.

Every (parity, local shard, source) triple owns a fixed-capacity landing region
(``region_rows``); overflow, like every other device-detected failure, is
reported to *all* ranks through the vote and surfaces as a *lost* exchange that
is re-run with a larger arena (:meth:`ReduceExchange.grow`).

Takes the place of the reference's combiner hand-off + pull shuffle
(``exec/bigmachine.go:1084-1210``, ``:857-1005``) for Reduce.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ..ops import gpu

c_void_p, c_int, c_ll, c_ull = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong

MAXW, MAXP = 16, 64
NSLOTS = 64

# error codes (csrc/reduce_exchange.cu: RxError)
OK, LANDING_OVERFLOW, TABLE_OVERFLOW, PRODUCER_FAILED, FINGERPRINT, TIMEOUT, DEST_OVERFLOW = range(7)
ERROR_NAMES = {LANDING_OVERFLOW: "landing region overflow", TABLE_OVERFLOW: "combine table overflow",
               PRODUCER_FAILED: "a rank reported failed producers",
               FINGERPRINT: "ranks issued different invocations", TIMEOUT: "peer timeout / abort",
               DEST_OVERFLOW: "destination table or output buffer overflow"}
#: failures that are retried (the exchange is *lost*), as opposed to fatal ones
RETRIABLE = (LANDING_OVERFLOW, TABLE_OVERFLOW, TIMEOUT, DEST_OVERFLOW)


class RxPeers(ctypes.Structure):
    _fields_ = [("world", c_int), ("rank", c_int), ("ctrl", c_void_p * MAXW),
                ("arena", c_void_p * MAXW), ("region_rows", c_ull), ("half_bytes", c_ull),
                ("shards_per_rank", c_int), ("regions", c_int)]


class ExchangeFailed(RuntimeError):
    """An exchange finished with a device-agreed error (same verdict on every rank)."""

    def __init__(self, code: int, need_rows: int = 0):
        super().__init__("reduce exchange failed: %s" % ERROR_NAMES.get(code, "error %d" % code))
        self.code = code
        self.need_rows = need_rows

    @property
    def retriable(self) -> bool:
        return self.code in RETRIABLE


class _Ring:
    """``NSLOTS`` report slots in pinned, device-mapped host memory."""

    def __init__(self, lib, slot_bytes: int):
        self.lib = lib
        self.slot_bytes = slot_bytes
        self.ptr = c_void_p()
        rc = lib.bs_host_alloc(c_ll(slot_bytes * NSLOTS), ctypes.byref(self.ptr))
        if rc != 0:
            raise RuntimeError("bs_host_alloc failed: cuda error %d" % rc)
        buf = (ctypes.c_byte * (slot_bytes * NSLOTS)).from_address(self.ptr.value)
        self.words = np.frombuffer(buf, dtype=np.int64).reshape(NSLOTS, slot_bytes // 8)

    def read(self, ticket: int) -> np.ndarray:
        return self.words[ticket % NSLOTS].copy()

    def close(self):
        if self.ptr.value:
            self.lib.bs_host_free(self.ptr)
            self.ptr = c_void_p()


class Report:
    __slots__ = ("ticket", "error", "need_rows", "sent_rows", "recv_rows", "out_rows", "dest_failed")

    def __init__(self, words: np.ndarray, offs: Tuple[int, int, int]):
        self.ticket = int(words[0])
        self.error, self.need_rows, self.sent_rows = int(words[1]), int(words[2]), int(words[3])
        r, o, f = offs
        self.recv_rows = words[r // 8:r // 8 + MAXP].tolist()
        self.out_rows = words[o // 8:o // 8 + MAXP].tolist()
        self.dest_failed = words[f // 8:f // 8 + MAXP].tolist()


class Pending:
    """The outcome of device work the host has only *enqueued*: resolved (one
    event wait + a read of pinned memory) when somebody needs sizes or errors."""

    def __init__(self, owner: "ReduceExchange", ring: _Ring, ticket: int, event: "torch.cuda.Event",
                 on_report=None):
        self.owner, self.ring, self.ticket, self.event = owner, ring, ticket, event
        self.report: Optional[Report] = None
        self.on_report = on_report
        self._lock = threading.Lock()

    def ready(self) -> bool:
        return self.report is not None or self.event.query()

    def resolve(self) -> Report:
        if self.report is None:
            with self._lock:
                if self.report is None:
                    self.event.synchronize()
                    rep = Report(self.ring.read(self.ticket), self.owner.report_offs)
                    if rep.ticket != self.ticket:
                        raise RuntimeError("report slot %d was overwritten (ticket %d, expected %d)"
                                           % (self.ticket % NSLOTS, rep.ticket, self.ticket))
                    self.report = rep
                    if self.on_report is not None:
                        cb, self.on_report = self.on_report, None
                        cb(rep)
        return self.report

    def rows(self, j: int = 0) -> int:
        rep = self.resolve()
        if rep.error:
            raise ExchangeFailed(rep.error, rep.need_rows)
        return rep.out_rows[j]


class _Graph:
    __slots__ = ("exec", "nodes", "keep")

    def __init__(self, exec_, nodes, keep):
        self.exec, self.nodes, self.keep = exec_, nodes, keep


class ReduceExchange:
    """Symmetric landing buffers + the exchange state of one rank."""

    def __init__(self, group, arena_bytes: int, device, timeout_s: float = 30.0):
        lib = gpu.require()
        self.lib = lib
        self.group = group
        self.world = dist.get_world_size(group) if group is not None else 1
        self.rank = dist.get_rank(group) if group is not None else 0
        if self.world > MAXW:
            raise RuntimeError("reduce exchange supports at most %d ranks" % MAXW)
        self.device = torch.device(device)
        sizes = [c_int() for _ in range(4)]
        lib.bs_rx_sizes(*[ctypes.byref(s) for s in sizes])
        self.ctrl_bytes, self.state_bytes, self.report_bytes, peers_b = [s.value for s in sizes]
        if peers_b != ctypes.sizeof(RxPeers):
            raise RuntimeError("ctypes/CUDA struct layout mismatch (RxPeers %d/%d)"
                               % (peers_b, ctypes.sizeof(RxPeers)))
        offs = [c_ll() for _ in range(3)]
        lib.bs_rx_report_offsets(*[ctypes.byref(o) for o in offs])
        self.report_offs = tuple(o.value for o in offs)
        self.timeout_ns = int(timeout_s * 1e9)
        self.stream = torch.cuda.Stream(device=self.device)
        self.state = torch.zeros(self.state_bytes, dtype=torch.uint8, device=self.device)
        self.xring = _Ring(lib, self.report_bytes)  # exchange reports, slot = epoch % NSLOTS
        self.pring = _Ring(lib, self.report_bytes)  # local publishes, slot = ticket % NSLOTS
        self.epoch = 0          # host mirror of the device-side epoch
        self.publish_seq = 0    # host mirror of the device-side publish ticket
        self.pending: Dict[int, Pending] = {}
        self.ppending: Dict[int, Pending] = {}
        self.graphs: Dict[tuple, _Graph] = {}
        self.seen: Dict[tuple, int] = {}
        self.use_graphs = os.environ.get("BIGSLICE_NO_GRAPHS") is None
        self.stats = {"exchanges": 0, "graph_launches": 0, "graph_captures": 0, "direct": 0}
        self._ctrl = c_void_p()
        self._arena = c_void_p()
        self._opened: List[int] = []
        self.arena_bytes = 0
        self.peers = RxPeers()
        self._map(int(arena_bytes))

    # -- memory ------------------------------------------------------------------
    @staticmethod
    def _chk(rc: int, what: str):
        if rc != 0:
            raise RuntimeError("%s failed: cuda error %d" % (what, rc))

    def _map(self, arena_bytes: int) -> None:
        """Allocate control block + arena, exchange IPC handles (collective)."""
        lib = self.lib
        arena_bytes = max(arena_bytes, 1 << 20) & ~0xFFF
        h_ctrl = (ctypes.c_ubyte * 64)()
        h_arena = (ctypes.c_ubyte * 64)()
        self._chk(lib.bs_ipc_alloc(c_ll(self.ctrl_bytes), ctypes.byref(self._ctrl), h_ctrl), "ipc_alloc rx ctrl")
        self._chk(lib.bs_ipc_alloc(c_ll(arena_bytes), ctypes.byref(self._arena), h_arena), "ipc_alloc rx arena")
        self.arena_bytes = arena_bytes
        mine = (bytes(h_ctrl), bytes(h_arena), torch.cuda.current_device(), os.getpid())
        everyone: List = [mine]
        if self.world > 1:
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=self.group)
        p = self.peers
        p.world, p.rank = self.world, self.rank
        for r, (hc, ha, devidx, pid) in enumerate(everyone):
            if r == self.rank:
                p.ctrl[r], p.arena[r] = self._ctrl.value, self._arena.value
                continue
            if devidx != torch.cuda.current_device():
                lib.bs_enable_peer(c_int(devidx))
            pc, pa = c_void_p(), c_void_p()
            self._chk(lib.bs_ipc_open((ctypes.c_ubyte * 64).from_buffer_copy(hc), ctypes.byref(pc)),
                      "ipc_open rx ctrl of rank %d" % r)
            self._chk(lib.bs_ipc_open((ctypes.c_ubyte * 64).from_buffer_copy(ha), ctypes.byref(pa)),
                      "ipc_open rx arena of rank %d" % r)
            p.ctrl[r], p.arena[r] = pc.value, pa.value
            self._opened += [pc.value, pa.value]
        self._chk(lib.bs_rx_state_init(c_void_p(self.state.data_ptr()), c_ull(self.timeout_ns),
                                       gpu._stream()), "rx_state_init")
        self.epoch = 0
        self.publish_seq = 0
        if self.world > 1:
            dist.barrier(group=self.group)

    def _unmap(self) -> None:
        for ptr in self._opened:
            self.lib.bs_ipc_close(c_void_p(ptr))
        self._opened = []
        for ptr in (self._ctrl, self._arena):
            if ptr.value:
                self.lib.bs_dev_free(ptr)
        self._ctrl, self._arena = c_void_p(), c_void_p()

    def reset(self, arena_bytes: Optional[int] = None) -> None:
        """Collective: drain, drop every mapping and start over at epoch 0 —
        after a lost exchange (epochs may have diverged) or to grow the arena."""
        torch.cuda.synchronize(self.device)
        for g in self.graphs.values():
            self.lib.bs_graph_destroy(c_void_p(g.exec))
        self.graphs.clear()
        self.seen.clear()
        self.pending.clear()
        self.ppending.clear()
        if self.world > 1:
            dist.barrier(group=self.group)  # nobody still writes into a peer's arena
        self._unmap()
        self._map(int(arena_bytes or self.arena_bytes))

    def close(self) -> None:
        try:
            torch.cuda.synchronize(self.device)
        except Exception:
            pass
        for g in self.graphs.values():
            self.lib.bs_graph_destroy(c_void_p(g.exec))
        self.graphs.clear()
        self._unmap()
        self.xring.close()
        self.pring.close()

    # -- geometry ------------------------------------------------------------------
    def geometry(self, nparts: int, regions: int = 1) -> Tuple[int, int]:
        """(shards per rank, rows per landing region) for ``nparts`` shards whose
        destination tables are split into ``regions`` L2-sized parts."""
        pl = (nparts + self.world - 1) // self.world
        half = (self.arena_bytes // 2) & ~0xFF
        rows = half // (pl * regions * self.world * 16)
        return pl, rows

    def _peers_for(self, nparts: int, regions: int = 1) -> RxPeers:
        pl, rows = self.geometry(nparts, regions)
        p = self.peers
        p.shards_per_rank, p.region_rows, p.regions = pl * regions, rows, regions
        p.half_bytes = (self.arena_bytes // 2) & ~0xFF
        return p

    def bytes_for(self, nparts: int, region_rows: int, regions: int = 1) -> int:
        pl = (nparts + self.world - 1) // self.world
        # (_map rounds the arena down to 4 KiB: round up here so the regions really hold region_rows)
        return (2 * (pl * regions * self.world * 16 * int(region_rows) + 256) + 0xFFF) & ~0xFFF

    @staticmethod
    def regions_for(table_bytes: int, nparts: int, l2_part: int = 48 << 20) -> int:
        """Table regions per destination shard: the smallest power of two that makes
        a region's part of the table at most ``l2_part`` bytes (so combining on
        arrival runs out of L2), bounded by the kernel's 64 virtual partitions."""
        r = 1
        while table_bytes // r > l2_part and r < 8 and nparts * r * 2 <= MAXP:
            r *= 2
        return r

    # -- the exchange ----------------------------------------------------------------
    def _enqueue_send(self, pe: RxPeers, table, nparts: int, keymode: int = 0) -> None:
        lib, st = self.lib, c_void_p(self.state.data_ptr())
        s = gpu._stream()
        bits = c_ull(table.identity_bits())
        self._chk(lib.bs_rx_scatter(ctypes.byref(pe), st, c_void_p(table.slots.data_ptr()), c_ll(table.total),
                                    c_int(nparts), c_int(1), bits, c_void_p(table.status.data_ptr()),
                                    c_int(keymode), s), "rx_scatter")
        self._chk(lib.bs_rx_wait(ctypes.byref(pe), st, c_int(nparts), s), "rx_wait")

    def _enqueue(self, pe: RxPeers, table, nparts: int, dests, max_rows: int, keymode: int = 0) -> None:
        self._enqueue_send(pe, table, nparts, keymode)
        self._enqueue_recv(pe, table, nparts, dests, max_rows)

    def _enqueue_recv(self, pe: RxPeers, table, nparts: int, dests, max_rows: int) -> None:
        lib, st = self.lib, c_void_p(self.state.data_ptr())
        s = gpu._stream()
        bits = c_ull(table.identity_bits())
        vt, op = c_int(gpu.VTYPE[table.vdtype]), c_int(gpu.OPS[table.op])
        n = len(dests)
        stp = (c_void_p * max(n, 1))()
        caps = (c_ll * max(n, 1))()
        R = int(pe.regions)
        for i, (j, dt, ok, ov, cap) in enumerate(dests):
            if dt.total % R:
                raise RuntimeError("destination table of %d slots does not split into %d regions" % (dt.total, R))
            rcap = dt.total // R
            for r in range(R):  # region after region: each part of the table stays in L2
                self._chk(lib.bs_rx_combine(ctypes.byref(pe), st, c_int(j * R + r), vt, op,
                                            c_void_p(dt.slots.data_ptr() + r * rcap * 16), c_ll(rcap),
                                            c_void_p(dt.status.data_ptr()), c_ll(max_rows), s), "rx_combine")
            self._chk(lib.bs_table_compact(c_void_p(dt.slots.data_ptr()), c_ll(dt.total), vt,
                                           c_void_p(ok.data_ptr()), c_void_p(ov.data_ptr()),
                                           c_void_p(dt.status.data_ptr()), c_int(1), bits, c_ll(cap), s),
                      "table_compact")
            stp[i], caps[i] = dt.status.data_ptr(), cap
        self._chk(lib.bs_rx_end(ctypes.byref(pe), st, c_int(n), stp, caps, bits, self.xring.ptr,
                                c_int(NSLOTS), c_int(1), s), "rx_end")

    def exchange_probe(self, table, nparts: int, vote: Tuple[int, int] = (0, 0), keymode: int = 0,
                       regions: int = 1) -> Pending:
        """First half of an exchange whose destination sizes are unknown (first
        run of a slice): scatter + wait, then publish how many rows arrived for
        every local shard.  Call on the exchange stream; follow with
        :meth:`exchange_finish` after sizing the destination tables.  ``regions``
        must be what the peers use for this exchange (the landing layout depends on
        it), whichever half-path each of them takes."""
        if nparts * regions > MAXP:
            raise RuntimeError("reduce exchange supports at most %d (shards x regions)" % MAXP)
        pe = self._peers_for(nparts, regions)
        lib, st = self.lib, c_void_p(self.state.data_ptr())
        s = gpu._stream()
        self._chk(lib.bs_rx_set_vote(st, c_ull(vote[0]), c_ull(vote[1]), s), "rx_set_vote")
        self._enqueue_send(pe, table, nparts, keymode)
        pl = pe.shards_per_rank // regions
        stp, caps = (c_void_p * pl)(), (c_ll * pl)()
        self._chk(lib.bs_rx_end(ctypes.byref(pe), st, c_int(pl), stp, caps, c_ull(0), self.xring.ptr,
                                c_int(NSLOTS), c_int(0), s), "rx_end(probe)")
        ev = torch.cuda.Event()
        ev.record()
        self.stats["direct"] += 1
        return Pending(self, self.xring, self.epoch + 1, ev)

    def exchange_finish(self, table, nparts: int, dests: Sequence[tuple], on_report=None,
                        regions: int = 1) -> Pending:
        """Second half: combine on arrival, compaction, acknowledgement and the
        closing agreement (every rank reports the same verdict)."""
        pe = self._peers_for(nparts, regions)
        self._enqueue_recv(pe, table, nparts, dests, int(pe.region_rows) * self.world)
        ev = torch.cuda.Event()
        ev.record()
        self.epoch += 1
        self.stats["exchanges"] += 1
        p = Pending(self, self.xring, self.epoch, ev, on_report)
        old = self.pending.pop(self.epoch - NSLOTS + 1, None)
        if old is not None and old.report is None:
            old.resolve()
        self.pending[self.epoch] = p
        return p

    def exchange(self, table, nparts: int, dests: Sequence[tuple], vote: Tuple[int, int] = (0, 0),
                 on_report=None, keymode: int = 0, regions: int = 1) -> Pending:
        """Enqueue one exchange of ``table`` on the exchange stream (the caller
        has made that stream wait for the producers' combines).

        ``dests``: ``(j, dest_table, out_keys, out_vals, out_cap)`` for every
        local shard ``rank + j*world`` — all buffers allocated by the caller
        *before* this call (nothing is allocated while a graph is captured).
        Collective and ordered: every rank issues the same sequence."""
        if nparts * regions > MAXP:
            raise RuntimeError("reduce exchange supports at most %d (shards x regions)" % MAXP)
        for j, dt, ok, ov, cap in dests:
            if dt.warm:
                raise RuntimeError("destination tables have no warm level")
        pe = self._peers_for(nparts, regions)
        pl, rows = pe.shards_per_rank, int(pe.region_rows)
        max_rows = rows * self.world
        lib = self.lib
        with torch.cuda.stream(self.stream):
            s = gpu._stream()
            self._chk(lib.bs_rx_set_vote(c_void_p(self.state.data_ptr()), c_ull(vote[0]), c_ull(vote[1]), s),
                      "rx_set_vote")
            key = (table.slots.data_ptr(), table.total, table.status.data_ptr(), nparts, rows, table.vdtype,
                   table.op, keymode, regions) + tuple(
                (j, dt.slots.data_ptr(), dt.total, dt.status.data_ptr(), ok.data_ptr(), ov.data_ptr(), cap)
                for j, dt, ok, ov, cap in dests)
            g = self.graphs.get(key) if self.use_graphs else None
            if g is not None:
                self._chk(lib.bs_graph_launch(c_void_p(g.exec), c_ll(g.nodes), s), "graph_launch")
                self.stats["graph_launches"] += 1
            elif self.use_graphs and self.seen.get(key, 0) >= 1:
                # second occurrence of the same launch sequence: capture it
                self._chk(lib.bs_graph_begin(s), "graph_begin")
                try:
                    self._enqueue(pe, table, nparts, dests, max_rows, keymode)
                finally:
                    ex, nodes = c_void_p(), c_ll()
                    rc = lib.bs_graph_end(s, ctypes.byref(ex), ctypes.byref(nodes))
                self._chk(rc, "graph_end")
                if len(self.graphs) >= 16:  # pointers keep changing: stop hoarding executables
                    k0 = next(iter(self.graphs))
                    lib.bs_graph_destroy(c_void_p(self.graphs.pop(k0).exec))
                g = _Graph(ex.value, nodes.value, None)
                self.graphs[key] = g
                self.stats["graph_captures"] += 1
                self._chk(lib.bs_graph_launch(c_void_p(g.exec), c_ll(g.nodes), s), "graph_launch")
                self.stats["graph_launches"] += 1
            else:
                self._enqueue(pe, table, nparts, dests, max_rows, keymode)
                self.seen[key] = self.seen.get(key, 0) + 1
                if len(self.seen) > 256:
                    self.seen.clear()
                self.stats["direct"] += 1
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.epoch += 1
        self.stats["exchanges"] += 1
        p = Pending(self, self.xring, self.epoch, ev, on_report)
        # the slot of this epoch is about to be reused NSLOTS exchanges from now:
        # whoever still cares about the old report reads it first
        old = self.pending.pop(self.epoch - NSLOTS + 1, None)
        if old is not None and old.report is None:
            old.resolve()
        self.pending[self.epoch] = p
        return p

    def publish(self, table_status: torch.Tensor, out_cap: int, on_report=None) -> Pending:
        """After a purely local combine -> compact on the *current* stream: the
        row count / overflow verdict goes to pinned memory instead of a sync."""
        s = gpu._stream()
        self._chk(self.lib.bs_table_publish(c_void_p(table_status.data_ptr()), c_ll(out_cap),
                                            c_void_p(self.state.data_ptr()), self.pring.ptr, c_int(NSLOTS), s),
                  "table_publish")
        ev = torch.cuda.Event()
        ev.record()
        self.publish_seq += 1
        p = Pending(self, self.pring, self.publish_seq, ev, on_report)
        old = self.ppending.pop(self.publish_seq - NSLOTS + 1, None)
        if old is not None and old.report is None:
            old.resolve()
        self.ppending[self.publish_seq] = p
        return p

    def harvest(self) -> None:
        """Resolve, without blocking, everything that has already completed
        (feeds the cardinality hints while the pipeline keeps running)."""
        for d in (self.pending, self.ppending):
            for t in [t for t, p in d.items() if p.report is None and p.event.query()]:
                try:
                    d[t].resolve()
                except Exception:
                    pass
            for t in [t for t, p in d.items() if p.report is not None]:
                del d[t]
