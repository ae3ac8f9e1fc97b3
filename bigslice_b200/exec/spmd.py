"""SPMD executor: one process (rank) per GPU, shuffles over NVLink.

Every rank runs the same driver program, registers the same Funcs and compiles
the identical task graph (the compile environment is broadcast from rank 0, the
Func registries are compared at start).  Shard ``s`` is owned by rank
``s % world``; a rank executes only its own tasks.  A shuffle dependency is
served by one *group exchange* per producer group:

* **fused** (CUDA, default): the ``csrc/exchange.cu`` kernels hash, partition
  and store rows directly into the destination GPU's receive arena through
  NVLink peer mappings — no NCCL on the data path.  With a builtin combiner the
  producer side first hash-aggregates locally (``csrc/hash_combine.cu``) and
  the destination combines what it receives.
* **nccl** — local partition + ``all_to_all_single``: the measured baseline.
* **object** — codec-encoded frames over the control-plane (gloo) group: any
  column type (strings, lists, python objects) and the CPU test path.

Exchanges are collective, so they are issued by a single per-invocation thread
in a canonical order (producer groups sorted by graph depth, then name), which
is identical on all ranks; tasks themselves still run under the normal
evaluator with ``parallelism`` local procs.

This executor takes the place of the reference's cluster executor
(``exec/bigmachine.go``, ``exec/slicemachine.go``): the "machines" are the GPUs
of one NVSwitch box.  Like the reference's machine-combiner mode it does not
recover lost tasks: a failure on any rank aborts the invocation everywhere.
"""
from __future__ import annotations

import os
import pickle
import queue
import threading
import time
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .. import context as _context
from .. import errors, func as _func, slicefunc
from ..frame import Frame, concat
from ..frame import columns as C
from ..sliceio import EmptyReader, FrameReader, MultiReader, Reader, encode_frame
from ..slices import default_partitioner
from .combiner import combine_frame
from .invocation import InvocationGraph, fingerprint
from ..utils.spans import span

from .local import LocalExecutor, TaskBuffer, split_by_partition
from .task import Task, TaskState, iter_tasks


class _TaggedReader(FrameReader):
    combined_unique = False


def init_process_group(device: Optional[str] = None) -> None:
    """Initialise torch.distributed from the torchrun environment if needed."""
    if dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if "MASTER_PORT" not in os.environ:
        # single-process use (no torchrun): any free port will do
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        os.environ["MASTER_PORT"] = str(s.getsockname()[1])
        s.close()
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    use_cuda = (device or "").startswith("cuda") or (device is None and torch.cuda.is_available())
    if use_cuda:
        local = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
        torch.cuda.set_device(local % max(torch.cuda.device_count(), 1))
        # threads + pinned staging buffers stay on the GPU's NUMA node
        from ..utils import topology
        topology.bind_to_gpu(torch.cuda.current_device())
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    else:
        dist.init_process_group("gloo")


class SpmdExecutor(LocalExecutor):
    def __init__(self, parallelism: int = 1, device: Optional[str] = "auto",
                 shuffle: str = "auto", arena_bytes: int = 1 << 30, zero_copy: bool = False):
        if device == "auto":
            device = "cuda" if torch.cuda.is_available() else None
        if device == "cuda":
            init_process_group("cuda")
            device = "cuda:%d" % torch.cuda.current_device()
        super().__init__(parallelism, device)
        self.shuffle = shuffle
        self.arena_bytes = arena_bytes
        self.zero_copy = zero_copy
        self.rank = 0
        self.world = 1
        self.ctrl_group = None
        self.peers = None
        self.exchanged: Dict[Tuple[int, int], List[Frame]] = {}
        self.group_done: Dict[int, threading.Event] = {}
        self.invocations = InvocationGraph()
        self._pending_fp: List[Tuple[int, str, int, int]] = []
        self._coll_thread: Optional[threading.Thread] = None
        self._coll_q: Optional[queue.Queue] = None
        self.group_err: Dict[int, BaseException] = {}
        self._glock = threading.Lock()
        self.exchange_log: List[dict] = []
        self.group_tables: Dict[int, Any] = {}       # gid -> streaming HashTable
        self.card_hints: Dict[str, int] = {}         # slice name -> distinct keys seen

    def name(self) -> str:
        return "spmd"

    def _pid(self):
        return "rank%d" % self.rank

    # -- lifecycle -----------------------------------------------------------
    def start(self, session):
        super().start(session)
        init_process_group(str(self.device) if self.device is not None else "cpu")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if dist.get_backend() != "gloo":
            self.ctrl_group = dist.new_group(backend="gloo")
        else:
            self.ctrl_group = dist.group.WORLD
        # Func registries must match on every rank (slicemachine.go:689-702).
        locs = _func.func_locations()
        all_locs: List = [None] * self.world
        dist.all_gather_object(all_locs, locs, group=self.ctrl_group)
        for r, other in enumerate(all_locs):
            if other != all_locs[0]:
                diff = _func.func_locations_diff(all_locs[0], other)
                raise errors.fatal("bigslice.Func registry differs between rank 0 and rank %d:\n%s"
                                   % (r, "\n".join(d for d in diff if d)))
        self.registry_size = len(locs)
        if self.shuffle == "auto":
            # GPUs: fused peer-memory kernels.  CPU ranks: the library all-to-all over
            # gloo for tensor frames (every row crosses the wire once), codec-encoded
            # frames over the control plane for anything else.
            self.shuffle = "fused" if (self.device is not None and self.device.type == "cuda") else "gloo"
        if self.shuffle == "fused":
            from ..parallel.peer import PeerGroup
            self.peers = PeerGroup(self.ctrl_group, self.arena_bytes, self.device)

        def shutdown():
            if self._coll_thread is not None:
                self._coll_q.put(None)
                self._coll_thread.join(timeout=5)
                self._coll_thread = None
            if self.peers is not None:
                self.peers.close()
        return shutdown

    def check_registry(self) -> None:
        """Fail if Funcs were registered after the session started."""
        n = len(_func.func_locations())
        if n != self.registry_size:
            raise errors.fatal("bigslice.Func registry changed after session start (%d -> %d)"
                               % (self.registry_size, n))

    def owner(self, shard: int) -> int:
        return shard % self.world

    def owns(self, task: Task) -> bool:
        return self.owner(task.name.shard) == self.rank

    # -- compile consistency ---------------------------------------------------
    def compile_invocation(self, einv, do_compile):
        """Rank 0 compiles with a writable environment and broadcasts it; the
        other ranks compile against the frozen copy (CompileEnv semantics)."""
        if einv.inv.func >= self.registry_size:
            raise errors.fatal("bigslice.Func #%d was registered after the session started; "
                               "ranks cannot agree on it" % einv.inv.func)
        self.invocations.add(einv)
        if self.world == 1:
            return do_compile(einv)
        # Every rank must issue the same invocation (the reference's workers compile
        # exactly the driver's invocation, bigmachine.go:546-640): the fingerprint
        # rides on the next control-plane vote instead of costing a collective here.
        with self._glock:
            self._pending_fp.append((einv.index, einv.location, einv.inv.func,
                                     int(fingerprint(einv.inv)[:15], 16)))
        if not einv.uses_cache():
            return do_compile(einv)  # nothing in a CompileEnv to agree on
        if self.rank == 0:
            out = do_compile(einv)
            box = [pickle.dumps(einv.env)]
            with span("compile/broadcast_env"):
                dist.broadcast_object_list(box, src=0, group=self.ctrl_group)
            return out
        box = [None]
        with span("compile/broadcast_env"):
            dist.broadcast_object_list(box, src=0, group=self.ctrl_group)
        einv.env = pickle.loads(box[0])
        return do_compile(einv)

    def _vote(self, bad: int, upto: Optional[int] = None) -> int:
        """Control-plane agreement before an exchange: did every rank's producers
        succeed, and did every rank issue the same invocations since the last vote."""
        acc, fps = self._take_fingerprint(upto)
        flag = torch.tensor([bad, acc, -acc], dtype=torch.int64)
        with span("xchg/status_vote"):
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.ctrl_group)
        if int(flag[1]) != -int(flag[2]):
            where = ", ".join("inv%d at %s (func #%d)" % f[:3] for f in fps) or "an earlier invocation"
            raise errors.fatal("ranks issued different invocations: %s on rank %d does not match the "
                               "other ranks; SPMD programs must run identical Session.run sequences"
                               % (where, self.rank))
        return int(flag[0])

    def begin_invocation(self, roots: Sequence[Task]) -> None:
        """Start the collective thread that performs this invocation's group
        exchanges in canonical order."""
        groups: Dict[int, List[Task]] = {}
        depth: Dict[int, int] = {}

        def task_depth(t: Task, memo: Dict[int, int]) -> int:
            if id(t) in memo:
                return memo[id(t)]
            d = 0
            for dep in t.deps:
                for dt in dep.tasks():
                    dd = task_depth(dt, memo)
                    if dt.group is not None:
                        dd += 1
                    d = max(d, dd)
            memo[id(t)] = d
            return d
        memo: Dict[int, int] = {}
        seen: List[Task] = []
        iter_tasks(roots, seen.append)
        for t in seen:
            if t.group is not None:
                gid = id(t.group[0])
                if gid not in groups:
                    groups[gid] = t.group
                    depth[gid] = max(task_depth(x, memo) for x in t.group)
        todo = []
        with self._glock:
            for gid, g in groups.items():
                if gid not in self.group_done:
                    self.group_done[gid] = threading.Event()
                    todo.append((depth[gid], str(g[0].name), gid, g))
        todo.sort(key=lambda x: (x[0], x[1]))
        if not todo:
            return
        # One long-lived collective thread serves every invocation, in submission
        # order: exchanges of different invocations never interleave differently
        # on different ranks, and no thread is created per run.
        with self._glock:
            if self._coll_thread is None:
                self._coll_q = queue.Queue()
                self._coll_thread = threading.Thread(target=self._collective_main, daemon=True,
                                                     name="bigslice-spmd-collective")
                self._coll_thread.start()
        self._coll_q.put(todo)

    # -- the collective thread ---------------------------------------------------
    def _collective_main(self) -> None:
        if self.device is not None and self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        while True:
            todo = self._coll_q.get()
            if todo is None:
                return
            self._collective_loop(todo)

    def _collective_loop(self, todo) -> None:
        failed: Optional[BaseException] = None
        for _, _, gid, group in todo:
            ev = self.group_done[gid]
            if failed is None:
                try:
                    self._exchange_group(gid, group)
                except BaseException as e:  # noqa: BLE001
                    failed = e
            if failed is not None:
                self.group_err[gid] = failed
            ev.set()

    def _local_status(self, group: List[Task]) -> int:
        """Wait for this rank's producers; 0 ok, 1 failed."""
        for t in group:
            if not self.owns(t):
                continue
            st = t.wait_state(TaskState.OK)
            if st != TaskState.OK:
                return 1
        return 0

    def _exchange_group(self, gid: int, group: List[Task]) -> None:
        with span("xchg/wait_producers"):
            bad = self._local_status(group)
        head = group[0]
        nparts = head.num_partition
        typ = head.type.typ
        with self._glock:
            table = self.group_tables.pop(gid, None)
        # The table path is a static property of the task (same answer on every
        # rank, also on ranks whose shards were empty or whose producers failed).
        if self._table_path_ok(head) and (table is not None or self.world > 1):
            if table is None:
                from ..ops import gpu
                table = gpu.table_pool().get(1024, head.type.typ.out(1).torch_dtype,
                                             head.combiner.builtin, self.device)
                table.track(nparts)
            with span("xchg/table_total"):
                return self._exchange_table(gid, head, table, nparts, bad)
        if self.world > 1:
            bad = self._vote(bad, head.name.inv_index)
        if bad:
            raise errors.fatal("spmd: a producer of %s failed on some rank" % group[0].name)
        frames: List[Frame] = []
        for t in group:
            if self.owns(t):
                with self.lock:
                    buf = self.buffers.pop(id(t), None)
                if buf is not None:
                    frames.extend(buf.parts[0])
        f = concat(frames, typ, self.device) if frames else Frame.empty(typ, self.device)
        ctx = slicefunc.Context(None, 0, self.device)
        t0 = time.perf_counter()
        ev0 = self._cuda_event()
        with _context.use(ctx):
            if head.combiner is not None and len(f):
                f = combine_frame(f, head.combiner, ctx)
            mode, parts = self._shuffle(f, head, nparts)
            if head.combiner is not None:
                parts = {p: [combine_frame(concat(fr, typ, self.device), head.combiner, ctx)]
                         if fr else [] for p, fr in parts.items()}
        with self._glock:
            for p in range(self.rank, nparts, self.world):
                self.exchanged[(gid, p)] = [x for x in parts.get(p, []) if len(x)]
        self._log_exchange(head, mode, t0, ev0, rows_in=len(f))

    # -- fast Reduce path: streaming hash-aggregation, table shuffled in place --
    def _table_path_ok(self, task: Task) -> bool:
        if self.device is None or self.device.type != "cuda" or task.combiner is None:
            return False
        if task.partitioner is not default_partitioner:
            return False
        if self.peers is None and not (self.world == 1 and task.num_partition == 1) \
                and self.shuffle != "nccl-table":
            return False  # the table is shuffled by the peer-memory kernels (or, as the
            #               like-for-like baseline, by the library all-to-all: "nccl-table")
        from ..ops import gpu
        typ = task.type.typ
        if task.combiner.builtin not in gpu.OPS or typ.num_out() != 2 or typ.prefix() != 1:
            return False
        return (typ.out(0).torch_dtype == torch.int64 and typ.out(0).kind == "int"
                and typ.out(1).torch_dtype in gpu.VTYPE and typ.out(1).kind != "vec"
                and task.num_partition <= 256)

    def _hint_key(self, task: Task) -> str:
        return "%s/%s" % (task.type.name(), task.combiner.builtin)

    def _group_table(self, task: Task, rows_hint: int):
        """The (per rank, per producer group) HBM table all local producer
        tasks aggregate into; sized from the cardinality seen last time."""
        from ..ops import gpu
        gid = id(task.group[0])
        with self._glock:
            t = self.group_tables.get(gid)
            if t is None:
                hint = self.card_hints.get(self._hint_key(task))
                cap = gpu.table_capacity(hint) if hint is not None else 2 * max(rows_hint, 1 << 16)
                # pooled: a table emptied by its last reader (clearing exchange or
                # compaction) is reused without an init pass
                t = gpu.table_pool().get(cap, task.type.typ.out(1).torch_dtype, task.combiner.builtin,
                                         self.device)
                if not (self.world == 1 and task.num_partition == 1):
                    t.track(task.num_partition)  # destinations counted while combining
                self.group_tables[gid] = t
            return t

    def _take_fingerprint(self, upto: Optional[int] = None) -> Tuple[int, List]:
        """Fingerprints of the invocations not yet agreed on, up to and including
        invocation ``upto`` — the one whose exchange carries the vote.  (The main
        thread may already have compiled later invocations, e.g. after this one
        failed locally; those belong to later votes on every rank.)"""
        with self._glock:
            fps = [f for f in self._pending_fp if upto is None or f[0] <= upto]
            self._pending_fp = [f for f in self._pending_fp if not (upto is None or f[0] <= upto)]
        acc = 0
        for _, _, _, fp in fps:
            acc = (acc * 1000003 + fp) & ((1 << 61) - 1)
        return acc, fps

    def _exchange_table(self, gid: int, head: Task, table, nparts: int, bad: int = 0) -> None:
        from ..ops import gpu
        from ..parallel.peer import ExchangeVoteError
        typ = head.type.typ
        t0 = time.perf_counter()
        ev0 = self._cuda_event()
        out: Dict[int, List[Frame]] = {}
        if self.world == 1 and bad:
            raise errors.fatal("spmd: a producer of %s failed" % head.name)
        if self.shuffle == "nccl-table" and not (self.world == 1 and nparts == 1):
            return self._exchange_table_library(gid, head, table, nparts, bad, t0, ev0)
        if self.world == 1 and nparts == 1:
            # the local table is already the answer
            clear = gpu.RECYCLE == "clear"
            k, v = table.compact(clear=clear)
            gpu.table_pool().put(table, clean=clear)
            out[0] = [Frame([k, v], typ)]
            sent = len(k)
            mode = "table-local"
        else:
            # the producer status and the invocation fingerprint travel with the
            # counts: all ranks agree on device, without a host collective
            acc, fps = self._take_fingerprint(head.name.inv_index)
            try:
                with span("xchg/exchange_table"):
                    got = self.peers.exchange_table(table, nparts, clear=gpu.RECYCLE == "clear",
                                                    vote=(1 if bad else 0, acc))
            except ExchangeVoteError as e:
                if e.kind == 3:
                    raise errors.fatal("spmd: a producer of %s failed on some rank" % head.name) from e
                where = ", ".join("inv%d at %s (func #%d)" % f[:3] for f in fps) or "an earlier invocation"
                raise errors.fatal("ranks issued different invocations (%s, rank %d); SPMD programs must "
                                   "run identical Session.run sequences" % (where, self.rank)) from e
            with span("xchg/status_sync"):
                st = table.status.tolist()
            if st[0]:
                self.card_hints[self._hint_key(head)] = 4 * max(self.card_hints.get(self._hint_key(head), 0), 1 << 16)
                raise errors.fatal("spmd: combine table overflow (%d rows): the number of distinct keys "
                                   "more than doubled since the last run of this slice; the estimate has "
                                   "been raised, re-run the invocation" % st[0])
            gpu.table_pool().put(table, clean=gpu.RECYCLE == "clear")
            sent = 0
            for p, (rk, rv) in got.items():
                sent += rk.numel()
                if rk.numel() == 0:
                    out[p] = []
                    continue
                # rows from different ranks may share keys: combine on arrival
                dkey = self._hint_key(head) + "/dest%d" % p
                dh = self.card_hints.get(dkey)
                with span("xchg/dest_combine"):
                    fk, fv = gpu.hash_combine(rk, rv, head.combiner.builtin,
                                              distinct_hint=min(rk.numel(), dh) if dh else rk.numel())
                self.card_hints[dkey] = max(int(fk.numel() * 1.15), 1 << 12)
                out[p] = [Frame([fk, fv], typ)]
            if st[1]:  # the reserved EMPTY key was present: route it on the host (rare)
                self._route_special(head, st[2], nparts, out, typ)
            mode = "fused-table"
        self.card_hints[self._hint_key(head)] = max(int(sent * 1.15), 1 << 12) if self.world == 1 \
            else max(int(self.peers.last_sent * 1.15), 1 << 12)
        with self._glock:
            for p in range(self.rank, nparts, self.world):
                self.exchanged[(gid, p)] = [x for x in out.get(p, []) if len(x)]
        self._log_exchange(head, mode, t0, ev0)

    def _exchange_table_library(self, gid: int, head: Task, table, nparts: int, bad: int, t0, ev0) -> None:
        """Like-for-like baseline of the Reduce shuffle (``shuffle="nccl-table"``):
        the same combine kernels on both sides, but the table is compacted and its
        rows travel through ``all_to_all_single`` (NCCL) instead of the peer-memory
        scatter.  Everything the fused path fuses shows up here as separate passes:
        compaction, hashing, local partition, three library collectives."""
        from ..ops import gpu
        from ..parallel.peer import collective_exchange_tensors
        typ = head.type.typ
        if self.world > 1:
            bad = self._vote(bad, head.name.inv_index)
        if bad:
            raise errors.fatal("spmd: a producer of %s failed on some rank" % head.name)
        k, v = table.compact()
        st = table.status.tolist()
        gpu.table_pool().put(table, clean=False)
        dest = gpu.hash_column(k) % nparts
        if self.world > 1:
            k, v, dest = collective_exchange_tensors([k, v, dest], dest % self.world, None)
        out: Dict[int, List[Frame]] = {}
        for p in range(self.rank, nparts, self.world):
            sel = dest == p if nparts > self.world else None
            pk, pv = (k[sel], v[sel]) if sel is not None else (k, v)
            if pk.numel() == 0:
                out[p] = []
                continue
            fk, fv = gpu.hash_combine(pk, pv, head.combiner.builtin, distinct_hint=pk.numel())
            out[p] = [Frame([fk, fv], typ)]
        if st[1]:
            self._route_special(head, st[2], nparts, out, typ)
        self.card_hints[self._hint_key(head)] = max(int(k.numel() * 1.15), 1 << 12)
        with self._glock:
            for p in range(self.rank, nparts, self.world):
                self.exchanged[(gid, p)] = [x for x in out.get(p, []) if len(x)]
        self._log_exchange(head, "nccl-table", t0, ev0)

    def _route_special(self, head: Task, bits: int, nparts: int, out, typ) -> None:
        """A row with the reserved sentinel key: all ranks contribute their
        partial over the control plane and its owner adds the row."""
        import struct
        from ..frame import hashing
        key = -(1 << 63)
        vt = typ.out(1).torch_dtype
        raw = struct.pack("<q", bits if bits < (1 << 63) else bits - (1 << 64))
        val = torch.frombuffer(bytearray(raw), dtype=vt)[0].item()
        vals = [None] * self.world
        dist.all_gather_object(vals, val, group=self.ctrl_group)
        d = int(hashing.hash64(torch.tensor([key]), 0)[0]) % nparts
        if self.owner(d) != self.rank:
            return
        op = head.combiner.builtin
        tot = vals[0]
        for x in vals[1:]:
            tot = tot + x if op == "add" else (min(tot, x) if op == "min" else max(tot, x))
        out.setdefault(d, []).append(Frame([torch.tensor([key], device=self.device),
                                            torch.tensor([tot], dtype=vt, device=self.device)], typ))

    def _cuda_event(self):
        if self.device is not None and self.device.type == "cuda":
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            return ev
        return None

    def _log_exchange(self, head: Task, mode: str, t0: float, ev0, **kw) -> None:
        """Record the exchange in the log and, device-timed with CUDA events,
        in the session trace (pid = rank, category "exchange")."""
        rec = {"group": str(head.name), "mode": mode, "seconds": time.perf_counter() - t0}
        rec.update(kw)
        if ev0 is not None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            ev1.synchronize()
            rec["device_ms"] = ev0.elapsed_time(ev1)
        self.exchange_log.append(rec)
        if self.tracer is not None:
            now = self.tracer._now_us()
            dur = int(rec.get("device_ms", rec["seconds"] * 1e3) * 1e3)
            self.tracer.complete(self._pid(), "exchange:%s[%s]" % (head.name, mode), now - dur, dur,
                                 cat="exchange", rows_in=kw.get("rows_in", 0))

    def _fused_ok(self, f: Frame, head: Task, nparts: int) -> bool:
        from ..parallel import peer
        if self.peers is None or nparts > peer.MAXP or f.num_out() > peer.MAXC or f.num_out() == 0:
            return False
        for c in f.cols():
            if not (isinstance(c, torch.Tensor) and c.is_cuda and c.dtype != torch.bool):
                return False
        if head.partitioner is default_partitioner:
            if f.prefix() > peer.MAXK:
                return False
            for k in range(f.prefix()):
                c = f.col(k)
                if c.dim() != 1 or c.element_size() not in (4, 8):
                    return False
        return True

    def _shuffle(self, f: Frame, head: Task, nparts: int):
        """Returns (mode, {local shard: [frames]})."""
        typ = f.typ
        if self.world == 1 and self.peers is None:
            ids = head.partitioner(f, nparts) if len(f) else torch.zeros(0, dtype=torch.int64)
            parts = split_by_partition(f, ids, nparts) if len(f) else [f] * nparts
            return "local", {p: [parts[p]] for p in range(nparts)}
        if self.shuffle == "fused" and self._fused_ok(f, head, nparts):
            ids = None
            if head.partitioner is not default_partitioner:
                ids = head.partitioner(f, nparts).to(self.device)
            got = self.peers.exchange(f.cols(), nparts, nkeys=f.prefix(), ids=ids,
                                      zero_copy=self.zero_copy)
            return "fused", {p: [Frame(cols, typ)] for p, cols in got.items()}
        ids = head.partitioner(f, nparts) if len(f) else torch.zeros(0, dtype=torch.int64)
        tensor_only = all(isinstance(c, torch.Tensor) for c in f.cols()) and f.num_out() > 0
        if self.shuffle in ("nccl", "fused") and tensor_only and f.device.type == "cuda" \
                and nparts <= self.world * 64:
            return "nccl", self._shuffle_collective(f, ids, nparts)
        if self.shuffle == "gloo" and tensor_only and f.device.type == "cpu" and self.world > 1 \
                and dist.get_backend() == "gloo":
            return "gloo", self._shuffle_collective(f, ids, nparts)
        return "object", self._shuffle_objects(f, ids, nparts)

    def _shuffle_collective(self, f: Frame, ids: torch.Tensor, nparts: int):
        """Baseline: library all-to-all of each column (plus the shard id)."""
        from ..parallel.peer import collective_exchange_tensors
        ids = ids.to(f.device)
        cols = f.cols() + [ids]
        got = collective_exchange_tensors(cols, ids % self.world, None)
        rid = got[-1]
        g = Frame(got[:-1], f.typ)
        out = {}
        for p in range(self.rank, nparts, self.world):
            out[p] = [g.mask(rid == p)]
        return out

    def _shuffle_objects(self, f: Frame, ids: torch.Tensor, nparts: int):
        """Any column type: codec-encoded per-shard frames over the control
        plane (all_gather_object)."""
        parts = split_by_partition(f, ids, nparts) if len(f) else []
        mine = {}
        for p, g in enumerate(parts):
            if len(g) and self.owner(p) != self.rank:
                mine[p] = encode_frame(g.to("cpu") if g.device.type != "cpu" else g)
        allsends: List = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(allsends, mine, group=self.ctrl_group)
        out: Dict[int, List[Frame]] = {}
        for p in range(self.rank, nparts, self.world):
            frs = []
            if parts and len(parts[p]):
                frs.append(parts[p])
            for r, sends in enumerate(allsends):
                if r == self.rank or not sends or p not in sends:
                    continue
                b = sends[p]
                frs.append(_decode_one(b, f.typ, self.device))
            out[p] = frs
        return out

    # -- task execution ---------------------------------------------------------
    def run(self, task: Task) -> None:
        if not self.owns(task):
            task.set(TaskState.RUNNING)
            task.set(TaskState.OK)  # executed by its owner rank
            return
        super().run(task)
        if task.state() == TaskState.LOST:
            # no lineage recovery across ranks: surface as a fatal error
            task.error(errors.fatal("spmd: task %s lost (no recovery in SPMD mode)" % task.name))

    def dep_readers(self, task: Task) -> List[Reader]:
        ins: List[Reader] = []
        for dep in task.deps:
            if dep.head.group is None:
                dt = dep.head
                if dt.state() != TaskState.OK:
                    raise errors.TaskLostError("dependency %s is not available" % dt.name)
                ins.append(self.reader(dt, dep.partition))
                continue
            gid = id(dep.head.group[0])
            ev = self.group_done.get(gid)
            if ev is None:
                raise errors.fatal("spmd: no exchange scheduled for %s" % dep.head.name)
            ev.wait()
            if gid in self.group_err:
                raise errors.fatal("spmd: exchange for %s failed" % dep.head.name, self.group_err[gid])
            with self._glock:
                frames = self.exchanged.get((gid, dep.partition), [])
            typ = dep.head.type.typ
            r: Reader
            if frames:
                r = _TaggedReader(concat(frames, typ, self.device))
            else:
                r = EmptyReader()
            if dep.head.combiner is not None:
                r.combined_unique = True
            ins.append(r)
        return ins

    def buffer_output(self, task: Task, out: Reader, ctx) -> TaskBuffer:
        if task.group is None:
            return super().buffer_output(task, out, ctx)
        # Shuffle producer: keep whole tiles; the group exchange partitions (and
        # combines) them in one fused pass.
        buf = TaskBuffer(1)
        nrec = 0
        table = None
        fast = self._table_path_ok(task)
        try:
            while True:
                with span("task/read"):
                    f = out.read(None)
                if f is None:
                    break
                nrec += len(f)
                if fast and len(f):
                    if table is None:
                        hint_rows = 0
                        src = task.slices[-1] if task.slices else None
                        rh = getattr(src, "rows_hint", None)
                        if rh is not None:
                            hint_rows = int(rh(task.name.shard))
                        with span("task/group_table"):
                            table = self._group_table(task, max(hint_rows, len(f)))
                    with span("task/combine_launch"):
                        table.combine(f.col(0), f.col(1))  # async: overlaps the next read
                else:
                    buf.append(0, f)
        finally:
            out.close()
        task.stats = {"inrecords": nrec, "write": nrec}
        return buf

    def reader(self, task: Task, partition: int) -> Reader:
        if not self.owns(task):
            return EmptyReader()
        return super().reader(task, partition)

    def discard(self, task: Task) -> None:
        with self.lock:
            self.buffers.pop(id(task), None)
        if task.group is not None and task is task.group[0]:
            gid = id(task)
            with self._glock:
                for k in [k for k in self.exchanged if k[0] == gid]:
                    del self.exchanged[k]
                self.group_done.pop(gid, None)
                self.group_err.pop(gid, None)

    # -- results -------------------------------------------------------------------
    def result_reader(self, result) -> Reader:
        """Rows of the shards this rank owns, in shard order (SPMD view)."""
        return MultiReader([self.reader(t, 0) for t in result.tasks if self.owns(t)])

    def gather(self, result) -> List[tuple]:
        """All rows of ``result`` on every rank, in shard order."""
        local = {}
        for t in result.tasks:
            if self.owns(t):
                rows = []
                for f in self.reader(t, 0):
                    rows.extend(f.to("cpu").rows())
                local[t.name.shard] = rows
        everyone: List = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(everyone, local, group=self.ctrl_group)
        else:
            everyone = [local]
        merged = {}
        for d in everyone:
            merged.update(d)
        out: List[tuple] = []
        for s in sorted(merged):
            out.extend(merged[s])
        return out


def _decode_one(b: bytes, typ, device) -> Frame:
    import io
    from ..sliceio import Decoder, read_all
    return read_all(Decoder(io.BytesIO(b), typ, device), typ, device)
