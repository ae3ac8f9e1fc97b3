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

The Reduce fast path is *asynchronous*: producers hash-aggregate into an HBM
table, the collective thread enqueues one CUDA graph (table scatter over NVLink,
combine on arrival, compaction; ``parallel/rx.py``) on the exchange stream and
hands the consumers :class:`~bigslice_b200.frame.deferred.DeferredFrame`s whose
row counts are read from pinned memory only when somebody needs them.  The
per-exchange agreement between ranks (producer status, invocation fingerprint)
is a shared-memory status board (``utils/shmboard.py``), not a collective.

This executor takes the place of the reference's cluster executor
(``exec/bigmachine.go``, ``exec/slicemachine.go``): the "machines" are the GPUs
of one NVSwitch box.  Failure model: a *fatal* producer error on any rank fails
the invocation on every rank; anything else — a producer that failed
transiently, a table or landing-region overflow, a peer that timed out — marks
the exchange *lost*: its producers are re-run and the exchange is repeated
(``exec/bigmachine.go:441-454``), up to the evaluator's consecutive-loss limit.
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
        if os.environ.get("BIGSLICE_DIST_BACKEND") == "gloo":
            # several ranks may share one device (tests on a single-GPU box): NCCL
            # refuses duplicate devices, the bootstrap needs nothing but gloo
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    else:
        dist.init_process_group("gloo")


class SpmdExecutor(LocalExecutor):
    def __init__(self, parallelism: int = 1, device: Optional[str] = "auto",
                 shuffle: str = "auto", arena_bytes: int = 1 << 30, zero_copy: bool = False,
                 exchange_timeout: float = 30.0):
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
        self.card_hints: Dict[tuple, int] = {}       # stable slice key -> cardinality seen last time
        self.rx = None                               # ReduceExchange (lazily created, collective)
        self.board = None                            # shared-memory status board (world > 1)
        self.xseq = 0                                # exchanges posted on the board
        self.exchange_timeout = float(os.environ.get('BIGSLICE_EXCHANGE_TIMEOUT', exchange_timeout))
        self._want_arena = 0                         # landing overflow seen: grow before the next exchange
        self._rx_reset = False                       # a peer timed out: re-establish epochs first
        self._lost_groups: set = set()
        self._cancels: Dict[int, threading.Event] = {}
        self._dest_tables: Dict[tuple, Any] = {}     # (hint key, capacity) -> destination table
        self._out_sets: Dict[tuple, list] = {}       # hint key -> [(keys, vals, weakref to the frame)]
        self.group_cancel: Dict[int, threading.Event] = {}
        self._attempts: Dict[tuple, int] = {}        # (invocation, group name) -> exchanges issued
        self._group_pend: Dict[int, Any] = {}        # gid -> outcome handle of the exchange in place
        self.fault_hook = None                       # tests: called as hook(stage, group_name)

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
        if self.world > 1:
            from ..utils.shmboard import Board
            name = "bigslice-board-%s-%d" % (os.environ.get("MASTER_PORT", "0"), session.index)
            if self.rank == 0:
                self.board = Board(name, self.world, 0, create=True)
            dist.barrier(group=self.ctrl_group)
            if self.rank != 0:
                self.board = Board(name, self.world, self.rank, create=False)
            dist.barrier(group=self.ctrl_group)

        def shutdown():
            if self._coll_thread is not None:
                self._coll_q.put(None)
                self._coll_thread.join(timeout=5)
                self._coll_thread = None
            if self.rx is not None:
                self.rx.close()
                self.rx = None
            if self.peers is not None:
                self.peers.close()
            if self.board is not None:
                self.board.close(unlink=self.rank == 0)
                self.board = None
        return shutdown

    def check_registry(self) -> None:
        """Fail if Funcs were registered after the session started."""
        n = len(_func.func_locations())
        if n != self.registry_size:
            raise errors.fatal("bigslice.Func registry changed after session start (%d -> %d)"
                               % (self.registry_size, n))

    def health(self) -> dict:
        """This rank's "machine" line: HBM, exchange counters, pool state (the
        reference reports mem/disk/load per machine, exec/slicemachine.go:238-257)."""
        rec = {"rank": self.rank, "world": self.world, "device": str(self.device),
               "exchanges": len(self.exchange_log), "hints": len(self.card_hints)}
        if self.device is not None and self.device.type == "cuda":
            free, total = torch.cuda.mem_get_info(self.device)
            rec.update(hbm_used_gib=round((total - free) / 2**30, 2), hbm_total_gib=round(total / 2**30, 2),
                       hbm_allocated_gib=round(torch.cuda.memory_allocated(self.device) / 2**30, 2))
        if self.rx is not None:
            rec.update(rx=dict(self.rx.stats), rx_arena_mib=self.rx.arena_bytes >> 20)
        return rec

    def status_lines(self) -> List[str]:
        h = self.health()
        mem = ""
        if "hbm_total_gib" in h:
            mem = " hbm %.1f/%.1f GiB" % (h["hbm_used_gib"], h["hbm_total_gib"])
        rx = ""
        if "rx" in h:
            rx = " exchanges %d (graph %d)" % (h["rx"]["exchanges"], h["rx"]["graph_launches"])
        return ["rank%d/%d %s%s%s" % (h["rank"], h["world"], h["device"], mem, rx)]

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

    def _group_depths(self, roots: Sequence[Task]):
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
        return groups, depth

    def begin_invocation(self, roots: Sequence[Task]) -> None:
        """Queue this invocation's group exchanges, in canonical order, for the
        collective thread."""
        groups, depth = self._group_depths(roots)
        cancel = threading.Event()
        todo = []
        with self._glock:
            self._cancels[id(roots[0]) if roots else 0] = cancel
            for gid, g in groups.items():
                if gid not in self.group_done:
                    self.group_done[gid] = threading.Event()
                    self.group_cancel[gid] = cancel
                    todo.append((depth[gid], str(g[0].name), gid, g))
        todo.sort(key=lambda x: (x[0], x[1]))
        if todo:
            self._submit(todo)

    def end_invocation(self, roots: Sequence[Task], failed: bool) -> None:
        """Called by the session when ``evaluate`` returns or raises: a failed
        evaluation releases the collective thread from producers that will never
        be submitted (it still takes part in the vote, with a bad status)."""
        with self._glock:
            cancel = self._cancels.pop(id(roots[0]) if roots else 0, None)
        if failed and cancel is not None:
            cancel.set()

    def _submit(self, todo) -> None:
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

    def _ensure_group(self, task: Task) -> None:
        """A producer of a group whose exchange is not (or no longer) scheduled is
        being run: the group was lost and is re-evaluated.  Every rank runs every
        task of the graph (tasks of other ranks as no-ops), so all ranks queue the
        repeated exchange at the same point of their sequence."""
        gid = id(task.group[0])
        with self._glock:
            if gid in self.group_done:
                return
            self.group_done[gid] = threading.Event()
            self.group_err.pop(gid, None)
            self.group_cancel[gid] = threading.Event()
        self._submit([(0, str(task.group[0].name), gid, task.group)])

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
        for _, _, gid, group in todo:
            ev = self.group_done.get(gid)
            try:
                self._exchange_group(gid, group)
            except errors.Error as e:
                # every rank still takes part in the later exchanges of this
                # invocation (with whatever status its producers reach): skipping
                # them here would leave the peers waiting
                self.group_err[gid] = e
            except BaseException as e:  # noqa: BLE001
                if _is_comm_failure(e):
                    # a peer went away mid-collective: the exchange is lost (it is
                    # retried, and gives up with TooManyTries if the peer stays gone)
                    self._rx_reset = True
                    self.group_err[gid] = errors.TaskLostError(
                        "spmd: exchange of %s lost a peer: %s" % (group[0].name, str(e)[:200]))
                    if ev is not None:
                        ev.set()
                    continue
                import traceback  # anything else is a bug or a CUDA failure, not a lost task
                self.group_err[gid] = errors.fatal("spmd: exchange of %s raised %s: %s\n%s" % (
                    group[0].name, type(e).__name__, e, traceback.format_exc()), e)
            if ev is not None:
                ev.set()

    def _local_status(self, gid: int, group: List[Task]) -> int:
        """Wait for this rank's producers: 0 all OK; 1 one failed for good (ERR)
        or the evaluation was abandoned.  LOST producers are being retried by the
        evaluator — keep waiting for them."""
        cancel = self.group_cancel.get(gid)
        for t in group:
            if not self.owns(t):
                continue
            while True:
                st = t.wait_state(TaskState.OK, timeout=0.05)
                if st == TaskState.OK:
                    break
                if st == TaskState.ERR:
                    return 1
                if cancel is not None and cancel.is_set():
                    return 1
        return 0

    def _exchange_group(self, gid: int, group: List[Task]) -> None:
        with span("xchg/wait_producers"):
            bad = self._local_status(gid, group)
        head = group[0]
        nparts = head.num_partition
        typ = head.type.typ
        if self.fault_hook is not None:
            self.fault_hook("before_exchange", str(head.name))
        with self._glock:
            table = self.group_tables.pop(gid, None)
        # The table path is a static property of the task (same answer on every
        # rank, also on ranks whose shards were empty or whose producers failed).
        if self._table_path_ok(head):
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
        from ..cogroup import broadcast_partitioner
        if head.partitioner is broadcast_partitioner:
            mode, g = self._replicate(f)
            with self._glock:
                for p in range(self.rank, nparts, self.world):
                    self.exchanged[(gid, p)] = [g] if len(g) else []
            return self._log_exchange(head, mode, t0, ev0, rows_in=len(f))
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
        if self.shuffle not in ("fused", "nccl-table"):
            return False
        from ..ops import gpu, keypack
        from ..parallel import rx
        # the planner decided the shape (exec/compile.py: _table_lowering); the
        # executor adds what depends on the device library
        if task.lowering != "reduce-table":
            return False
        typ = task.type.typ
        if self.shuffle == "nccl-table" and not keypack.layout_for(typ).plain_int64:
            return False  # the library baseline hashes plain int64 keys only
        return typ.out(typ.prefix()).torch_dtype in gpu.VTYPE and task.num_partition <= rx.MAXP

    def _hint_key(self, task: Task, what: str = "local") -> tuple:
        """Cardinality hints are keyed on the slice's *creation site* — not on its
        name, whose per-site index changes with every ``Session.run`` (each run
        re-executes the Func body) — plus the combiner and the shard count."""
        nm = task.type.name()
        return (getattr(nm, "op", str(nm)), getattr(nm, "file", ""), getattr(nm, "line", 0),
                task.combiner.builtin, task.num_partition, what)

    def _group_table(self, task: Task, rows_hint: int):
        """The (per rank, per producer group) HBM table all local producer tasks
        aggregate into: sized from the cardinality seen last time, else for the
        worst case (every row of every local producer a distinct key)."""
        from ..ops import gpu
        gid = id(task.group[0])
        with self._glock:
            t = self.group_tables.get(gid)
            if t is None:
                hint = self.card_hints.get(self._hint_key(task))
                if hint is not None:
                    cap = gpu.table_capacity(int(hint * 1.15) + 1024)
                else:
                    rows = 0
                    for g in task.group:
                        if not self.owns(g):
                            continue
                        src = g.slices[-1] if getattr(g, "slices", None) else None
                        rh = getattr(src, "rows_hint", None)
                        rows += int(rh(g.name.shard)) if rh is not None else max(rows_hint, 1)
                    cap = gpu.table_capacity(max(rows, rows_hint, 1 << 15))
                # pooled: a table emptied by its last reader (clearing exchange or
                # compaction) is reused without an init pass
                typ = task.type.typ
                t = gpu.table_pool().get(cap, typ.out(typ.prefix()).torch_dtype, task.combiner.builtin,
                                         self.device)
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

    def _tag_exchange(self, acc: int, head: Task) -> int:
        """Fold *which* exchange this is — invocation, group, attempt — into the
        fingerprint the ranks compare.  Ranks whose exchange sequences diverged (one
        of them retrying a group the others consider done) then fail the vote on
        both sides instead of pairing two different exchanges."""
        import zlib
        key = (head.name.inv_index, str(head.name))
        with self._glock:
            n = self._attempts[key] = self._attempts.get(key, 0) + 1
            if len(self._attempts) > 4096:
                for k in list(self._attempts)[:2048]:
                    del self._attempts[k]
        tag = zlib.crc32(("%d/%s/%d" % (key[0], key[1], n)).encode())
        return (acc * 1000003 + tag) & ((1 << 61) - 1)

    def _rx(self):
        """The Reduce exchange state (created collectively on first use: every
        rank takes the table path for the same groups)."""
        if self.rx is None:
            from ..parallel.rx import ReduceExchange
            self.rx = ReduceExchange(self.ctrl_group if self.world > 1 else None,
                                     self.arena_bytes if self.world > 1 else 1 << 20, self.device,
                                     timeout_s=self.exchange_timeout)
        return self.rx

    def _peers_alive(self) -> None:
        """Before a collective re-initialisation: every rank checks in on the
        status board (bounded wait) so that a dead peer yields a lost task, not a
        hang inside a barrier."""
        if self.board is None:
            return
        self.xseq += 1
        self.board.post(self.xseq, 0, -1)
        votes = self.board.wait_all(self.xseq, self.exchange_timeout)
        if votes is None:
            raise errors.TaskLostError("spmd: rank(s) %s unreachable" % self.board.missing(self.xseq))
        others = [r for r, v in enumerate(votes) if v[1] != -1]
        if others:
            # they are voting on an ordinary exchange: entering the collective reset now
            # would leave this rank in a barrier nobody else joins
            raise errors.fatal("spmd: exchange sequences diverged: rank %d re-establishes the exchange after a "
                               "lost one while rank(s) %s carry on with the next" % (self.rank, others))

    def _board_vote(self, bad: int, acc: int, fps, head: Task, extra: int = 0) -> List[int]:
        """Host-side agreement for one exchange: post my status + fingerprint (and
        one free integer) and read everyone else's from shared memory (no
        collective, no GPU).  Returns every rank's integer."""
        if self.board is None:
            if bad:
                raise errors.fatal("spmd: a producer of %s failed" % head.name)
            return [extra]
        self.xseq += 1
        self.board.post(self.xseq, int(bad), acc, int(extra))
        with span("xchg/board_vote"):
            votes = self.board.wait_all(self.xseq, self.exchange_timeout)
        if votes is None:
            self._rx_reset = True
            raise errors.TaskLostError("spmd: rank(s) %s did not reach the exchange of %s within %.0f s"
                                       % (self.board.missing(self.xseq), head.name, self.exchange_timeout))
        if any(v[0] for v in votes):
            raise errors.fatal("spmd: a producer of %s failed on some rank" % head.name)
        if any(v[1] != acc for v in votes):
            where = ", ".join("inv%d at %s (func #%d)" % f[:3] for f in fps) or "an earlier invocation"
            raise errors.fatal("ranks issued different invocations, or their exchange sequences diverged "
                               "(%s; rank %d is at the exchange of %s); SPMD programs must run identical "
                               "Session.run sequences" % (where, self.rank, head.name))
        return [v[2] for v in votes]

    def _exchange_table(self, gid: int, head: Task, table, nparts: int, bad: int = 0) -> None:
        from ..frame.deferred import DeferredFrame
        from ..ops import gpu
        from ..ops import keypack
        typ = head.type.typ
        lay = keypack.layout_for(typ)
        vdtype, op = typ.out(typ.prefix()).torch_dtype, head.combiner.builtin
        # result buffers hold <packed key word, value>; anything but a plain int64
        # key is unpacked into its key columns when the frame is first looked at
        fin = None if lay.plain_int64 else (
            lambda d, n, lay=lay: keypack.unpack(d[0][:n], lay) + [d[1][:n]])
        t0 = time.perf_counter()
        ev0 = self._cuda_event()
        if self.shuffle == "nccl-table" and not (self.world == 1 and nparts == 1):
            if table is None:
                table = gpu.table_pool().get(1024, vdtype, op, self.device)
            return self._exchange_table_library(gid, head, table, nparts, bad, t0, ev0)
        rx = self._rx()
        rx.harvest()
        lkey = self._hint_key(head)
        out: Dict[int, List[Frame]] = {}
        if self.world == 1 and nparts == 1:
            # the local table is already the answer: asynchronous compaction, the row
            # count goes to pinned memory (no host synchronisation)
            if bad:
                raise errors.fatal("spmd: a producer of %s failed" % head.name)
            if table is None:
                out[0] = []
            else:
                hint = self.card_hints.get(lkey)
                cap = min(table.total + 1, int(hint * 1.25) + 4096) if hint is not None else table.total + 1
                ok = torch.empty(cap, dtype=torch.int64, device=self.device)
                ov = torch.empty(cap, dtype=vdtype, device=self.device)
                table.compact_into(ok, ov, clear=True)

                def learned(rep, key=lkey):
                    if rep.error == 0:
                        self.card_hints[key] = max(int(rep.out_rows[0]), 1 << 10)
                    else:
                        self.card_hints[key] = 4 * max(self.card_hints.get(key, 0), int(rep.out_rows[0]), 1 << 14)
                pend = rx.publish(table.status, cap, on_report=learned)
                with self._glock:
                    self._group_pend[gid] = pend
                gpu.table_pool().put(table, clean=True)
                out[0] = [DeferredFrame([ok, ov], typ, self._resolver(pend, gid, head, 0), finalize=fin)]
            mode = "table-local"
        else:
            if self._rx_reset or self._want_arena:
                # a lost exchange left epochs or the arena in doubt: start over (collective)
                with span("xchg/reset"):
                    try:
                        self._peers_alive()
                        rx.reset(max(self._want_arena, rx.arena_bytes))
                    except errors.Error:
                        raise
                    except BaseException as e:  # noqa: BLE001
                        if _is_comm_failure(e):
                            raise errors.TaskLostError("spmd: cannot re-establish the exchange: %s" % e)
                        raise
                self._rx_reset, self._want_arena = False, 0
            if self.world == 1 and rx.arena_bytes < self.arena_bytes:
                rx.reset(self.arena_bytes)  # single rank: the arena is allocated on first use
            if table is None:  # this rank had no rows: an empty table still takes part
                table = gpu.table_pool().get(1024, vdtype, op, self.device)
            acc, fps = self._take_fingerprint(head.name.inv_index)
            acc = self._tag_exchange(acc, head)
            # order the exchange stream after the producers' combine kernels
            after = torch.cuda.Event()
            after.record(torch.cuda.default_stream(self.device))
            rx.stream.wait_event(after)
            pl, _ = rx.geometry(nparts)
            locals_ = [j for j in range(pl) if self.rank + j * self.world < nparts]
            dkeys = [self._hint_key(head, "dest%d" % (self.rank + j * self.world)) for j in locals_]
            hints = [self.card_hints.get(k) for k in dkeys]
            pend = None
            first_err = None
            first_run = not all(h is not None for h in hints)
            # destination tables that do not fit in L2 are combined region by region
            # (rows are bucketed by region at the source); the region count decides the
            # landing layout, so it must be the same on every rank whichever path each
            # takes: it only ever changes at a board vote (below), from the table sizes
            # all ranks posted there
            regions = int(self.card_hints.get(self._hint_key(head, "regions"), 1))
            with torch.cuda.stream(rx.stream):
                if all(h is not None for h in hints):
                    dests = []
                    for j, h, dk in zip(locals_, hints, dkeys):
                        # The destination table and the output buffers are tied to the
                        # slice (not pooled): the same pointers every step, so the
                        # captured graph is replayed instead of re-recorded.
                        dcap = gpu.quantise_capacity(
                            (gpu.table_capacity(int(h * 1.15) + 1024) + regions - 1) // regions) * regions
                        dt = self._dest_tables.get((dk, dcap, vdtype, op))
                        if dt is None:
                            for k in [k for k in self._dest_tables if k[0] == dk]:
                                del self._dest_tables[k]
                            dt = self._dest_tables[(dk, dcap, vdtype, op)] = gpu.HashTable(
                                dcap, vdtype, op, self.device, warm=0)
                        cap = gpu.quantise_capacity(int(h * 1.25) + 4096)
                        ok, ov = self._out_buffers(dk, cap, vdtype)
                        dests.append((j, dt, ok, ov, cap))
                    with span("xchg/enqueue"):
                        pend = rx.exchange(table, nparts, dests, vote=(1 if bad else 0, acc),
                                           on_report=self._learner(head, lkey, dkeys, nparts),
                                           keymode=lay.mode, regions=regions)
                else:
                    # first run of this slice: one synchronisation to learn how much
                    # arrives, then everything is sized exactly
                    with span("xchg/first_run_sync"):
                        probe = rx.exchange_probe(table, nparts, vote=(1 if bad else 0, acc),
                                                  keymode=lay.mode, regions=regions)
                        rep = probe.resolve()
                    dests = []
                    if rep.error == 0:
                        for j in locals_:
                            rows = int(rep.recv_rows[j])
                            # (quantised capacities are multiples of 64: they split into regions)
                            dt = gpu.table_pool().get(gpu.table_capacity(rows + 1024), vdtype, op,
                                                      self.device, warm=0)
                            cap = rows + 16
                            dests.append((j, dt, torch.empty(cap, dtype=torch.int64, device=self.device),
                                          torch.empty(cap, dtype=vdtype, device=self.device), cap))
                    else:
                        from ..parallel.rx import ExchangeFailed
                        first_err = ExchangeFailed(rep.error, rep.need_rows)
                        self._learner(head, lkey, dkeys, nparts)(rep)
                    pend = rx.exchange_finish(table, nparts, dests,
                                              on_report=None if first_err else
                                              self._learner(head, lkey, dkeys, nparts), regions=regions)
                done = pend.event
            with self._glock:
                self._group_pend[gid] = pend
            table.ready = done  # the next user's stream waits for the scatter on the device
            gpu.table_pool().put(table, clean=True)
            if first_run:
                for j, dt, ok, ov, cap in dests:
                    gpu.table_pool().put(dt, clean=True)
            if self.fault_hook is not None:
                self.fault_hook("after_enqueue", str(head.name))
            # host-side agreement (shared memory): a fatal producer failure on any
            # rank fails the invocation everywhere, now, without draining the GPU
            # the vote also carries the size of my largest destination table: every rank
            # derives the same region count for the *next* run of this slice from it
            mine = max([gpu.table_capacity(int(self.card_hints[k] * 1.15) + 1024) * 16
                        for k in dkeys if self.card_hints.get(k) is not None] or [0])
            sizes = self._board_vote(bad, acc, fps, head, extra=mine)
            if all(x > 0 for x in sizes):
                self.card_hints[self._hint_key(head, "regions")] = rx.regions_for(max(sizes), nparts)
            if first_err is not None:
                if first_err.retriable:
                    raise errors.TaskLostError("spmd: exchange of %s was lost (%s)" % (head.name, first_err))
                raise errors.fatal("spmd: exchange of %s failed: %s" % (head.name, first_err))
            import weakref
            for (j, dt, ok, ov, cap), dk in zip(dests, dkeys):
                fr = DeferredFrame([ok, ov], typ, self._resolver(pend, gid, head, j), ready_event=done,
                                   finalize=fin)
                if not first_run:
                    for ent in self._out_sets.get(dk, ()):
                        if ent[0] is ok:
                            ent[2] = weakref.ref(fr)
                out[self.rank + j * self.world] = [fr]
            mode = "fused-table"
        with self._glock:
            for p in range(self.rank, nparts, self.world):
                self.exchanged[(gid, p)] = list(out.get(p, []))
        self._log_exchange(head, mode, t0, ev0, sync=False)

    def _out_buffers(self, dk: tuple, cap: int, vdtype):
        """Output buffers of a destination shard: the first set whose previous
        result frame is gone (discarded / garbage collected) is reused, so a
        run-discard loop sees the same pointers every step.  Reuse is ordered on the
        device: the exchange stream waits for everything queued on the default
        stream (the consumers of the earlier result) before it writes."""
        sets = self._out_sets.setdefault(dk, [])
        for ent in sets:
            if ent[0].numel() == cap and ent[1].dtype == vdtype and (ent[2] is None or ent[2]() is None):
                ent[2] = None
                return ent[0], ent[1]
        if len(sets) >= 8:  # results are being kept: stop tracking old sets
            sets[:] = [e for e in sets if e[2] is not None and e[2]() is not None][-4:]
        ok = torch.empty(cap, dtype=torch.int64, device=self.device)
        ov = torch.empty(cap, dtype=vdtype, device=self.device)
        sets.append([ok, ov, None])
        return ok, ov

    def _learner(self, head: Task, lkey, dkeys, nparts: int):
        """Report callback: cardinalities for the next run's table/buffer sizes;
        what a failed exchange needs to succeed next time."""
        from ..parallel import rx as _rx

        def learned(rep):
            if rep.error == 0:
                self.card_hints[lkey] = max(int(rep.sent_rows), 1 << 10)
                for i, k in enumerate(dkeys):
                    self.card_hints[k] = max(int(rep.out_rows[i]), 1 << 10)
                return
            if rep.error == _rx.LANDING_OVERFLOW and self.rx is not None:
                self._want_arena = max(self._want_arena,
                                       self.rx.bytes_for(nparts, int(rep.need_rows * 1.5) + 4096))
            elif rep.error == _rx.TABLE_OVERFLOW:
                self.card_hints[lkey] = 4 * max(self.card_hints.get(lkey, 0), 1 << 16)
            elif rep.error == _rx.DEST_OVERFLOW:
                for k in dkeys:  # next attempt takes the exactly-sized first-run path
                    self.card_hints.pop(k, None)
            elif rep.error == _rx.TIMEOUT:
                self._rx_reset = True
        return learned

    def _resolver(self, pend, gid: int, head: Task, j: int):
        """Length of a deferred result frame; a failed exchange is *lost* (its
        producers are re-run by the evaluator) unless the verdict is fatal."""
        from ..parallel.rx import ExchangeFailed

        def resolve() -> int:
            try:
                return pend.rows(j)
            except ExchangeFailed as e:
                if e.retriable:
                    self._mark_group_lost(gid, head, pend)
                    raise errors.TaskLostError("spmd: exchange of %s was lost (%s); recomputing"
                                               % (head.name, e)) from e
                raise errors.fatal("spmd: exchange of %s failed: %s" % (head.name, e)) from e
        return resolve

    def _mark_group_lost(self, gid: int, head: Task, pend=None) -> None:
        """The group's exchange did not deliver: forget it and mark the producers
        LOST so that the evaluator re-runs them (and, through ``_ensure_group``,
        the exchange).  ``pend`` names the attempt that failed: a consumer that
        still holds a frame of an *earlier* attempt finds a newer exchange in place
        and only re-runs itself — one failed attempt is repeated exactly once, on
        every rank (the ranks must issue the same number of exchanges)."""
        with self._glock:
            if gid not in self.group_done:
                return  # already handled by another consumer
            if pend is not None and self._group_pend.get(gid) is not pend:
                return  # superseded: the repeated exchange is already queued or done
            self._group_pend.pop(gid, None)
            self.group_done.pop(gid, None)
            self.group_err.pop(gid, None)
            for k in [k for k in self.exchanged if k[0] == gid]:
                del self.exchanged[k]
        for t in head.group or [head]:
            with self.lock:
                self.buffers.pop(id(t), None)
            if t.state() == TaskState.OK:
                t.set(TaskState.LOST)

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
        # (a row with the reserved EMPTY key is part of the compacted output and
        # travels like any other row)
        self.card_hints[self._hint_key(head)] = max(int(k.numel()), 1 << 10)
        with self._glock:
            for p in range(self.rank, nparts, self.world):
                self.exchanged[(gid, p)] = [x for x in out.get(p, []) if len(x)]
        self._log_exchange(head, "nccl-table", t0, ev0)

    def _cuda_event(self):
        if self.device is not None and self.device.type == "cuda":
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            return ev
        return None

    def _log_exchange(self, head: Task, mode: str, t0: float, ev0, sync: bool = True, **kw) -> None:
        """Record the exchange in the log and, device-timed with CUDA events,
        in the session trace (pid = rank, category "exchange").  ``sync=False``
        (asynchronous exchanges): host time only, nothing waits for the device."""
        rec = {"group": str(head.name), "mode": mode, "seconds": time.perf_counter() - t0}
        rec.update(kw)
        if len(self.exchange_log) > 4096:
            del self.exchange_log[:2048]
        if ev0 is not None and sync:
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

    def _replicate(self, f: Frame):
        """Every rank ends up with the rows of all ranks (a replicated
        dependency, e.g. the small side of a BroadcastJoin): NVSwitch multicast
        when the columns are device tensors, the library otherwise."""
        if self.world == 1:
            return "replicate-local", f
        tensor_only = all(isinstance(c, torch.Tensor) for c in f.cols()) and f.num_out() > 0
        if tensor_only and f.device.type == "cuda":
            from ..parallel.broadcast import Replicator
            if getattr(self, "_replicator", None) is None:
                self._replicator = Replicator(dist.group.WORLD, self.arena_bytes, self.device, self.ctrl_group)
            cols = self._replicator.all_gather(f.cols())
            return "replicate-" + self._replicator.mode, Frame(cols, f.typ)
        if tensor_only:
            from ..parallel.broadcast import Replicator
            rep = Replicator(self.ctrl_group, 0, "cpu", self.ctrl_group)
            return "replicate-library", Frame(rep.all_gather(f.cols()), f.typ)
        # any column type: codec-encoded frames over the control plane
        blobs = [None] * self.world
        dist.all_gather_object(blobs, encode_frame(f.to("cpu") if f.device.type != "cpu" else f) if len(f) else b"",
                               group=self.ctrl_group)
        parts = [_decode_one(b, f.typ, self.device) for b in blobs if b]
        return "replicate-object", (concat(parts, f.typ, self.device) if parts else f)

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
        """Any column type (host strings, lists, python objects): codec-encoded
        per-shard frames, each sent **only to the rank that owns the shard** — one
        all-to-all of byte counts and one of the bytes over the control-plane group
        (round 1 all-gathered every rank's whole send set to every rank)."""
        parts = split_by_partition(f, ids, nparts) if len(f) else []
        per_rank: List[Dict[int, bytes]] = [dict() for _ in range(self.world)]
        for p, g in enumerate(parts):
            if len(g) and self.owner(p) != self.rank:
                per_rank[self.owner(p)][p] = encode_frame(g.to("cpu") if g.device.type != "cpu" else g)
        received: List[Dict[int, bytes]] = []
        if self.world > 1:
            blobs = [pickle.dumps(d, protocol=pickle.HIGHEST_PROTOCOL) if d else b"" for d in per_rank]
            send_counts = torch.tensor([len(b) for b in blobs], dtype=torch.int64)
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts, group=self.ctrl_group)
            payload = b"".join(blobs)
            send = torch.frombuffer(bytearray(payload), dtype=torch.uint8) if payload \
                else torch.empty(0, dtype=torch.uint8)
            recv = torch.empty(int(recv_counts.sum()), dtype=torch.uint8)
            dist.all_to_all_single(recv, send, recv_counts.tolist(), send_counts.tolist(), group=self.ctrl_group)
            raw = recv.numpy().tobytes()
            at = 0
            for n in recv_counts.tolist():
                received.append(pickle.loads(raw[at:at + n]) if n else {})
                at += n
        out: Dict[int, List[Frame]] = {}
        for p in range(self.rank, nparts, self.world):
            frs = []
            if parts and len(parts[p]):
                frs.append(parts[p])
            for r, sends in enumerate(received):
                if r == self.rank or p not in sends:
                    continue
                frs.append(_decode_one(sends[p], f.typ, self.device))
            out[p] = frs
        return out

    # -- task execution ---------------------------------------------------------
    def run(self, task: Task) -> None:
        if task.group is not None:
            self._ensure_group(task)
        if not self.owns(task):
            task.set(TaskState.RUNNING)
            task.set(TaskState.OK)  # executed by its owner rank
            return
        if self.fault_hook is not None:
            self.fault_hook("run", str(task.name))
        # LOST stays LOST: the evaluator re-runs the task (a transient producer
        # failure is retried locally, before the group's exchange happens)
        super().run(task)

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
            for dt in dep.head.group:
                if dt.state() != TaskState.OK:
                    raise errors.TaskLostError("dependency %s is not available" % dt.name)
            ev = self.group_done.get(gid)
            if ev is None:
                raise errors.TaskLostError("spmd: the exchange of %s is not available" % dep.head.name)
            ev.wait()
            err = self.group_err.get(gid)
            if err is not None:
                if isinstance(err, errors.TaskLostError) or not errors.is_fatal(err):
                    self._mark_group_lost(gid, dep.head)
                    raise errors.TaskLostError("spmd: exchange for %s was lost: %s" % (dep.head.name, err))
                raise errors.fatal("spmd: exchange for %s failed" % dep.head.name, err)
            with self._glock:
                frames = self.exchanged.get((gid, dep.partition))
            if frames is None:
                raise errors.TaskLostError("spmd: the exchange of %s is not available" % dep.head.name)
            typ = dep.head.type.typ
            r: Reader
            if frames and all(isinstance(f, _deferred()) for f in frames):
                r = _DeferredReader(frames)  # hands the frames over without touching their length
            elif frames:
                r = _TaggedReader(concat(frames, typ, self.device))
            else:
                r = EmptyReader()
            if dep.head.combiner is not None:
                r.combined_unique = True
            ins.append(r)
        return ins

    def buffer_output(self, task: Task, out: Reader, ctx) -> TaskBuffer:
        if task.group is None:
            if task.num_partition <= 1 and task.combiner is None:
                return self._buffer_passthrough(task, out)
            return super().buffer_output(task, out, ctx)
        # Shuffle producer: keep whole tiles; the group exchange partitions (and
        # combines) them in one fused pass.
        buf = TaskBuffer(1)
        nrec = 0
        table = None
        fast = self._table_path_ok(task)
        if fast:
            from ..ops import keypack
            lay = keypack.layout_for(task.type.typ)
            npre = task.type.typ.prefix()
        try:
            while True:
                with span("task/read"):
                    f = out.read(None)
                if f is None:
                    break
                nrec += len(f)
                if fast and len(f):
                    if table is None:
                        with span("task/group_table"):
                            table = self._group_table(task, len(f))
                    with span("task/combine_launch"):
                        # async: overlaps the next read
                        table.combine(f.col(0) if lay.plain_int64 else
                                      keypack.pack([f.col(i) for i in range(npre)], lay), f.col(npre))
                else:
                    buf.append(0, f)
        finally:
            out.close()
        task.stats = {"inrecords": nrec, "write": nrec}
        return buf

    def _buffer_passthrough(self, task: Task, out: Reader) -> TaskBuffer:
        """Unpartitioned output: frames are kept as they come.  Deferred frames
        (results of an asynchronous exchange) are stored *without* asking for their
        length, so a Reduce step runs end to end with no host synchronisation."""
        DF = _deferred()
        buf = TaskBuffer(1)
        nrec, lazy = 0, False
        try:
            while True:
                f = out.read(None)
                if f is None:
                    break
                if isinstance(f, DF) and not f.resolved():
                    buf.parts.add(0, f)
                    lazy = True
                    continue
                if len(f):
                    nrec += len(f)
                    buf.append(0, f)
        finally:
            out.close()
        task.stats = {"write": buf.nrows} if lazy else {"inrecords": nrec, "write": buf.nrows}
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
                self._group_pend.pop(gid, None)

    # -- results -------------------------------------------------------------------
    def result_reader(self, result) -> Reader:
        """Rows of the shards this rank owns, in shard order (SPMD view).  A
        result whose output (or whose exchange) was lost since the run is
        re-evaluated when it is read, as in the reference's fault-tolerant scans
        (``exec/bigmachine.go:1500-1513``); every rank must then read it too."""
        return MultiReader([_ReevalReader(self, t) for t in result.tasks if self.owns(t)])

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


def _is_comm_failure(e: BaseException) -> bool:
    """torch.distributed / transport errors that mean "a peer is unreachable"."""
    for name in ("DistNetworkError", "DistBackendError", "DistStoreError"):
        cls = getattr(dist, name, None)
        if cls is not None and isinstance(e, cls):
            return True
    if isinstance(e, (ConnectionError, TimeoutError, BrokenPipeError)):
        return True
    msg = str(e)
    return isinstance(e, RuntimeError) and any(
        s in msg for s in ("Connection closed by peer", "Connection reset", "Broken pipe", "timed out",
                           "Timed out", "Socket Timeout", "NCCL communicator was aborted", "ncclRemoteError"))


def _deferred():
    from ..frame.deferred import DeferredFrame
    return DeferredFrame


class _DeferredReader(Reader):
    """Hands deferred frames to the consumer as they are; the consumer's stream
    waits for the producing stream on the device."""
    combined_unique = False

    def __init__(self, frames: List[Frame]):
        self.frames = list(frames)
        self.cur: Optional[FrameReader] = None

    def read(self, n=None):
        while True:
            if self.cur is not None:
                f = self.cur.read(n)
                if f is not None:
                    return f
                self.cur = None
            if not self.frames:
                return None
            f = self.frames.pop(0)
            f.wait_ready()
            if n is None:
                return f
            self.cur = FrameReader(f)


class _ReevalReader(Reader):
    """Reads a task's output; if it (or the exchange behind it) turns out to be
    lost, the task is re-evaluated and the read starts over."""

    def __init__(self, ex: "SpmdExecutor", task: Task):
        self.ex, self.task = ex, task
        self.frames: Optional[List[Frame]] = None
        self.tries = 0

    def _load(self) -> List[Frame]:
        from .eval import evaluate
        while True:
            try:
                if self.task.state() != TaskState.OK:
                    evaluate(self.ex, [self.task])
                r = self.ex.reader(self.task, 0)
                out = []
                for f in r:
                    len(f)  # resolve deferred frames here, where a loss can be handled
                    out.append(f)
                r.close()
                return out
            except errors.TaskLostError:
                self.tries += 1
                if self.tries > 5:
                    raise
                if self.task.state() == TaskState.OK:
                    self.task.set(TaskState.LOST)

    def read(self, n=None):
        if self.frames is None:
            self.frames = self._load()
            self.cur = None
        while True:
            if self.cur is not None:
                f = self.cur.read(n)
                if f is not None:
                    return f
                self.cur = None
            if not self.frames:
                return None
            self.cur = FrameReader(self.frames.pop(0))


def _decode_one(b: bytes, typ, device) -> Frame:
    import io
    from ..sliceio import Decoder, read_all
    return read_all(Decoder(io.BytesIO(b), typ, device), typ, device)
