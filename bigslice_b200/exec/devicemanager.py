"""Multi-device resource manager and executor (single process, many GPUs).

``DeviceManager`` is the scheduler's view of the box: every device has a proc
capacity, a load, a health state and the set of tasks whose output lives on
it.  Tasks are *offered* the least-loaded healthy device that can fit them
(lower invocation index first among waiters; big requests are not starved:
first-fit over a priority-ordered wait list).  A device-level failure (CUDA
error, OOM) puts the device on *probation* — no new work for
``ProbationTimeout`` seconds or until it completes a task successfully — and a
lost device marks every task resident on it ``LOST`` so the evaluator
recomputes them elsewhere.

``MultiDeviceExecutor`` runs tasks through it; task outputs stay on the device
that computed them and are copied to the consumer's device on read (NVLink
P2P copy).  This is the single-process analogue of the reference's cluster
resource manager (``exec/slicemachine.go:26-817``: sliceMachine, machineManager,
Offer/Done, probation, loss) with GPUs as machines; the SPMD executor is the
high-throughput path, this one keeps the reference's elastic/fault model.
"""
from __future__ import annotations

import heapq
import itertools
import threading
import time
from typing import Callable, Dict, List, Optional, Sequence, Set

import torch

from .. import context as _context
from .. import errors
from ..sliceio import FuncReader, Reader
from .local import LocalExecutor
from .task import Task, TaskState

#: Seconds a device stays out of scheduling after a device-level error
#: (reference ``exec.ProbationTimeout`` = 30 s).
ProbationTimeout = 30.0

OK, PROBATION, LOST = "ok", "probation", "lost"


class Device:
    def __init__(self, index: int, device, maxprocs: int):
        self.index, self.device, self.maxprocs = index, device, maxprocs
        self.curprocs = 0
        self.health = OK
        self.probation_until = 0.0
        self.tasks: Set[int] = set()
        self.task_objs: Dict[int, Task] = {}
        self.done_count = 0
        self.err_count = 0

    def load(self) -> float:
        return self.curprocs / float(self.maxprocs)

    def __repr__(self):
        return "dev%d(%s %d/%d %s)" % (self.index, self.device, self.curprocs, self.maxprocs, self.health)


class DeviceManager:
    def __init__(self, devices: Sequence, procs_per_device: int = 1, max_load: float = 0.95):
        self.devs = [Device(i, d, max(1, int(round(procs_per_device * max_load + 0.4999))) if max_load < 1
                            else procs_per_device) for i, d in enumerate(devices)]
        for d in self.devs:
            d.maxprocs = max(1, min(procs_per_device, d.maxprocs))
        self.cond = threading.Condition()
        self.waiters: List = []  # heap of (priority, seq, procs, slot)
        self.seq = itertools.count()

    # -- scheduling ----------------------------------------------------------
    def _eligible(self, procs: int) -> Optional[Device]:
        now = time.time()
        best = None
        for d in self.devs:
            if d.health == LOST:
                continue
            if d.health == PROBATION:
                if now < d.probation_until:
                    continue
                d.health = OK  # probation expired
            need = min(procs, d.maxprocs)
            if d.curprocs + need > d.maxprocs:
                continue
            # least loaded first; ties go to the device that has done the least
            if best is None or (d.load(), d.done_count + d.err_count) < \
                    (best.load(), best.done_count + best.err_count):
                best = d
        return best

    def offer(self, priority: int, procs: int, timeout: Optional[float] = None) -> Device:
        """Block until a device can take ``procs``; returns it with the procs
        already reserved."""
        slot = {"dev": None}
        with self.cond:
            heapq.heappush(self.waiters, (priority, next(self.seq), procs, slot))
            deadline = None if timeout is None else time.time() + timeout
            while True:
                self._dispatch()
                if slot["dev"] is not None:
                    return slot["dev"]
                if all(d.health == LOST for d in self.devs):
                    self._remove(slot)
                    raise errors.fatal("no devices left")
                wait = 0.25
                if deadline is not None:
                    wait = min(wait, deadline - time.time())
                    if wait <= 0:
                        self._remove(slot)
                        raise errors.Error("offer timed out", errors.Kind.Timeout)
                self.cond.wait(wait)

    def _remove(self, slot) -> None:
        self.waiters = [w for w in self.waiters if w[3] is not slot]
        heapq.heapify(self.waiters)

    def _dispatch(self) -> None:
        """Serve waiters in priority order; a request that does not fit is
        shelved so smaller ones behind it can still be placed (first-fit)."""
        shelved = []
        while self.waiters:
            pr, sq, procs, slot = heapq.heappop(self.waiters)
            d = self._eligible(procs)
            if d is None:
                shelved.append((pr, sq, procs, slot))
                continue
            d.curprocs += min(procs, d.maxprocs)
            slot["dev"] = d
        for w in shelved:
            heapq.heappush(self.waiters, w)
        if any(w[3]["dev"] is None for w in shelved) or True:
            self.cond.notify_all()

    def done(self, d: Device, procs: int, err: Optional[BaseException] = None) -> None:
        with self.cond:
            d.curprocs -= min(procs, d.maxprocs)
            if err is None:
                d.done_count += 1
                if d.health == PROBATION:
                    d.health = OK  # a success ends probation early
            else:
                d.err_count += 1
                if d.health != LOST:
                    d.health = PROBATION
                    d.probation_until = time.time() + ProbationTimeout
            self.cond.notify_all()

    def assign(self, d: Device, task: Task) -> None:
        with self.cond:
            if d.health == LOST:
                task.set(TaskState.LOST)
                return
            d.tasks.add(id(task))
            d.task_objs[id(task)] = task

    def unassign(self, task: Task) -> None:
        with self.cond:
            for d in self.devs:
                d.tasks.discard(id(task))
                d.task_objs.pop(id(task), None)

    def lose(self, index: int) -> List[Task]:
        """The device is gone: everything resident on it is LOST."""
        with self.cond:
            d = self.devs[index]
            d.health = LOST
            tasks = list(d.task_objs.values())
            d.tasks.clear()
            d.task_objs.clear()
            self.cond.notify_all()
        for t in tasks:
            if t.state() == TaskState.OK:
                t.set(TaskState.LOST)
        return tasks

    def revive(self, index: int) -> None:
        with self.cond:
            self.devs[index].health = OK
            self.cond.notify_all()

    def poll(self) -> List[dict]:
        """Per-device health snapshot: procs, load, tasks resident, and for CUDA
        devices free/total HBM and bytes held by the caching allocator (the
        analogue of the reference's periodic mem/disk/load polling of machines,
        ``exec/slicemachine.go:238-257``)."""
        out = []
        with self.cond:
            devs = list(self.devs)
        for d in devs:
            rec = {"index": d.index, "device": str(d.device), "health": d.health, "procs": d.curprocs,
                   "maxprocs": d.maxprocs, "load": round(d.load(), 3), "tasks": len(d.tasks),
                   "done": d.done_count, "errors": d.err_count}
            if d.device.type == "cuda" and d.health != LOST:
                try:
                    free, total = torch.cuda.mem_get_info(d.device)
                    rec.update(hbm_free=free, hbm_total=total,
                               hbm_allocated=torch.cuda.memory_allocated(d.device))
                except Exception as e:  # noqa: BLE001 - a dying device must not kill the poller
                    rec["poll_error"] = str(e)[:80]
            out.append(rec)
        return out

    def status(self) -> List[str]:
        """One line per device (rendered under ``machines`` in /debug/status)."""
        lines = []
        for r in self.poll():
            mem = ""
            if "hbm_total" in r:
                mem = " hbm %.1f/%.1f GiB" % ((r["hbm_total"] - r["hbm_free"]) / 2**30, r["hbm_total"] / 2**30)
            lines.append("dev%d(%s %d/%d %s)%s load %.2f tasks %d done %d err %d" % (
                r["index"], r["device"], r["procs"], r["maxprocs"], r["health"], mem, r["load"],
                r["tasks"], r["done"], r["errors"]))
        return lines


class MultiDeviceExecutor(LocalExecutor):
    def __init__(self, devices: Optional[Sequence] = None, procs_per_device: int = 1,
                 max_load: float = 1.0, fail_hook: Optional[Callable] = None):
        if devices is None:
            n = torch.cuda.device_count() if torch.cuda.is_available() else 0
            devices = [torch.device("cuda", i) for i in range(n)] or [torch.device("cpu")]
        devices = [torch.device(d) if not isinstance(d, torch.device) else d for d in devices]
        super().__init__(len(devices) * procs_per_device, None)
        self.mgr = DeviceManager(devices, procs_per_device, max_load)
        self.fail_hook = fail_hook
        self.where: Dict[int, Device] = {}

    def name(self) -> str:
        return "multidevice"

    def max_workers(self) -> int:
        return max(8, self.parallelism * 2 + 8)

    def run(self, task: Task) -> None:
        procs = task.pragma.procs()
        inv = task.name.inv_index
        if task.pragma.exclusive():
            procs = max(d.maxprocs for d in self.mgr.devs)
        dev = self.mgr.offer(inv, procs)
        err: Optional[BaseException] = None
        try:
            if dev.device.type == "cuda":
                torch.cuda.set_device(dev.device)
            if self.fail_hook is not None:
                self.fail_hook(dev, task)
            self.last_device_error = None
            self._run(task, dev.device if dev.device.type != "cpu" else None)
            if task.state() == TaskState.LOST and self.last_device_error is not None:
                err = self.last_device_error
        except BaseException as e:  # injected / device-level failure
            err = e
            if task.state() < TaskState.OK:
                task.set(TaskState.LOST)
        finally:
            self.mgr.done(dev, procs, err)
        if task.state() == TaskState.OK:
            with self.lock:
                self.where[id(task)] = dev
            self.mgr.assign(dev, task)

    def reader(self, task: Task, partition: int) -> Reader:
        r = super().reader(task, partition)
        target = _context.current_device()

        def read(n):
            f = r.read(n)
            if f is None or target is None or f.device == target:
                return f
            return f.to(target)  # P2P copy over NVLink between resident GPUs
        return FuncReader(read, r.close)

    def discard(self, task: Task) -> None:
        self.mgr.unassign(task)
        with self.lock:
            self.where.pop(id(task), None)
        super().discard(task)

    def lose_device(self, index: int) -> int:
        """Simulate / handle the loss of a device."""
        tasks = self.mgr.lose(index)
        with self.lock:
            for t in tasks:
                self.buffers.pop(id(t), None)
                self.where.pop(id(t), None)
        return len(tasks)


def MultiDevice(devices=None, procs_per_device: int = 1, **kw):
    """Session option: one process driving several devices with the resource
    manager (elastic, loss-tolerant)."""
    from .session import _set_executor
    return _set_executor(lambda s: MultiDeviceExecutor(devices, procs_per_device,
                                                       s.max_load_ or 1.0, **kw))
