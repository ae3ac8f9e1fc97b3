"""Fault-injecting executor for tests: the "killable machine" harness.

``ChaosExecutor`` is a :class:`LocalExecutor` that randomly fails task runs
(the task ends ``LOST`` before producing output), randomly drops the stored
output of tasks that already completed (spontaneous post-success loss), and can
``kill()`` everything it holds at once — the analogue of the reference's
``testsystem.Kill`` / ``rpc.InjectFailures`` / ``stressExecutor`` seams
(SURVEY §4.2).  The evaluator must still drive every computation to the right
answer by recomputing lost lineage.

Reference seams: ``exec/evalstress_test.go:22-185`` (stressExecutor: random
delays, LOST with rate <= 20 %, spontaneous post-success loss),
``exec/chaosmonkey_test.go:22-235`` (machine kills during a Reduce),
``exec/bigmachine_test.go:625-676`` (testsystem helpers).
"""
from __future__ import annotations

import random
import threading
from typing import Optional

from .local import LocalExecutor
from .task import Task, TaskState


class ChaosExecutor(LocalExecutor):
    def __init__(self, parallelism: int = 4, device=None, fail_rate: float = 0.1,
                 loss_rate: float = 0.05, seed: int = 0):
        super().__init__(parallelism, device)
        self.fail_rate = fail_rate
        self.loss_rate = loss_rate
        self.rng = random.Random(seed)
        self.rng_lock = threading.Lock()
        self.injected_failures = 0
        self.injected_losses = 0
        self.kills = 0

    def name(self) -> str:
        return "chaos"

    def _roll(self) -> float:
        with self.rng_lock:
            return self.rng.random()

    def run(self, task: Task) -> None:
        if self._roll() < self.fail_rate:
            # the "machine" died before the task produced anything
            self.injected_failures += 1
            task.set(TaskState.RUNNING)
            task.set(TaskState.LOST)
            return
        super().run(task)
        if task.state() == TaskState.OK and self._roll() < self.loss_rate:
            victim = self._random_completed()
            if victim is not None:
                self.injected_losses += 1
                super().discard(victim)

    def _random_completed(self) -> Optional[Task]:
        with self.lock:
            ids = list(self.buffers)
        if not ids or self.session is None:
            return None
        with self.rng_lock:
            tid = self.rng.choice(ids)
        with self.session._lock:
            roots = list(self.session.roots.values())
        found = []

        from .task import iter_tasks
        iter_tasks(roots, lambda t: found.append(t) if id(t) == tid else None)
        return found[0] if found else None

    def kill(self) -> int:
        """Lose every stored output at once (a machine going away)."""
        self.kills += 1
        with self.lock:
            ids = set(self.buffers)
            self.buffers.clear()
        n = 0
        if self.session is not None:
            with self.session._lock:
                roots = list(self.session.roots.values())
            from .task import iter_tasks

            def mark(t: Task):
                nonlocal n
                if id(t) in ids and t.state() == TaskState.OK:
                    t.set(TaskState.LOST)
                    n += 1
            iter_tasks(roots, mark)
        return n


def Chaos(fail_rate: float = 0.1, loss_rate: float = 0.05, seed: int = 0, device=None):
    """Session option selecting the fault-injecting executor."""
    from .session import _set_executor
    return _set_executor(lambda s: ChaosExecutor(s.p, device, fail_rate, loss_rate, seed))
