"""Evaluator tests with fake executors and scripted state transitions (model:
reference exec/eval_test.go:25-552 and exec/evalstress_test.go:22-185)."""
import random
import threading
import time

import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200 import errors
from bigslice_b200.exec import eval as ev
from bigslice_b200.exec.compile import ExecInvocation, compile_slice
from bigslice_b200.exec.eval import Executor, evaluate
from bigslice_b200.exec.task import Task, TaskDep, TaskName, TaskState, iter_tasks
from bigslice_b200.func import Invocation

OK, ERR, LOST, RUNNING, INIT = (TaskState.OK, TaskState.ERR, TaskState.LOST,
                                TaskState.RUNNING, TaskState.INIT)


class ManualExecutor(Executor):
    """Only flips WAITING->RUNNING; the test body drives the rest."""

    def __init__(self):
        self.ran = []
        self.lock = threading.Lock()

    def run(self, task):
        with self.lock:
            self.ran.append(task)
        task.set(RUNNING)


def simple_graph(nshard=2, stages=2):
    def build():
        s = bs.Const(nshard, torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64))
        for _ in range(stages - 1):
            s = bs.Reduce(s, bs.ops.add)
        return s
    return compile_slice(ExecInvocation(Invocation(1, 0, (), False, "t")), build())


def all_tasks(roots):
    out = []
    iter_tasks(roots, out.append)
    return out


def start_eval(ex, roots):
    res = {}

    def go():
        try:
            evaluate(ex, roots)
            res["err"] = None
        except BaseException as e:
            res["err"] = e
    th = threading.Thread(target=go, daemon=True)
    th.start()
    return th, res


def wait_state(task, state, timeout=5):
    assert task.wait_state(state, timeout) >= state, "%s stuck in %s" % (task.name, task.state())


def test_task_err_aborts_eval():
    roots = simple_graph(1, 2)
    ex = ManualExecutor()
    th, res = start_eval(ex, roots)
    dep = roots[0].deps[0].head
    wait_state(dep, RUNNING)
    dep.error(errors.fatal("some error"))
    th.join(5)
    assert not th.is_alive() and res["err"] is not None and "some error" in str(res["err"])
    assert roots[0].state() == INIT


def test_all_roots_evaluated_and_deps_first():
    roots = simple_graph(3, 2)
    ex = ManualExecutor()
    th, res = start_eval(ex, roots)
    producers = roots[0].deps[0].tasks()
    for p in producers:
        wait_state(p, RUNNING)
    assert all(r.state() == INIT for r in roots)  # consumers wait for the whole phase
    for p in producers[:-1]:
        p.set(OK)
    time.sleep(0.1)
    assert all(r.state() == INIT for r in roots)
    producers[-1].set(OK)
    for r in roots:
        wait_state(r, RUNNING)
        r.set(OK)
    th.join(5)
    assert res["err"] is None and not th.is_alive()


def test_resubmit_lost_task():
    roots = simple_graph(1, 1)
    ex = ManualExecutor()
    th, res = start_eval(ex, roots)
    t = roots[0]
    wait_state(t, RUNNING)
    t.set(LOST)
    deadline = time.time() + 5
    while len(ex.ran) < 2 and time.time() < deadline:
        time.sleep(0.01)
    assert len(ex.ran) == 2
    wait_state(t, RUNNING)
    t.set(OK)
    th.join(5)
    assert res["err"] is None


@pytest.mark.parametrize("concurrent", [1, 5])
def test_resubmit_lost_interior_task(concurrent):
    roots = simple_graph(2, 2)
    ex = ManualExecutor()
    evals = [start_eval(ex, roots) for _ in range(concurrent)]
    producers = roots[0].deps[0].tasks()
    for p in producers:
        wait_state(p, RUNNING)
        p.set(OK)
    for r in roots:
        wait_state(r, RUNNING)
    # The consumer discovers its input is gone: producer lost, consumer lost.
    producers[0].set(LOST)
    roots[0].set(LOST)
    deadline = time.time() + 5
    while time.time() < deadline and producers[0].state() != RUNNING:
        time.sleep(0.01)
    assert producers[0].state() == RUNNING  # recomputed before the consumer
    assert roots[0].state() in (LOST, INIT)
    producers[0].set(OK)
    wait_state(roots[0], RUNNING)
    roots[0].set(OK)
    roots[1].set(OK)
    for th, res in evals:
        th.join(5)
        assert not th.is_alive() and res["err"] is None
    assert ex.ran.count(producers[0]) == 2  # exactly one re-run despite N evals


def test_persistent_task_loss_gives_too_many_tries():
    roots = simple_graph(1, 1)

    class AlwaysLost(Executor):
        n = 0

        def run(self, task):
            AlwaysLost.n += 1
            task.set(RUNNING)
            task.set(LOST)
    with pytest.raises(errors.Error) as ei:
        evaluate(AlwaysLost(), roots)
    assert errors.is_kind(ei.value, errors.Kind.TooManyTries)
    assert AlwaysLost.n == ev.MAX_CONSECUTIVE_LOST


def test_multi_phase_eval_many_shards():
    nshard, stages = 200, 6
    roots = simple_graph(nshard, stages)

    class Instant(Executor):
        def run(self, task):
            task.set(RUNNING)
            for d in task.deps:
                assert all(t.state() == OK for t in d.tasks())
            task.set(OK)
    t0 = time.time()
    evaluate(Instant(), roots)
    assert all(t.state() == OK for t in all_tasks(roots))
    assert len(all_tasks(roots)) == nshard * stages
    assert time.time() - t0 < 60


def test_eval_stress_random_loss():
    """Random delays, random LOST (<=20%), spontaneous post-success loss; every
    root must have been OK at completion (evalstress_test.go)."""
    rng = random.Random(0)
    roots = simple_graph(12, 4)
    lock = threading.Lock()

    class Stress(Executor):
        def run(self, task):
            task.set(RUNNING)
            time.sleep(rng.random() * 0.002)
            for d in task.deps:
                if any(t.state() != OK for t in d.tasks()):
                    task.set(LOST)
                    return
            with lock:
                r = rng.random()
            if r < 0.15:
                task.set(LOST)
                return
            task.set(OK)
            if r > 0.97 and task.deps:  # spontaneous loss of an input after success
                victim = task.deps[0].tasks()[0]
                if victim.state() == OK:
                    victim.set(LOST)

        def max_workers(self):
            return 16
    orig = ev.MAX_CONSECUTIVE_LOST
    ev.MAX_CONSECUTIVE_LOST = 1000
    try:
        evaluate(Stress(), roots)
    finally:
        ev.MAX_CONSECUTIVE_LOST = orig
    assert all(r.state() == OK for r in roots)


def test_py_and_native_core_agree():
    from bigslice_b200 import _native
    core = ev._PyCore()
    core.add_group(1, [10, 11])
    core.wait_on(20, 1)
    core.set_count(20, 1)
    core.set_count(10, 0)
    core.set_count(11, 0)
    assert sorted(core.take_todo()) == [10, 11]
    core.done_ok(10)
    assert core.take_todo() == []
    core.done_ok(11)
    assert core.take_todo() == [20] and not core.has_group(1)
    nat = _native.lib()
    if nat is not None and hasattr(nat, "EvalCore"):
        c = nat.EvalCore()
        c.add_group(1, [10, 11]); c.wait_on(20, 1); c.set_count(20, 1)  # noqa: E702
        c.set_count(10, 0); c.set_count(11, 0)  # noqa: E702
        assert sorted(c.take_todo()) == [10, 11]
        c.done_ok(10); assert c.take_todo() == []  # noqa: E702
        c.done_ok(11); assert c.take_todo() == [20] and not c.has_group(1)  # noqa: E702
