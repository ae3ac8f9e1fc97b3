"""Fault-injection tests on real combinators (model: reference
exec/bigmachine_test.go:TestBigmachineExecutorLost, exec/session_test.go:
TestScanFaultTolerance/TestDiscard, exec/chaosmonkey_test.go)."""
import random
import threading
import time

import numpy as np
import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200 import errors, slicetest
from bigslice_b200.exec import eval as ev
from bigslice_b200.exec.chaos import Chaos, ChaosExecutor
from bigslice_b200.exec.task import TaskState

N = 3000
rng = random.Random(3)
KEYS = torch.tensor([rng.randrange(64) for _ in range(N)])
WANT = {}
for k in KEYS.tolist():
    WANT[k] = WANT.get(k, 0) + 1


@bs.Func
def counts(nshard: int):
    return bs.Reduce(bs.Const(nshard, KEYS, torch.ones(N, dtype=torch.int64)), bs.ops.add)


@bs.Func
def regroup(prev: bs.Slice):
    return bs.Cogroup(bs.Map(prev, lambda k, v: (k % 8, v)))


@bs.Func
def big(nshard: int):
    return bs.Map(bs.Const(nshard, torch.arange(200000)), lambda x: x + 1)


@pytest.fixture
def chaos_session(monkeypatch):
    monkeypatch.setattr(ev, "MAX_CONSECUTIVE_LOST", 1000)
    sess = bs.start(Chaos(fail_rate=0.2, loss_rate=0.1, seed=1), bs.Parallelism(4))
    yield sess
    sess.shutdown()


def test_reduce_and_cogroup_survive_random_loss(chaos_session):
    res = chaos_session.run(counts, 6)
    assert dict(res.rows()) == WANT
    res2 = chaos_session.run(regroup, res)
    got = {k: sum(v) for k, v in res2.rows()}
    want = {}
    for k, c in WANT.items():
        want[k % 8] = want.get(k % 8, 0) + c
    assert got == want
    ex = chaos_session.executor
    assert ex.injected_failures > 0  # faults really happened


def test_machine_kill_then_reuse_recomputes(chaos_session):
    res = chaos_session.run(counts, 4)
    lost = chaos_session.executor.kill()
    assert lost > 0 and all(t.state() == TaskState.LOST for t in res.tasks)
    assert dict(res.rows()) == WANT  # scanning re-evaluates the lost lineage
    chaos_session.executor.kill()
    res2 = chaos_session.run(regroup, res)  # a dependent run recomputes its inputs
    assert sum(sum(v) for _, v in res2.rows()) == N


def test_scan_fault_tolerance_kill_while_scanning(monkeypatch):
    monkeypatch.setattr(ev, "MAX_CONSECUTIVE_LOST", 1000)
    sess = bs.start(Chaos(fail_rate=0.0, loss_rate=0.0), bs.Parallelism(4))
    res = sess.run(big, 8)
    stop = threading.Event()

    def killer():
        while not stop.is_set():
            sess.executor.kill()
            time.sleep(0.005)
    th = threading.Thread(target=killer, daemon=True)
    th.start()
    try:
        total = 0
        n = 0
        for (v,) in res.scanner():
            total += v
            n += 1
    finally:
        stop.set()
        th.join()
    assert n == 200000 and total == sum(range(1, 200001))
    assert sess.executor.kills > 1


def test_discard_chaos_race(monkeypatch):
    """Run, discard and kill concurrently, then verify (TestDiscardChaos)."""
    monkeypatch.setattr(ev, "MAX_CONSECUTIVE_LOST", 1000)
    sess = bs.start(Chaos(fail_rate=0.05, loss_rate=0.05, seed=5), bs.Parallelism(4))
    errs = []
    stop = time.time() + 3.0

    def worker(i):
        try:
            while time.time() < stop:
                r = sess.run(counts, 3 + i)
                if i % 2:
                    r.discard()
                assert dict(r.rows()) == WANT
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
    ths = [threading.Thread(target=worker, args=(i,)) for i in range(3)]
    [t.start() for t in ths]
    while time.time() < stop:
        sess.executor.kill()
        time.sleep(0.05)
    [t.join() for t in ths]
    assert not errs, errs[:1]
    assert dict(sess.run(counts, 5).rows()) == WANT


def test_persistent_failure_is_too_many_tries():
    sess = bs.start(Chaos(fail_rate=1.0, loss_rate=0.0), bs.Parallelism(2))
    with pytest.raises(errors.Error) as ei:
        sess.run(counts, 2)
    assert errors.is_kind(ei.value, errors.Kind.TooManyTries)
