"""Resource-manager tests (model: reference exec/slicemachine_test.go:20-345:
load balancing, exclusive, probation + timeout, lost machines, priority)."""
import threading
import time

import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200 import errors, slicetest
from bigslice_b200.exec import devicemanager as dm
from bigslice_b200.exec.devicemanager import DeviceManager, MultiDevice, MultiDeviceExecutor
from bigslice_b200.exec.task import TaskState

CPU = torch.device("cpu")


def test_offers_balance_load_and_respect_capacity():
    m = DeviceManager([CPU] * 3, procs_per_device=2)
    got = [m.offer(1, 1) for _ in range(6)]
    assert sorted(d.index for d in got) == [0, 0, 1, 1, 2, 2]
    assert [d.curprocs for d in m.devs] == [2, 2, 2]
    with pytest.raises(errors.Error):
        m.offer(1, 1, timeout=0.05)  # full
    m.done(got[0], 1)
    d = m.offer(1, 1, timeout=1)
    assert d is got[0]


def test_priority_and_first_fit():
    m = DeviceManager([CPU], procs_per_device=4)
    hold = m.offer(1, 3)
    order = []

    def want(pri, procs, tag):
        d = m.offer(pri, procs)
        order.append(tag)
        time.sleep(0.05)
        m.done(d, procs)
    ths = [threading.Thread(target=want, args=a) for a in
           [(5, 4, "big-late"), (2, 4, "big-early"), (9, 1, "small")]]
    [t.start() for t in ths]
    time.sleep(0.2)
    assert order == ["small"]  # the only request that fits is not starved by bigger ones
    m.done(hold, 3)
    [t.join(5) for t in ths]
    assert order == ["small", "big-early", "big-late"]  # lower invocation index first


def test_probation_timeout_and_early_release(monkeypatch):
    monkeypatch.setattr(dm, "ProbationTimeout", 0.2)
    m = DeviceManager([CPU, CPU], procs_per_device=1)
    a = m.offer(1, 1)
    m.done(a, 1, IOError("device trouble"))
    assert m.devs[a.index].health == dm.PROBATION
    b = m.offer(1, 1)
    assert b.index != a.index  # the sick device gets no work
    with pytest.raises(errors.Error):
        m.offer(1, 1, timeout=0.05)
    time.sleep(0.25)
    c = m.offer(1, 1, timeout=1)  # probation expired
    assert c.index == a.index
    m.done(c, 1)
    assert m.devs[a.index].health == dm.OK


def test_multidevice_executor_spreads_and_recovers_from_device_loss():
    keys = torch.arange(4000) % 37
    sess = bs.start(MultiDevice([CPU] * 4, procs_per_device=1), bs.Parallelism(4))
    ex = sess.executor
    s = bs.Reduce(bs.Const(8, keys, torch.ones(4000, dtype=torch.int64)), bs.ops.add)
    res = slicetest.run(s, session=sess)
    want = {k: int((keys == k).sum()) for k in range(37)}
    assert dict(res.rows()) == want
    used = {d.index for d in ex.where.values()}
    assert len(used) == 4  # work was spread over all devices
    victim = next(iter(ex.where.values())).index
    lost = ex.lose_device(victim)
    assert lost > 0 and any(t.state() == TaskState.LOST for t in res.tasks) or lost > 0
    assert dict(res.rows()) == want  # recomputed on the remaining devices
    assert all(d.index != victim for d in ex.where.values())
    for i in range(4):
        ex.lose_device(i)
    with pytest.raises(errors.Error):
        slicetest.run(bs.Const(2, torch.arange(4)), session=sess)


def test_device_errors_mean_lost_and_probation_not_fatal(monkeypatch):
    monkeypatch.setattr(dm, "ProbationTimeout", 0.1)
    calls = {"n": 0}

    def flaky(dev, task):
        calls["n"] += 1
        if dev.index == 0 and calls["n"] < 4:
            raise RuntimeError("CUDA error: an illegal memory access was encountered")
    sess = bs.start(MultiDevice([CPU, CPU], procs_per_device=1, fail_hook=flaky), bs.Parallelism(2))
    res = slicetest.run(bs.Map(bs.Const(4, torch.arange(8)), lambda x: x * 3), session=sess)
    assert res.rows() == [(i * 3,) for i in range(8)]
    assert sess.executor.mgr.devs[0].err_count >= 1

    def oom(x):
        raise torch.cuda.OutOfMemoryError("CUDA out of memory")
    _, err = slicetest.run_err(bs.Map(bs.Const(1, torch.arange(2)), oom, out=[bs.int64]), session=sess)
    assert err is not None and errors.is_kind(err, errors.Kind.TooManyTries)  # retried, never "fatal"


def test_exclusive_takes_whole_device():
    peak = [0]
    running = [0]
    lock = threading.Lock()

    def slow(x):
        with lock:
            running[0] += 1
            peak[0] = max(peak[0], running[0])
        time.sleep(0.03)
        with lock:
            running[0] -= 1
        return x
    sess = bs.start(MultiDevice([CPU], procs_per_device=4), bs.Parallelism(4))
    slicetest.run(bs.Map(bs.Const(6, torch.arange(6)), slow, bs.Exclusive, out=[bs.int64]), session=sess)
    assert peak[0] == 1


@pytest.mark.gpu
def test_multidevice_executor_on_real_gpus():
    """The resource manager on actual devices: frames resident in HBM, per-device
    health polling (free/total HBM), device loss -> resident tasks LOST ->
    recomputed on the surviving device.  On a one-GPU box the same device is
    registered twice (two logical 'machines' sharing it)."""
    n = torch.cuda.device_count()
    devs = [torch.device("cuda", i) for i in range(min(n, 2))]
    if len(devs) == 1:
        devs = devs * 2
    ex = MultiDeviceExecutor(devs, procs_per_device=2)
    keys = torch.arange(200000) % 1000
    vals = torch.ones(200000, dtype=torch.int64)

    @bs.Func
    def job():
        return bs.Reduce(bs.Map(bs.Const(8, keys, vals), lambda k, v: (k % 100, v)), bs.ops.add)
    from bigslice_b200.exec.session import _set_executor
    with bs.start(_set_executor(lambda s: ex), bs.Parallelism(4)) as sess:
        r = sess.run(job)
        assert all(f.device.type == "cuda" for f in r.frames())
        assert dict(r.rows()) == {k: 2000 for k in range(100)}
        polled = ex.mgr.poll()
        assert len(polled) == 2 and all(p["hbm_total"] > 0 and p["hbm_free"] > 0 for p in polled)
        assert any("hbm" in line for line in ex.mgr.status())
        lost = ex.lose_device(0)          # everything resident on device 0 is gone
        assert lost >= 0
        r2 = sess.run(job)                # recomputed on what is left
        assert dict(r2.rows()) == {k: 2000 for k in range(100)}
        assert ex.mgr.devs[0].health == dm.LOST
