"""utils: host span recorder, topology binding (no GPU needed)."""
import os
import threading
import time

from bigslice_b200.utils import spans, topology


def test_spans_record_and_summarise():
    spans.enable(True)
    try:
        m = spans.mark()
        with spans.span("outer"):
            time.sleep(0.01)
            with spans.span("inner"):
                time.sleep(0.005)

        def worker():
            with spans.span("inner"):
                time.sleep(0.002)
        t = threading.Thread(target=worker, name="w1")
        t.start()
        t.join()
        rows = {r["span"]: r for r in spans.summary(m, per=2)}
        assert rows["inner"]["count_per"] == 1.0 and rows["outer"]["count_per"] == 0.5
        assert rows["outer"]["ms_per"] >= 7.0 and rows["inner"]["ms_per"] >= 3.0
        assert {th for (_n, th, _a, _b) in spans.raw(m)} >= {"w1"}
        assert "outer" in spans.format_summary(m)
    finally:
        spans.enable(False)
    m = spans.mark()
    with spans.span("ignored"):
        pass
    assert spans.raw(m) == []


def test_topology_binding_is_safe_without_nvml():
    before = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    cpus = topology.bind_to_gpu(0)
    if not cpus and before is not None:
        assert os.sched_getaffinity(0) == before  # untouched when NVML has no answer
    os.environ["BIGSLICE_NO_NUMA_BIND"] = "1"
    try:
        assert topology.bind_to_gpu(0) == []
    finally:
        del os.environ["BIGSLICE_NO_NUMA_BIND"]
        if before is not None:
            os.sched_setaffinity(0, before)


def test_status_board_vote_and_diverged_exchange_sequences():
    """Two 'ranks' (threads) vote through the shared-memory board.  Equal exchange
    tags agree and return every rank's integer; a rank that repeats an exchange on
    its own (attempt 2 against the peer's attempt 1 of the next group) makes the
    vote fail on BOTH sides instead of pairing two different exchanges."""
    import pytest
    from bigslice_b200 import errors
    from bigslice_b200.exec.spmd import SpmdExecutor
    from bigslice_b200.utils.shmboard import Board

    class Name:
        def __init__(self, inv, s):
            self.inv_index, self.s = inv, s

        def __str__(self):
            return self.s

    class Head:
        def __init__(self, inv, s):
            self.name = Name(inv, s)

    name = "bigslice-test-board-%d" % os.getpid()
    exs = []
    for r in range(2):
        ex = SpmdExecutor(device=None)
        ex.rank, ex.world, ex.exchange_timeout = r, 2, 5.0
        ex.board = Board(name, 2, r, create=(r == 0))
        exs.append(ex)
    try:
        G, H = Head(3, "inv3_reduce_0"), Head(3, "inv3_reduce_1")
        assert exs[0]._tag_exchange(7, G) == exs[1]._tag_exchange(7, G)     # same exchange, same tag
        assert exs[0]._tag_exchange(7, G) != exs[1]._tag_exchange(7, H)     # attempt 2 of G vs attempt 1 of H
        assert exs[0]._tag_exchange(7, H) != exs[1]._tag_exchange(7, G)     # (both ranks have now issued G, G, H)

        def vote(ex, acc, head, extra, out):
            try:
                out.append(ex._board_vote(0, acc, [], head, extra=extra))
            except errors.Error as e:
                out.append(e)
        # in step: both at attempt 3 of G
        outs = [[], []]
        accs = [ex._tag_exchange(0, G) for ex in exs]
        ts = [threading.Thread(target=vote, args=(exs[r], accs[r], G, 100 + r, outs[r])) for r in range(2)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert outs[0] == outs[1] == [[100, 101]]
        # diverged: rank 1 repeats G on its own while rank 0 moved on to H
        outs = [[], []]
        accs = [exs[0]._tag_exchange(0, H), exs[1]._tag_exchange(0, G)]
        heads = [H, G]
        ts = [threading.Thread(target=vote, args=(exs[r], accs[r], heads[r], 0, outs[r])) for r in range(2)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        for o in outs:
            assert isinstance(o[0], errors.Error) and "diverged" in str(o[0])
            with pytest.raises(errors.Error):
                raise o[0]
    finally:
        for r in (1, 0):
            exs[r].board.close(unlink=(r == 0))


def test_a_failed_exchange_attempt_is_repeated_once_however_many_consumers_hold_its_frames():
    import threading as th
    from bigslice_b200.exec.spmd import SpmdExecutor
    from bigslice_b200.exec.task import TaskState

    class T:
        def __init__(self):
            self.st, self.group = TaskState.OK, None

        def state(self):
            return self.st

        def set(self, s):
            self.st = s

    ex = SpmdExecutor(device=None)
    prods = [T(), T()]
    for p in prods:
        p.group = prods
    gid = id(prods[0])
    attempt1, attempt2 = object(), object()
    ex.group_done[gid] = th.Event()
    ex._group_pend[gid] = attempt1
    ex.exchanged[(gid, 0)] = ["frames of attempt 1"]
    # first consumer of the failed attempt: the group is forgotten, producers are LOST
    ex._mark_group_lost(gid, prods[0], attempt1)
    assert gid not in ex.group_done and (gid, 0) not in ex.exchanged
    assert all(p.state() == TaskState.LOST for p in prods)
    # the evaluator re-runs the producers; the repeated exchange is queued, then in place
    for p in prods:
        p.set(TaskState.OK)
    ex.group_done[gid] = th.Event()
    ex._mark_group_lost(gid, prods[0], attempt1)           # stale frame while the retry is queued
    assert gid in ex.group_done and all(p.state() == TaskState.OK for p in prods)
    ex._group_pend[gid] = attempt2
    ex.exchanged[(gid, 0)] = ["frames of attempt 2"]
    ex._mark_group_lost(gid, prods[0], attempt1)           # stale frame after the retry completed
    assert ex.exchanged[(gid, 0)] == ["frames of attempt 2"] and all(p.state() == TaskState.OK for p in prods)
    # the new attempt failing is a new loss
    ex._mark_group_lost(gid, prods[0], attempt2)
    assert gid not in ex.group_done and all(p.state() == TaskState.LOST for p in prods)
