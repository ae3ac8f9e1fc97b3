"""Store / retry-reader tests (model: reference exec/store_test.go,
exec/bigmachine_test.go:TestReadRetries)."""
import os

import numpy as np
import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200 import errors, sliceio, slicetest
from bigslice_b200.exec.local import LocalExecutor
from bigslice_b200.exec.session import _set_executor
from bigslice_b200.exec.store import FileStore, MemoryStore
from bigslice_b200.exec.task import TaskName
from bigslice_b200.frame import Frame


def frames():
    return [Frame.of(torch.arange(i * 10, i * 10 + 10), np.array(["s%d" % j for j in range(10)], dtype=object))
            for i in range(3)]


@pytest.mark.parametrize("kind", ["memory", "file"])
def test_store_commit_open_stat_discard(kind, tmp_path):
    store = MemoryStore() if kind == "memory" else FileStore(str(tmp_path / "st"))
    name = TaskName(1, "inv1_const_map", 2, 4)
    fs = frames()
    with pytest.raises(errors.Error) as ei:
        store.stat(name, 0)
    assert errors.is_kind(ei.value, errors.Kind.NotExist)
    w = store.create(name, 0)
    for f in fs:
        w.write(f)
    w.commit()
    w2 = store.create(name, 1)
    w2.write(fs[0])
    w2.discard()  # never becomes visible
    info = store.stat(name, 0)
    assert info.records == 30 and info.size > 0
    got = sliceio.read_all(store.open(name, 0, fs[0].typ), fs[0].typ)
    assert got.col(0).tolist() == list(range(30))
    part = sliceio.read_all(store.open(name, 0, fs[0].typ, offset=17), fs[0].typ)
    assert part.col(0).tolist() == list(range(17, 30))  # resumable reads
    with pytest.raises(errors.Error):
        store.open(name, 1, fs[0].typ)
    store.discard(name, 0)
    with pytest.raises(errors.Error):
        store.stat(name, 0)


def test_retry_reader_recovers_from_transient_failures():
    f = Frame.of(torch.arange(100))
    state = {"opens": 0, "reads": 0}

    class Flaky(sliceio.Reader):
        def __init__(self, off):
            self.r = sliceio.FrameReader(f.slice(off, 100))

        def read(self, n=None):
            state["reads"] += 1
            if state["reads"] % 3 != 0:  # fails 2 of every 3 reads
                raise IOError("transient")
            return self.r.read(10)

    def opener(off):
        state["opens"] += 1
        return Flaky(off)
    r = sliceio.RetryReader(opener, retries=5, backoff=0.0)
    got = sliceio.read_all(r, f.typ)
    assert got.col(0).tolist() == list(range(100)) and state["opens"] > 5
    always = sliceio.RetryReader(lambda off: sliceio.ErrReader(IOError("down")), retries=2, backoff=0.0)
    with pytest.raises(errors.Error) as ei:
        always.read()
    assert errors.is_kind(ei.value, errors.Kind.TooManyTries)
    fatal = sliceio.RetryReader(lambda off: sliceio.ErrReader(errors.fatal("bad")), backoff=0.0)
    with pytest.raises(errors.Error) as ei:
        fatal.read()
    assert errors.is_fatal(ei.value)


def test_local_executor_with_file_store(tmp_path):
    store = FileStore(str(tmp_path / "out"))
    sess = bs.start(_set_executor(lambda s: LocalExecutor(s.p, None, store)), bs.Parallelism(3))
    keys = torch.arange(1000) % 17
    s = bs.Reduce(bs.Const(4, keys, torch.ones(1000, dtype=torch.int64)), bs.ops.add)
    res = slicetest.run(s, session=sess)
    assert dict(res.rows()) == {k: int((keys == k).sum()) for k in range(17)}
    nfiles = sum(len(fs) for _, _, fs in os.walk(tmp_path / "out"))
    assert nfiles >= 4 * 4 + 4  # producer partitions + results live in files
    assert sess.executor.stored_bytes() == 0  # nothing resident in memory
    res.discard()
    assert sum(len(fs) for _, _, fs in os.walk(tmp_path / "out")) == 0
    assert dict(res.rows())  # recomputed from lineage
