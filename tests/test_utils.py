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
