"""Worker for the rank-death test: rank 1 dies (os._exit, no goodbye) right
before its second exchange; rank 0 must see a TaskLost-classified failure within
the exchange timeout budget instead of hanging."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bigslice_b200 as bs  # noqa: E402
from bigslice_b200.models import wordcount as wc  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    dev = os.environ.get("BS_TEST_DEVICE", "cpu")
    sess = bs.start(bs.Spmd(device=None if dev == "cpu" else "cuda", arena_bytes=16 << 20,
                            exchange_timeout=float(os.environ.get("BS_TEST_TIMEOUT", "3"))),
                    bs.Parallelism(2))
    ex = sess.executor
    n = 20000
    k = torch.arange(n) % 1000 + rank
    v = torch.ones(n, dtype=torch.int64)
    if dev != "cpu":
        k, v = k.cuda(), v.cuda()
    e = (k[:0], v[:0])
    wc.register_source("rows", wc.DeviceRows([(k, v) if s == rank else e for s in range(ex.world)]))
    first = sess.run(wc.wordcount, "rows")
    ok_first = sum(len(f) for f in first.frames()) > 0
    runs = {"n": 0}

    def hook(stage, name):
        if stage == "before_exchange":
            runs["n"] += 1
            if rank == 1:
                os._exit(17)       # the rank is gone: no exception, no cleanup
    ex.fault_hook = hook
    t0 = time.time()
    out = {"rank": rank, "ok_first": ok_first}
    try:
        res = sess.run(wc.wordcount, "rows")
        sum(len(f) for f in res.frames())
        out["error"] = None
    except bs.errors.Error as e2:
        out["error"] = str(e2)[:300]
        out["kind"] = e2.kind.name
        out["fatal"] = bs.errors.is_fatal(e2)
        out["lost_classified"] = isinstance(e2, (bs.errors.TaskLostError, bs.errors.TooManyTriesError)) \
            or "lost" in str(e2)
    out["seconds"] = time.time() - t0
    with open(os.environ["BS_TEST_OUT"] + ".%d" % rank, "w") as fh:
        json.dump(out, fh)
    os._exit(0)   # peers are dead: skip the collective shutdown


if __name__ == "__main__":
    main()
