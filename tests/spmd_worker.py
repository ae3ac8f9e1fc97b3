"""Worker program for multi-process SPMD tests (launched with torch.distributed.run)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bigslice_b200 as bs  # noqa: E402

N = 4000
rng = np.random.default_rng(7)
KEYS = torch.from_numpy(rng.integers(0, 97, N))
WORDS = np.array(["w%d" % (i % 13) for i in range(N)], dtype=object)
scanned = []


@bs.Func
def const_map_scan(nshard: int):
    s = bs.Const(nshard, torch.arange(100))
    s = bs.Map(s, lambda x: x * 2)

    def sink(shard, scanner):
        for (v,) in scanner:
            scanned.append(v)
    return bs.Scan(s, sink)


@bs.Func
def wordcount(nshard: int):
    s = bs.Const(nshard, KEYS, torch.ones(N, dtype=torch.int64))
    return bs.Reduce(s, bs.ops.add)


@bs.Func
def strcount(nshard: int):
    s = bs.Const(nshard, WORDS, torch.ones(N, dtype=torch.int64))
    return bs.Reduce(s, lambda a, b: a + b)


@bs.Func
def join(nshard: int):
    a = bs.Const(nshard, KEYS, torch.arange(N, dtype=torch.float32))
    b = bs.Const(nshard, torch.arange(0, 97, 2), torch.arange(0, 97, 2) * 10)
    return bs.Cogroup(a, b)


@bs.Func
def reshard(nshard: int, out: int):
    return bs.Reshard(bs.Const(nshard, KEYS, torch.arange(N)), out)


@bs.Func
def chain(prev: bs.Slice):
    return bs.Reduce(bs.Map(prev, lambda k, v: (k % 5, v)), bs.ops.add)


def main():
    device = os.environ.get("BS_TEST_DEVICE", "cpu")
    shuffle = os.environ.get("BS_TEST_SHUFFLE", "auto")
    sess = bs.start(bs.Spmd(device=None if device == "cpu" else "cuda", shuffle=shuffle,
                            arena_bytes=64 << 20), bs.Parallelism(2))
    ex = sess.executor
    world, rank = ex.world, ex.rank
    out = {"rank": rank, "world": world}
    sess.run(const_map_scan, 2 * world)
    out["scan_local"] = sorted(scanned)
    r = sess.run(wordcount, world)
    rows = ex.gather(r)
    want = {}
    for k in KEYS.tolist():
        want[k] = want.get(k, 0) + 1
    out["wordcount_ok"] = dict(rows) == want
    out["wordcount_local_keys"] = sorted(k for k, _ in r.rows())
    r2 = sess.run(chain, r)
    want2 = {}
    for k, v in want.items():
        want2[k % 5] = want2.get(k % 5, 0) + v
    out["chain_ok"] = dict(ex.gather(r2)) == want2
    rows = ex.gather(sess.run(strcount, world))
    wants = {}
    for w in WORDS.tolist():
        wants[w] = wants.get(w, 0) + 1
    out["strcount_ok"] = dict(rows) == wants
    rows = ex.gather(sess.run(join, 2 * world))
    got = {k: (sorted(a), sorted(b)) for k, a, b in rows}
    wantj = {}
    for i, k in enumerate(KEYS.tolist()):
        wantj.setdefault(k, ([], []))[0].append(float(i))
    for k in range(0, 97, 2):
        wantj.setdefault(k, ([], []))[1].append(k * 10)
    out["join_ok"] = got == {k: (sorted(a), sorted(b)) for k, (a, b) in wantj.items()}
    res = sess.run(reshard, world, 2 * world + 1)
    rows = ex.gather(res)
    out["reshard_ok"] = sorted(rows) == sorted(zip(KEYS.tolist(), range(N)))
    local = res.rows()
    from bigslice_b200.frame import hashing
    out["reshard_placement_ok"] = all(
        int(hashing.hash64(torch.tensor([k]), 0)[0]) % (2 * world + 1) % world == rank
        for k, _ in local)
    out["modes"] = sorted(set(e["mode"] for e in ex.exchange_log))
    try:
        from bigslice_b200.ops import gpu
        out["launches"] = gpu.launches() if device != "cpu" else 0
    except Exception:
        out["launches"] = 0
    sess.shutdown()
    path = os.environ["BS_TEST_OUT"] + ".%d" % rank
    with open(path, "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
