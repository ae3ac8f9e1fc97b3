"""slicer: self-verifying integration / stress workloads.

    python -m bigslice_b200.cmd.slicer [session flags] <cogroup|reduce|memiter|oom> [options]

* ``cogroup`` — N sharded inputs with deterministic keys; verifies every key's
  group sizes after the join.
* ``reduce``  — sum per key over generated rows; verifies each key's total.
* ``memiter`` — repeated runs that reuse and discard results; reports process /
  device memory so leaks show up.
* ``oom``     — allocates until the executor fails, to exercise error paths.

Capability parity: reference ``cmd/slicer/main.go:18-73``, ``cogroup.go``,
``reduce.go``, ``memiter.go``, ``oom.go``.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

import bigslice_b200 as bs
from bigslice_b200 import sliceconfig


@bs.Func
def _cogroup_func(nshard: int, nkey: int, ninput: int):
    ins = []
    def make_read(i):
        def read(shard, state, keys: bs.int64, vals: bs.int64):
            # shard s of input i emits keys s, s+nshard, ... each (i+1) times
            if getattr(state, "done", False):
                return 0, bs.EOF
            state.done = True
            ks = torch.arange(shard, nkey, nshard).repeat_interleave(i + 1)
            n = ks.numel()
            keys[:n] = ks
            vals[:n] = i
            return n, bs.EOF
        return read
    for i in range(ninput):
        ins.append(bs.ReaderFunc(nshard, make_read(i), batch=(nkey // nshard + 1) * (i + 1)))
    return bs.Cogroup(*ins)


@bs.Func
def _reduce_func(nshard: int, nkey: int, rows_per_shard: int):
    def read(shard, state, keys: bs.int64, vals: bs.int64):
        if getattr(state, "done", False):
            return 0, bs.EOF
        state.done = True
        g = torch.Generator().manual_seed(shard)
        keys[:rows_per_shard] = torch.randint(0, nkey, (rows_per_shard,), generator=g)
        vals[:rows_per_shard] = 1
        return rows_per_shard, bs.EOF
    return bs.Reduce(bs.ReaderFunc(nshard, read, batch=rows_per_shard), bs.ops.add)


@bs.Func
def _memiter_base(nshard: int, n: int):
    return bs.Const(nshard, torch.arange(n), torch.ones(n, dtype=torch.int64))


@bs.Func
def _memiter_step(prev: bs.Slice):
    return bs.Reduce(bs.Map(prev, lambda k, v: ((k * 7 + 1) % 100003, v)), bs.ops.add)


@bs.Func
def _oom_func(nshard: int, rows: int):
    return bs.Map(bs.Const(nshard, torch.arange(nshard)),
                  lambda x: torch.zeros(rows, dtype=torch.int64)[: len(x)] + x)


def _gather(sess, res):
    g = getattr(sess.executor, "gather", None)
    return g(res) if g is not None else res.rows()


def cmd_cogroup(sess, a) -> int:
    res = sess.run(_cogroup_func, a.nshard, a.nkey, a.ninput)
    rows = _gather(sess, res)
    bad = 0
    seen = set()
    for row in rows:
        k, groups = row[0], row[1:]
        seen.add(k)
        for i, g in enumerate(groups):
            if len(g) != i + 1 or any(v != i for v in g):
                bad += 1
    missing = a.nkey - len(seen)
    print("cogroup: %d keys, %d bad groups, %d missing" % (len(seen), bad, missing))
    return 1 if (bad or missing) else 0


def cmd_reduce(sess, a) -> int:
    res = sess.run(_reduce_func, a.nshard, a.nkey, a.rows)
    rows = _gather(sess, res)
    total = sum(v for _, v in rows)
    want = a.nshard * a.rows
    ok = total == want and len(set(k for k, _ in rows)) == len(rows)
    print("reduce: %d keys, total %d (want %d): %s" % (len(rows), total, want, "ok" if ok else "FAIL"))
    return 0 if ok else 1


def cmd_memiter(sess, a) -> int:
    import psutil
    proc = psutil.Process(os.getpid())
    base = sess.run(_memiter_base, a.nshard, a.rows)
    for i in range(a.iters):
        r = sess.run(_memiter_step, base)
        n = len(_gather(sess, r))
        r.discard()
        dev = torch.cuda.memory_allocated() if torch.cuda.is_available() else 0
        print("iter %d: %d keys rss=%.1fMB device=%.1fMB" % (i, n, proc.memory_info().rss / 1e6, dev / 1e6))
    return 0


def cmd_oom(sess, a) -> int:
    rows = a.rows
    while True:
        try:
            sess.run(_oom_func, a.nshard, rows).discard()
        except BaseException as e:  # noqa: BLE001
            print("oom: failed at %d rows/task: %s" % (rows, type(e).__name__))
            return 0
        rows *= 2
        if rows > (1 << 40):
            return 1


def main(argv=None) -> int:
    sess, rest = sliceconfig.parse(argv)
    ap = argparse.ArgumentParser(prog="slicer")
    sub = ap.add_subparsers(dest="cmd", required=True)
    for name in ("cogroup", "reduce", "memiter", "oom"):
        p = sub.add_parser(name)
        p.add_argument("--nshard", type=int, default=8)
        p.add_argument("--nkey", type=int, default=1 << 14)
        p.add_argument("--ninput", type=int, default=3)
        p.add_argument("--rows", type=int, default=1 << 16)
        p.add_argument("--iters", type=int, default=5)
    a = ap.parse_args(rest)
    t0 = time.time()
    try:
        rc = {"cogroup": cmd_cogroup, "reduce": cmd_reduce, "memiter": cmd_memiter,
              "oom": cmd_oom}[a.cmd](sess, a)
    finally:
        sess.shutdown()
    print("%s: %.2fs" % (a.cmd, time.time() - t0))
    return rc


if __name__ == "__main__":
    sys.exit(main())
