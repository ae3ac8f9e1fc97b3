"""badfuncs: negative tests for Func-registry determinism.

    python -m torch.distributed.run --nproc-per-node 2 -m bigslice_b200.cmd.badfuncs <ok|toolate|random>

* ``ok``      — all Funcs registered up front, identically on every rank.
* ``toolate`` — a Func registered *after* the session started: using it fails
  because ranks cannot have agreed on it.
* ``random``  — ranks register a different number of Funcs: session start must
  fail with a registry diff.

Capability parity: reference ``cmd/badfuncs/main.go:22-119`` (workers verify
``FuncLocations`` at boot, ``exec/slicemachine.go:689-702``).
"""
from __future__ import annotations

import os
import sys

import torch

import bigslice_b200 as bs
from bigslice_b200 import errors


@bs.Func
def base(n: int):
    return bs.Const(n, torch.arange(16))


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    mode = argv[0] if argv else "ok"
    rank = int(os.environ.get("RANK", "0"))
    if mode == "random":
        # rank-dependent registration order/number: must be rejected
        for i in range(rank + 1):
            bs.Func(lambda n, i=i: bs.Const(n, torch.arange(4) + i))
    try:
        sess = bs.start(bs.Spmd(device=None if not torch.cuda.is_available() else "cuda"))
    except errors.Error as e:
        print("rank %d: session rejected: %s" % (rank, str(e).splitlines()[0]))
        return 0 if mode == "random" else 1
    if mode == "random" and sess.executor.world > 1:
        print("rank %d: mismatched registries were not detected" % rank)
        return 1
    rc = 0
    if mode == "toolate":
        late = bs.Func(lambda n: bs.Const(n, torch.arange(4)))
        try:
            sess.run(late, 2)
            print("rank %d: late Func was not detected" % rank)
            rc = 1
        except errors.Error as e:
            print("rank %d: late Func rejected: %s" % (rank, str(e).splitlines()[0]))
    else:
        rows = sess.executor.gather(sess.run(base, 2))
        print("rank %d: ok, %d rows" % (rank, len(rows)))
    sess.shutdown()
    return rc


if __name__ == "__main__":
    sys.exit(main())
