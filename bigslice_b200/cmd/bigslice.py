"""The ``bigslice`` command: build, run and set up.

    python -m bigslice_b200.cmd.bigslice build            # compile native kernels/runtime in-tree
    python -m bigslice_b200.cmd.bigslice run prog.py [--gpus N] [args...]
    python -m bigslice_b200.cmd.bigslice setup [--gpus N]  # write ~/.bigslice/config

``run`` executes a bigslice program: a python file that defines
``bigslice_main(sess, args)`` (a ``main`` is synthesised around it, like the
reference does for ``BigsliceMain`` packages) or an ordinary script.  With
``--gpus N`` (N>1) it launches one rank per GPU through ``torch.distributed.
run`` on 127.0.0.1.  ``setup`` writes a profile for the local box (the
reference's ``setup-ec2`` provisions cloud security groups; here the "cluster"
is the GPUs of one NVSwitch box).

Capability parity: reference ``cmd/bigslice/main.go:43-74``,
``bigslicecmd/build.go:69-221``, ``bigslicecmd/run.go:22-49``, ``ec2.go:58-199``.
"""
from __future__ import annotations

import argparse
import os
import runpy
import subprocess
import sys

from .. import sliceconfig


def cmd_build(a) -> int:
    from .. import build
    cuda, native = build.build_all(force=a.force, verbose=a.verbose)
    print(cuda)
    print(native)
    return 0


def _run_program(path: str, args) -> int:
    ns = runpy.run_path(path, run_name="__bigslice__")
    main = ns.get("bigslice_main") or ns.get("BigsliceMain")
    if main is None:
        # ordinary script: run it as __main__
        sys.argv = [path] + list(args)
        runpy.run_path(path, run_name="__main__")
        return 0
    sess, rest = sliceconfig.parse(args)
    try:
        rc = main(sess, rest)
    finally:
        sess.shutdown()
    return int(rc or 0)


def cmd_run(a) -> int:
    if a.gpus and a.gpus > 1 and "RANK" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node=%d" % a.gpus, "--master-addr", "127.0.0.1",
               "--master-port", str(a.port), "-m", "bigslice_b200.cmd.bigslice", "run",
               a.program] + (["-spmd"] if "-spmd" not in a.args else []) + a.args
        return subprocess.call(cmd)
    return _run_program(a.program, a.args)


def cmd_setup(a) -> int:
    import torch
    n = a.gpus or (torch.cuda.device_count() if torch.cuda.is_available() else 0)
    path = sliceconfig.Path
    os.makedirs(os.path.dirname(path), exist_ok=True)
    lines = ["# written by `bigslice setup`",
             "bigslice.system = %s" % ("spmd" if n > 1 else "local"),
             "bigslice.device = %s" % ("cuda" if n >= 1 else ""),
             "bigslice.parallelism = %d" % (max(n, 1) * 2),
             "bigslice.max-load = 0.95",
             "bigslice.arena-bytes = %d" % (4 << 30)]
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("wrote %s (%d GPUs)" % (path, n))
    return 0


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="bigslice")
    sub = ap.add_subparsers(dest="cmd", required=True)
    b = sub.add_parser("build")
    b.add_argument("--force", action="store_true")
    b.add_argument("-v", "--verbose", action="store_true")
    r = sub.add_parser("run")
    r.add_argument("program")
    r.add_argument("--gpus", type=int, default=0)
    r.add_argument("--port", type=int, default=29533)
    r.add_argument("args", nargs=argparse.REMAINDER)
    s = sub.add_parser("setup")
    s.add_argument("--gpus", type=int, default=0)
    sub.add_parser("setup-ec2")
    a = ap.parse_args(argv)
    if a.cmd == "build":
        return cmd_build(a)
    if a.cmd == "run":
        return cmd_run(a)
    if a.cmd == "setup":
        return cmd_setup(a)
    sys.stderr.write("setup-ec2 has no analogue on a single NVSwitch box; use `bigslice setup`\n")
    return 2


if __name__ == "__main__":
    sys.exit(main())
