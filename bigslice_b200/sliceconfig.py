"""Session configuration from flags and a profile file.

``parse()`` reads command-line flags (``-local``, ``-device``, ``-spmd``,
``-parallelism``, ``-max-load``, ``-trace-path``, ``-status``, ``-http``,
``-set key=value`` …) layered over the profile ``$HOME/.bigslice/config``
(``key = value`` lines, keys prefixed ``bigslice.``), starts the session and,
when ``-http`` is given, serves ``/debug/*`` (status, tasks, trace).

Capability parity: reference ``sliceconfig/sliceconfig.go:39-65`` and
``exec/config.go:15-40`` (params parallelism [1024], system, eventer, max-load
[0.95], trace-path); hidden tunables
``-bigslice-internal-default-chunk-rows`` / ``-…-sort-canary-rows``
(``internal/defaultsize/size.go:15-19``).
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import Dict, List, Optional, Sequence, Tuple

from . import exec as bsexec
from .exec.graph import DebugMux
from .internal import defaultsize
from .status import Status

#: Default profile location (``sliceconfig.Path``).
Path = os.path.join(os.path.expanduser("~"), ".bigslice", "config")

DEFAULTS = {"parallelism": "1024", "system": "local", "max-load": "0.95", "trace-path": "",
            "device": "", "shuffle": "auto", "arena-bytes": str(1 << 30)}


def load_profile(path: Optional[str] = None) -> Dict[str, str]:
    """Parse ``key = value`` lines (``#`` comments); ``bigslice.`` prefix optional."""
    path = path or Path
    out: Dict[str, str] = {}
    if not os.path.exists(path):
        return out
    with open(path) as fh:
        for line in fh:
            line = line.split("#", 1)[0].strip()
            if not line or "=" not in line:
                continue
            k, v = (x.strip() for x in line.split("=", 1))
            if k.startswith("bigslice."):
                k = k[len("bigslice."):]
            out[k] = v.strip('"')
    return out


def options_from(params: Dict[str, str], status: Optional[Status] = None) -> List:
    """Translate resolved parameters into session options (``exec/config.go``)."""
    opts: List = []
    system = params.get("system", "local")
    device = params.get("device") or None
    if system in ("local", ""):
        opts.append(bsexec.LocalDevice(device) if device else bsexec.Local)
    elif system in ("spmd", "bigmachine", "gpu"):
        opts.append(bsexec.Spmd(device=device or "auto", shuffle=params.get("shuffle", "auto"),
                                arena_bytes=int(params.get("arena-bytes", 1 << 30))))
    else:
        raise ValueError("sliceconfig: unknown system %r" % system)
    opts.append(bsexec.Parallelism(int(params.get("parallelism", "1024"))))
    opts.append(bsexec.MaxLoad(float(params.get("max-load", "0.95"))))
    if params.get("trace-path"):
        opts.append(bsexec.TracePath(params["trace-path"]))
    if status is not None:
        opts.append(bsexec.Status(status))
    return opts


def parse(argv: Optional[Sequence[str]] = None, profile: Optional[str] = None
          ) -> Tuple["bsexec.Session", List[str]]:
    """Configure and start a session; returns ``(session, remaining_args)``."""
    ap = argparse.ArgumentParser(add_help=False)
    ap.add_argument("-local", "--local", action="store_true",
                    help="run in-process on the local executor")
    ap.add_argument("-spmd", "--spmd", action="store_true", help="one rank per GPU (torchrun)")
    ap.add_argument("-device", "--device", default=None)
    ap.add_argument("-parallelism", "--parallelism", default=None)
    ap.add_argument("-max-load", "--max-load", default=None)
    ap.add_argument("-trace-path", "--trace-path", default=None)
    ap.add_argument("-http", "--http", default=None, help="host:port for /debug endpoints")
    ap.add_argument("-status", "--status", action="store_true")
    ap.add_argument("-set", "--set", action="append", default=[], metavar="key=value")
    ap.add_argument("-bigslice-internal-default-chunk-rows", type=int, default=None)
    ap.add_argument("-bigslice-internal-default-sort-canary-rows", type=int, default=None)
    ns, rest = ap.parse_known_args(list(sys.argv[1:] if argv is None else argv))
    params = dict(DEFAULTS)
    params.update(load_profile(profile))
    for kv in ns.set:
        k, _, v = kv.partition("=")
        if k.startswith("bigslice."):
            k = k[len("bigslice."):]
        params[k] = v
    if ns.local:
        params["system"] = "local"
        params.setdefault("parallelism", "1024")
    if ns.spmd:
        params["system"] = "spmd"
    for k in ("device", "parallelism", "max-load", "trace-path"):
        v = getattr(ns, k.replace("-", "_"))
        if v is not None:
            params[k] = v
    if ns.local and int(params["parallelism"]) > (os.cpu_count() or 8) * 4:
        params["parallelism"] = str((os.cpu_count() or 8))
    if ns.bigslice_internal_default_chunk_rows:
        defaultsize.set_chunk(ns.bigslice_internal_default_chunk_rows)
    if ns.bigslice_internal_default_sort_canary_rows:
        defaultsize.set_sort_canary(ns.bigslice_internal_default_sort_canary_rows)
    status = Status() if (ns.status or ns.http) else None
    sess = bsexec.start(*options_from(params, status))
    if ns.http:
        host, _, port = ns.http.rpartition(":")
        mux = DebugMux()
        sess.handle_debug(mux)
        sess.debug_server = mux.serve(host or "127.0.0.1", int(port))
    return sess, rest


Parse = parse
