"""Debug endpoints: task graph, trace and status over HTTP.

``DebugMux`` is a tiny path→handler table that can be served with
``serve()`` (``http.server``) or queried directly in tests.  Sessions register

* ``/debug/tasks/graph`` — JSON task graph (nodes with state, links),
* ``/debug/tasks``       — a self-contained HTML page rendering that graph,
* ``/debug/trace``       — the Chrome trace JSON so far,
* ``/debug/status``      — text status.

Capability parity: reference ``exec/graph.go:15-211`` and
``exec/session.go:376-389``.
"""
from __future__ import annotations

import json
import threading
from http.server import BaseHTTPRequestHandler, HTTPServer
from typing import Callable, Dict, List, Tuple

from .task import Task, iter_tasks


class DebugMux:
    def __init__(self):
        self.handlers: Dict[str, Callable[[], Tuple[str, str]]] = {}

    def handle(self, path: str, fn: Callable[[], Tuple[str, str]]) -> None:
        """Register ``fn() -> (content_type, body)`` at ``path``."""
        self.handlers[path] = fn

    def get(self, path: str) -> Tuple[str, str]:
        return self.handlers[path]()

    def serve(self, host: str = "127.0.0.1", port: int = 0) -> HTTPServer:
        mux = self

        class H(BaseHTTPRequestHandler):
            def do_GET(self):
                path = self.path.split("?")[0]
                if path in ("/debug", "/debug/"):
                    body = "\n".join(sorted(mux.handlers)) + "\n"
                    ct = "text/plain"
                elif path in mux.handlers:
                    ct, body = mux.handlers[path]()
                else:
                    self.send_error(404)
                    return
                data = body.encode()
                self.send_response(200)
                self.send_header("Content-Type", ct)
                self.send_header("Content-Length", str(len(data)))
                self.end_headers()
                self.wfile.write(data)

            def log_message(self, *a):
                pass

        srv = HTTPServer((host, port), H)
        threading.Thread(target=srv.serve_forever, daemon=True).start()
        return srv


def task_graph_json(roots: List[Task]) -> str:
    nodes, links, index = [], [], {}
    tasks: List[Task] = []
    iter_tasks(roots, tasks.append)
    tasks.sort(key=lambda t: str(t.name))
    for i, t in enumerate(tasks):
        index[id(t)] = i
        nodes.append({"name": str(t.name), "state": str(t.state()),
                      "partitions": t.num_partition})
    for t in tasks:
        for d in t.deps:
            for dt in d.tasks():
                links.append({"source": index[id(t)], "target": index[id(dt)],
                              "partition": d.partition})
    return json.dumps({"nodes": nodes, "links": links})


_HTML = """<!doctype html><html><head><meta charset="utf-8"><title>bigslice tasks</title>
<style>body{font-family:monospace} .OK{color:green}.ERROR{color:red}.LOST{color:orange}
.RUNNING{color:blue}</style></head><body><h3>tasks</h3><div id="g"></div><script>
fetch('/debug/tasks/graph').then(r=>r.json()).then(g=>{
 const el=document.getElementById('g');
 g.nodes.forEach((n,i)=>{const d=document.createElement('div');
  const deps=g.links.filter(l=>l.source==i).map(l=>g.nodes[l.target].name+'['+l.partition+']');
  d.innerHTML='<span class="'+n.state+'">'+n.state+'</span> '+n.name+
   (deps.length?' &larr; '+deps.join(', '):'');el.appendChild(d);});});
</script></body></html>"""


def register(sess, mux: DebugMux) -> None:
    def roots():
        with sess._lock:
            return list(sess.roots.values())

    mux.handle("/debug/tasks/graph", lambda: ("application/json", task_graph_json(roots())))
    mux.handle("/debug/tasks", lambda: ("text/html", _HTML))
    mux.handle("/debug/trace", lambda: ("application/json", sess.tracer.marshal()))
    def status_text():
        """Status groups plus the executor's machine lines (per-GPU HBM, load,
        exchange counters — reference: the ``machines`` status group,
        ``exec/slicemachine.go:238-257``) and the last invocation's plan."""
        parts = []
        lines = None
        ex = sess.executor
        if hasattr(ex, "status_lines"):
            lines = ex.status_lines()
        elif hasattr(ex, "mgr"):
            lines = ex.mgr.status()
        if lines:
            parts.append("machines\n" + "\n".join("  " + ln for ln in lines))
        if sess.status_:
            parts.append(sess.status_.marshal())
        return "\n".join(parts)
    mux.handle("/debug/status", lambda: ("text/plain", status_text()))
