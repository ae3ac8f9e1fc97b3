"""Build a ``csrc/*.cu`` file for the CPU against the SIMT emulator (``simt.h``).

``emulate("strings")`` returns a ``ctypes`` library exporting the *same* ``bs_*``
entry points as the CUDA build — the unmodified kernels and host wrappers of the
file, with

* ``#include "common.cuh"`` replaced by ``simt.h`` plus the portable part of
  ``csrc/common.cuh`` (everything that is not inline PTX);
* inline ``asm volatile(...)`` statements dropped (kernels that depend on them —
  TMA, multicast stores — are not runnable, the rest of the file still is);
* every ``kernel<<<grid, block, smem, stream>>>(args);`` rewritten to
  ``bs_emu_launch(grid, block, smem, [=] { kernel(args); });`` and ``extern __shared__``
  arrays pointed at the emulator's dynamic shared-memory buffer.

``asan=True`` compiles with AddressSanitizer; run such a library in a subprocess
with ``LD_PRELOAD`` set to ``asan_preload()``: every out-of-bounds access of the
device code is then a hard failure (a memcheck of the kernels without a GPU).
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "csrc")


def _match(text: str, i: int, open_: str, close: str) -> int:
    """Index just past the bracket that closes the one at ``text[i]``."""
    depth = 0
    while i < len(text):
        c = text[i]
        if c == open_:
            depth += 1
        elif c == close:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced %s%s" % (open_, close))


def _split_top(s: str):
    out, depth, cur = [], 0, ""
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += c
    out.append(cur.strip())
    return out


def portable_common() -> str:
    """csrc/common.cuh without the CUDA runtime include and without the helpers
    that are inline PTX."""
    text = open(os.path.join(CSRC, "common.cuh")).read()
    text = text.replace("#include <cuda_runtime.h>", "").replace("#pragma once", "")
    pat = re.compile(r"^__device__ __forceinline__ [^;{]*?\b(\w+)\s*\([^;{]*\{", re.M)
    funcs, i = [], 0      # (start, end, name, body)
    while True:
        m = pat.search(text, i)
        if not m:
            break
        end = _match(text, m.end() - 1, "{", "}")
        funcs.append((m.start(), end, m.group(1), text[m.start():end]))
        i = end
    dropped = {name for _, _, name, body in funcs
               if "asm volatile" in body or "__cvta_generic_to_shared" in body}
    while True:  # ... and whatever calls something that was dropped
        more = {name for _, _, name, body in funcs if name not in dropped
                and any(re.search(r"\b%s\s*\(" % d, body) for d in dropped)}
        if not more:
            break
        dropped |= more
    out, i = [], 0
    for start, end, name, body in funcs:
        out.append(text[i:start])
        if name not in dropped:
            out.append(body)
        i = end
    out.append(text[i:])
    return "".join(out)


def strip_asm(text: str) -> str:
    while True:
        i = text.find("asm volatile")
        if i < 0:
            return text
        j = text.index("(", i)
        end = _match(text, j, "(", ")")
        semi = text.index(";", end)
        text = text[:i] + "/* asm dropped by the emulator */" + text[semi:]


def rewrite_launches(text: str) -> str:
    while True:
        i = text.find("<<<")
        if i < 0:
            return text
        # kernel name (with optional template arguments) ends right before <<<
        k = i
        while k > 0 and text[k - 1].isspace():
            k -= 1
        if text[k - 1] == ">":
            depth, k2 = 0, k - 1
            while True:
                if text[k2] == ">":
                    depth += 1
                elif text[k2] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                k2 -= 1
            k = k2
        start = k
        while start > 0 and (text[start - 1].isalnum() or text[start - 1] == "_"):
            start -= 1
        name = text[start:i].strip()
        j = text.index(">>>", i)
        cfg = _split_top(text[i + 3:j])
        p = text.index("(", j)
        end = _match(text, p, "(", ")")
        args = text[p + 1:end - 1]
        semi = text.index(";", end)
        smem = cfg[2] if len(cfg) > 2 else "0"
        call = "bs_emu_launch((long long)(%s), (long long)(%s), (long long)(%s), [=] { %s(%s); })" % (
            cfg[0], cfg[1], smem, name, args)
        text = text[:start] + call + text[semi:]


def rewrite_dynamic_smem(text: str) -> str:
    """``extern __shared__ T name[];`` -> a pointer into the emulator's dynamic
    shared-memory buffer (sized by the launch's third configuration argument)."""
    return re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s*)?([\w ]+?)\s+(\w+)\[\];",
                  r"\1* \2 = (\1*)bs_dyn_smem;", text)


def source_for(name: str) -> str:
    text = open(os.path.join(CSRC, name + ".cu")).read()
    if '#include "table.cuh"' in text:   # the table header is part of the translation unit
        table = open(os.path.join(CSRC, "table.cuh")).read().replace("#pragma once", "")
        text = text.replace('#include "table.cuh"', table)
    text = text.replace('#include "common.cuh"', '#include "simt.h"\n' + portable_common() + '\n#include "ptx_emu.h"\n')
    text = rewrite_launches(strip_asm(rewrite_dynamic_smem(text)))
    if not re.search(r"^long long bs_launch_counter", text, re.M):
        text += '\nextern "C" { long long bs_launch_counter = 0; }\n'
    return text


def asan_preload() -> str:
    """LD_PRELOAD value for running an ``asan=True`` build inside python ("" without g++)."""
    libs = []
    for lib in ("libasan.so", "libstdc++.so"):
        try:
            p = subprocess.run(["g++", "-print-file-name=" + lib], capture_output=True, text=True).stdout.strip()
        except OSError:
            return ""
        libs.append(os.path.realpath(p))
    return ":".join(libs)


_built = {}


def build(name: str, asan: bool = False) -> str:
    """Compile csrc/<name>.cu for the emulator; returns the path of the shared object.
    Objects are cached by content (source after rewriting + emulator headers + flags), so
    the sanitizer and schedule-fuzzing subprocesses of a test run reuse them."""
    key = (name, asan)
    if key in _built:
        return _built[key]
    src = source_for(name)
    flags = ["-O1", "-g", "-fwrapv", "-std=c++17", "-shared", "-fPIC", "-w", "-I", HERE]
    if asan:
        flags = ["-fsanitize=address", "-fno-omit-frame-pointer"] + flags
    h = hashlib.sha1()
    for part in [src, " ".join(flags)] + [open(os.path.join(HERE, f)).read() for f in ("simt.h", "ptx_emu.h")]:
        h.update(part.encode())
    cache = os.path.join(tempfile.gettempdir(), "bsemu-cache-%d" % os.getuid())
    os.makedirs(cache, exist_ok=True)
    so = os.path.join(cache, "%s-%s.so" % (name, h.hexdigest()[:16]))
    if not os.path.exists(so):
        d = tempfile.mkdtemp(prefix="bsemu-")
        cpp, tmp = os.path.join(d, name + "_emu.cpp"), os.path.join(d, name + "_emu.so")
        with open(cpp, "w") as fh:
            fh.write(src)
        r = subprocess.run(["g++"] + flags + ["-o", tmp, cpp], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("emulator build of %s.cu failed:\n%s" % (name, r.stderr[-3000:]))
        os.replace(tmp, so)   # atomic: concurrent test processes may build the same object
    _built[key] = so
    return so


def emulate(name: str) -> ctypes.CDLL:
    return ctypes.CDLL(build(name))
