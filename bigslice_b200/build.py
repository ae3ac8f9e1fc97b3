"""In-tree builds of the native components.

* ``_bscuda.so``  — the sm_100a kernels in ``csrc/*.cu`` (plain C ABI, loaded
  with ctypes by :mod:`bigslice_b200.ops.gpu`), compiled with
  ``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo``.
* ``_bsnative``   — the host C++ runtime in ``csrc/native.cpp`` (pybind11):
  evaluator core, CRC32, host murmur3, codec helpers.

Both land inside the package directory so they travel with the source tree.
``python -m bigslice_b200.build`` builds both.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(ROOT, "csrc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
              "--expt-relaxed-constexpr", "-std=c++17", "-Xcompiler", "-fPIC"]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


class _BuildLock:
    """Serialises in-tree builds across processes (several ranks may import a
    fresh checkout at once); the winner builds, the others find the result."""

    def __init__(self, name: str):
        self.path = os.path.join(PKG, ".build-%s.lock" % name)
        self.fh = None

    def __enter__(self):
        import fcntl
        self.fh = open(self.path, "w")
        fcntl.flock(self.fh, fcntl.LOCK_EX)
        return self

    def __exit__(self, *a):
        import fcntl
        fcntl.flock(self.fh, fcntl.LOCK_UN)
        self.fh.close()
        return False


def cuda_lib_path() -> str:
    return os.path.join(PKG, "_bscuda.so")


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu into bigslice_b200/_bscuda.so for sm_100a."""
    out = cuda_lib_path()
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh"))
    if not force and not _newer(out, deps):
        return out
    with _BuildLock("cuda"):
        if not force and not _newer(out, deps):
            return out  # another process built it while we waited
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        tmp = out + ".tmp.%d" % os.getpid()
        cmd = [nvcc] + NVCC_FLAGS + ["-shared", "-o", tmp] + srcs
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
        os.replace(tmp, out)  # readers never see a half-written library
        if verbose:
            sys.stderr.write(r.stderr)
    return out


def native_lib_path() -> str:
    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return os.path.join(PKG, "_bsnative" + ext)


def build_native(force: bool = False) -> str:
    """Compile csrc/native.cpp into the pybind11 module bigslice_b200._bsnative."""
    import pybind11
    out = native_lib_path()
    src = os.path.join(CSRC, "native.cpp")
    if not force and not _newer(out, [src]):
        return out
    with _BuildLock("native"):
        if not force and not _newer(out, [src]):
            return out
        cxx = shutil.which("g++") or "g++"
        inc = [pybind11.get_include(), sysconfig.get_paths()["include"]]
        tmp = out + ".tmp.%d" % os.getpid()
        cmd = [cxx, "-O3", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden"] + \
              ["-I" + i for i in inc] + [src, "-o", tmp]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
        os.replace(tmp, out)
    return out


def build_all(force: bool = False, verbose: bool = False):
    return build_cuda(force, verbose), build_native(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
