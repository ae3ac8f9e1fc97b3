"""Loader for the host-side C++ runtime (``csrc/native.cpp`` -> ``_bsnative``).

The native module holds the evaluator's scheduling core, CRC32, the host
murmur3 batch hasher and frame-codec helpers.  Pure-python fallbacks exist for
every function so the framework works before ``build()`` has run; the loader
returns ``None`` in that case.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import sysconfig

_lib = None
_tried = False


def lib():
    """The ``_bsnative`` extension module, or None when not built."""
    global _lib, _tried
    if _tried:
        return _lib
    _tried = True
    if os.environ.get("BIGSLICE_NO_NATIVE"):
        return None
    alt = os.environ.get("BIGSLICE_NATIVE_DIR")
    if alt:
        # an alternative build of the module (scripts/sanitize_native.sh: ASan/TSan)
        if "bigslice_b200._bsnative" in sys.modules:   # one pybind11 module per process
            _lib = sys.modules["bigslice_b200._bsnative"]
            return _lib
        path = os.path.join(alt, "_bsnative" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
        spec = importlib.util.spec_from_file_location("bigslice_b200._bsnative", path)
        _lib = importlib.util.module_from_spec(spec)
        sys.modules["bigslice_b200._bsnative"] = _lib
        spec.loader.exec_module(_lib)
        return _lib
    try:
        _lib = importlib.import_module("bigslice_b200._bsnative")
    except Exception:
        _lib = None
        if not os.environ.get("BIGSLICE_NO_AUTOBUILD"):
            try:  # a fresh checkout (the .so is not tracked): build in-tree once
                from . import build
                build.build_native()
                importlib.invalidate_caches()
                _lib = importlib.import_module("bigslice_b200._bsnative")
            except Exception:
                _lib = None
    return _lib


def reset_for_tests():
    global _lib, _tried
    _lib, _tried = None, False
