"""Loader for the host-side C++ runtime (``csrc/native.cpp`` -> ``_bsnative``).

The native module holds the evaluator's scheduling core, CRC32, the host
murmur3 batch hasher and frame-codec helpers.  Pure-python fallbacks exist for
every function so the framework works before ``build()`` has run; the loader
returns ``None`` in that case.
"""
from __future__ import annotations

import importlib
import os
import sys

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
