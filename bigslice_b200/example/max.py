"""Example: per-key integer maximum, and how to test slice code with the kit.

Capability parity: reference ``example/max.go:10-21`` (``IntMax``) — its test
``example/max_test.go`` is mirrored by ``tests/test_examples.py`` using
hypothesis as the property-testing engine.
"""
from __future__ import annotations

import bigslice_b200 as bs


def int_max(slice_: bs.Slice) -> bs.Slice:
    """For a ``<key, int>`` slice return the maximum value per key."""
    return bs.Reduce(slice_, bs.ops.max)


IntMax = int_max
