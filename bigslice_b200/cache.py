"""Cache / CachePartial / ReadCache: checkpoint a slice's shards to files.

* ``Cache(s, prefix)`` — if *all* shard files exist, read them instead of
  computing ``s`` (and everything it depends on); otherwise compute and write
  every shard through.
* ``CachePartial(s, prefix)`` — per-shard: existing shards are read, missing
  ones computed and written (resumes an interrupted run).
* ``ReadCache(typ, nshard, prefix)`` — read a cache written earlier; all
  shards must exist.

Capability parity: reference ``cache.go:16-95``; semantics pinned by
``cache_test.go:116-351``.
"""
from __future__ import annotations

from typing import List

from . import slicetype, typecheck
from .internal.slicecache import Cacheable, FileShardCache, ShardCache
from .sliceio import Reader
from .slices import Dep, Slice, _check_slice, make_name
from .slicetype import Type


class _CacheSlice(Slice, Cacheable):
    def __init__(self, s: Slice, op: str, cache: ShardCache):
        Slice.__init__(self, make_name(op), s.typ, s.num_shard(), [Dep(s)], ())
        self._cache = cache

    def cache(self) -> ShardCache:
        return self._cache

    def reader(self, shard, deps: List[Reader]) -> Reader:
        return deps[0]


def Cache(slice_: Slice, prefix: str) -> Slice:
    """All-or-nothing file cache of ``slice_`` under ``prefix``."""
    _check_slice("cache", slice_)
    c = FileShardCache(prefix, slice_.num_shard(), slice_.typ)
    c.require_all_cached()
    return _CacheSlice(slice_, "cache", c)


def CachePartial(slice_: Slice, prefix: str) -> Slice:
    """Per-shard file cache of ``slice_`` under ``prefix``."""
    _check_slice("cachepartial", slice_)
    c = FileShardCache(prefix, slice_.num_shard(), slice_.typ)
    return _CacheSlice(slice_, "cachepartial", c)


class _ReadCache(Slice, Cacheable):
    def __init__(self, typ: Type, nshard: int, cache: FileShardCache):
        Slice.__init__(self, make_name("readcache"), typ, nshard)
        self._cache = cache

    def cache(self) -> ShardCache:
        return self._cache

    def reader(self, shard, deps):
        # Only reached when a shard file is missing: surface a run error.
        return self._cache.cache_reader(shard)


def ReadCache(typ, nshard: int, prefix: str) -> Slice:
    """A slice that reads a previously written cache."""
    typ = slicetype.as_type(typ)
    if nshard < 1:
        typecheck.panic("readcache: shard must be >= 1")
    c = FileShardCache(prefix, nshard, typ)
    return _ReadCache(typ, nshard, c)
