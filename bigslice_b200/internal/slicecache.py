"""Per-shard file caches for slices.

A cached slice stores shard ``i`` of ``n`` at ``<prefix>-<i:04d>-of-<n:04d>``
as a compressed frame-codec stream.  Reading a cached shard replaces
computing it (and the compiler drops its dependencies); computing an uncached
shard writes the file through as batches are produced and commits it
atomically at EOF.

Capability parity: reference ``internal/slicecache/slicecache.go:14-119``
(ShardCache, Empty, FileShardCache, Cacheable) and ``sliceio.go:16-96``
(write-through and cache readers).  Compression is deflate (zstd is not
available in this image); the format is this framework's own codec.
"""
from __future__ import annotations

import gzip
import itertools
import os
from typing import List, Optional

from ..sliceio import Decoder, Encoder, ErrReader, Reader
from ..slicetype import Type


class ShardCache:
    """Cache interface used by the compiler."""

    def is_cached(self, shard: int) -> bool:
        return False

    def writethrough_reader(self, shard: int, reader: Reader) -> Reader:
        return reader

    def cache_reader(self, shard: int) -> Reader:
        raise RuntimeError("empty shard cache has no shards")

    IsCached, WritethroughReader, CacheReader = is_cached, writethrough_reader, cache_reader


class _Empty(ShardCache):
    pass


Empty = _Empty()


class Cacheable:
    """Mixin for slices discoverable by the compiler as cacheable."""

    def cache(self) -> ShardCache:
        raise NotImplementedError


def shard_path(prefix: str, shard: int, nshard: int) -> str:
    return "%s-%04d-of-%04d" % (prefix, shard, nshard)


_tmp_seq = itertools.count()


class _WritethroughReader(Reader):
    def __init__(self, reader: Reader, path: str):
        self.reader = reader
        self.path = path
        # unique per writer: a task of an abandoned (failed) evaluation may still be
        # writing this shard while its recomputation has already started
        self.tmp = "%s.tmp.%d.%d" % (path, os.getpid(), next(_tmp_seq))
        self.fh = None
        self.enc: Optional[Encoder] = None
        self.done = False

    def _open(self):
        d = os.path.dirname(self.path)
        if d:
            os.makedirs(d, exist_ok=True)
        self.fh = gzip.open(self.tmp, "wb", compresslevel=1)
        self.enc = Encoder(self.fh)

    def read(self, n=None):
        if self.done:
            return None
        try:
            f = self.reader.read(n)
        except BaseException:
            self._discard()
            raise
        if self.fh is None:
            self._open()
        if f is None:
            self.done = True
            self.fh.close()
            self.fh = None
            os.replace(self.tmp, self.path)  # atomic commit
            return None
        if len(f):
            self.enc.write(f)
        return f

    def _discard(self):
        self.done = True
        if self.fh is not None:
            try:
                self.fh.close()
            finally:
                self.fh = None
        try:
            os.unlink(self.tmp)
        except OSError:
            pass

    def close(self):
        if not self.done:
            self._discard()  # incomplete read leaves no valid cache
        self.reader.close()


class FileShardCache(ShardCache):
    """Shard cache backed by files under a path prefix."""

    def __init__(self, prefix: str, nshard: int, typ: Type, require_all: bool = False):
        self.prefix = prefix
        self.nshard = nshard
        self.typ = typ
        self.require_all = require_all
        self.shard_is_cached: List[bool] = [
            os.path.exists(shard_path(prefix, i, nshard)) for i in range(nshard)]

    def require_all_cached(self) -> None:
        """All-or-nothing: if any shard is missing, treat none as cached
        (``Cache`` semantics; ``CachePartial`` keeps per-shard state)."""
        if not all(self.shard_is_cached):
            self.shard_is_cached = [False] * self.nshard

    def is_cached(self, shard: int) -> bool:
        return self.shard_is_cached[shard]

    def writethrough_reader(self, shard: int, reader: Reader) -> Reader:
        if self.shard_is_cached[shard]:
            return reader
        return _WritethroughReader(reader, shard_path(self.prefix, shard, self.nshard))

    def cache_reader(self, shard: int) -> Reader:
        path = shard_path(self.prefix, shard, self.nshard)
        try:
            fh = gzip.open(path, "rb")
        except OSError as e:
            return ErrReader(e)
        from .. import context as _context
        return Decoder(fh, self.typ, _context.current_device(), closer=fh.close)
