"""Task-output stores: where executors keep partitioned task output.

* :class:`MemoryStore` — frames kept in process (HBM or host) — tests/default.
* :class:`FileStore`   — one file per (task, partition) under
  ``{prefix}/{hash byte}/{op}/{shard}-of-{n}/p{partition}``: a frame-codec
  stream followed by an 8-byte little-endian record-count trailer; writes go to
  a temp name and are committed atomically; ``open(offset=rows)`` resumes a read
  part-way (for retrying readers).  Lets an executor park outputs on NVMe when
  HBM/host memory is short.

Capability parity: reference ``exec/store.go:23-285`` (Store, writeCommitter,
memoryStore, fileStore, sliceInfo{Size,Records}).
"""
from __future__ import annotations

import os
import shutil
import struct
import threading
import zlib
from typing import Dict, List, Optional, Tuple

from .. import errors
from ..frame import Frame
from ..sliceio import Decoder, Encoder, FrameReader, MultiReader, Reader
from ..slicetype import Type
from .task import TaskName


class SliceInfo:
    """Size in bytes and number of records of a stored partition."""

    def __init__(self, size: int, records: int):
        self.size, self.records = size, records


class WriteCommitter:
    """Accepts frames; ``commit`` publishes them, ``discard`` drops them."""

    def write(self, f: Frame) -> None:
        raise NotImplementedError

    def commit(self, records: Optional[int] = None) -> None:
        raise NotImplementedError

    def discard(self) -> None:
        raise NotImplementedError


class Store:
    def create(self, task: TaskName, partition: int) -> WriteCommitter:
        raise NotImplementedError

    def open(self, task: TaskName, partition: int, typ: Type, offset: int = 0, device=None) -> Reader:
        raise NotImplementedError

    def stat(self, task: TaskName, partition: int) -> SliceInfo:
        raise NotImplementedError

    def discard(self, task: TaskName, partition: int) -> None:
        raise NotImplementedError


class _SkipReader(Reader):
    """Skips the first ``offset`` rows of a stream."""

    def __init__(self, r: Reader, offset: int):
        self.r, self.left = r, offset

    def read(self, n=None):
        while True:
            f = self.r.read(n)
            if f is None:
                return None
            if self.left <= 0:
                return f
            if len(f) <= self.left:
                self.left -= len(f)
                continue
            f = f.slice(self.left, len(f))
            self.left = 0
            return f

    def close(self):
        self.r.close()


class MemoryStore(Store):
    def __init__(self):
        self.lock = threading.Lock()
        self.data: Dict[Tuple[str, int], List[Frame]] = {}

    def create(self, task, partition):
        store, key = self, (str(task), partition)

        class W(WriteCommitter):
            def __init__(self):
                self.frames: List[Frame] = []

            def write(self, f):
                if len(f):
                    self.frames.append(f)

            def commit(self, records=None):
                with store.lock:
                    store.data[key] = self.frames

            def discard(self):
                self.frames = []
        return W()

    def open(self, task, partition, typ, offset=0, device=None):
        with self.lock:
            frames = self.data.get((str(task), partition))
        if frames is None:
            raise errors.Error("store: %s[%d] does not exist" % (task, partition), errors.Kind.NotExist)
        r: Reader = MultiReader([FrameReader(f) for f in frames])
        return _SkipReader(r, offset) if offset else r

    def stat(self, task, partition):
        with self.lock:
            frames = self.data.get((str(task), partition))
        if frames is None:
            raise errors.Error("store: %s[%d] does not exist" % (task, partition), errors.Kind.NotExist)
        return SliceInfo(sum(f.nbytes() for f in frames), sum(len(f) for f in frames))

    def discard(self, task, partition):
        with self.lock:
            self.data.pop((str(task), partition), None)


class FileStore(Store):
    def __init__(self, prefix: str):
        self.prefix = prefix
        os.makedirs(prefix, exist_ok=True)

    def path(self, task: TaskName, partition: int) -> str:
        h = zlib.crc32(str(task).encode()) & 0xFF
        return os.path.join(self.prefix, "%02x" % h, task.op,
                            "%03d-of-%03d" % (task.shard, task.num_shard), "p%03d" % partition)

    def create(self, task, partition):
        path = self.path(task, partition)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = "%s.tmp.%d.%d" % (path, os.getpid(), threading.get_ident())

        class W(WriteCommitter):
            def __init__(self):
                self.fh = open(tmp, "wb")
                self.enc = Encoder(self.fh)
                self.records = 0

            def write(self, f):
                if len(f):
                    self.enc.write(f)
                    self.records += len(f)

            def commit(self, records=None):
                self.fh.write(struct.pack("<q", self.records if records is None else records))
                self.fh.close()
                os.replace(tmp, path)

            def discard(self):
                self.fh.close()
                try:
                    os.unlink(tmp)
                except OSError:
                    pass
        return W()

    def _trailer(self, path: str) -> Tuple[int, int]:
        try:
            size = os.path.getsize(path)
            with open(path, "rb") as fh:
                fh.seek(size - 8)
                (records,) = struct.unpack("<q", fh.read(8))
        except (OSError, struct.error):
            raise errors.Error("store: %s does not exist" % path, errors.Kind.NotExist)
        return size - 8, records

    def open(self, task, partition, typ, offset=0, device=None):
        path = self.path(task, partition)
        size, _ = self._trailer(path)
        fh = open(path, "rb")

        class Limited:
            """File view that ends before the record-count trailer."""

            def __init__(self):
                self.left = size

            def read(self, n=-1):
                if n is None or n < 0 or n > self.left:
                    n = self.left
                b = fh.read(n)
                self.left -= len(b)
                return b
        r: Reader = Decoder(Limited(), typ, device, closer=fh.close)
        return _SkipReader(r, offset) if offset else r

    def stat(self, task, partition):
        size, records = self._trailer(self.path(task, partition))
        return SliceInfo(size, records)

    def discard(self, task, partition):
        try:
            os.unlink(self.path(task, partition))
        except OSError:
            pass

    def cleanup(self):
        shutil.rmtree(self.prefix, ignore_errors=True)
