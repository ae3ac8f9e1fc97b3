"""Global batch-size tunables.

The reference streams 128-row batches (``internal/defaultsize/size.go:15-18``:
chunk 128, sort canary 256) because every row costs a reflective call.  Here a
batch is a tensor tile handed to a kernel, so the defaults are large; both are
overridable by environment variable / :mod:`bigslice_b200.sliceconfig` flags
(``--bigslice-internal-default-chunk-rows`` etc.).
"""
import os

#: Rows per streamed batch on the host path.
Chunk = int(os.environ.get("BIGSLICE_CHUNK_ROWS", 1 << 16))
#: Rows read to estimate row size before sizing external-sort runs.
SortCanary = int(os.environ.get("BIGSLICE_SORT_CANARY_ROWS", 1 << 12))
#: Rows per batch when a frame lives in HBM (whole-shard tiles).
DeviceChunk = int(os.environ.get("BIGSLICE_DEVICE_CHUNK_ROWS", 1 << 26))


def set_chunk(n: int) -> None:
    global Chunk
    Chunk = int(n)


def set_sort_canary(n: int) -> None:
    global SortCanary
    SortCanary = int(n)
