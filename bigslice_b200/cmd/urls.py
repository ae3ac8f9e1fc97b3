"""urls: count URL domains and report the most frequent ones.

    python -m bigslice_b200.cmd.urls [session flags] [--input FILE ...] [--synthetic N] [--top K]

The reference demo reads GDELT CSVs from S3, extracts the domain of every URL
and counts domains with Reduce (``cmd/urls/urls.go:37-126``).  There is no
network here, so inputs are local files (one URL per line / any column
containing a URL) or a synthetic generator with Zipf-distributed domains.

Pipeline: ``ScanReader/ReaderGen -> Tokenize -> UrlDomains -> StringIds ->
Reduce(add) -> per-shard top-K``.  On a GPU executor the lines are moved to HBM
as a CSR string column once; tokenising, host extraction (lower-cased) and the
64-bit dictionary ids are string kernels (``csrc/strings.cu``), the count runs on
the int64 hash-aggregation + fused NVLink exchange path, and the id -> domain
dictionary stays on the device until the top-K rows are printed.
"""
from __future__ import annotations

import argparse
import sys
from typing import Dict
from urllib.parse import urlparse

import numpy as np
import torch

import bigslice_b200 as bs
from bigslice_b200 import sliceconfig
from bigslice_b200.frame import hashing

from bigslice_b200 import text as _text  # noqa: E402

#: id -> domain, learned while the pipeline runs (device resident on GPU executors)
DOMAINS = _text.IdDictionary()


def domain_of(url: str) -> str:
    if "://" not in url:
        url = "http://" + url
    try:
        return urlparse(url).netloc.lower()
    except ValueError:
        return ""


def synthetic_urls(n: int, ndomain: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, ndomain + 1)
    cdf = np.cumsum(w) / w.sum()
    r = np.searchsorted(cdf, rng.random(n))
    out = np.empty(n, dtype=object)
    out[:] = ["https://site%d.example.com/page/%d" % (d, i) for i, d in enumerate(r.tolist())]
    return out


@bs.Func
def count_domains(nshard: int, files: list, synthetic: int, ndomain: int, top: int):
    if files:
        def opener():
            import io
            return io.StringIO("".join(open(f).read() for f in files))
        lines = bs.ScanReader(nshard, opener)
    else:
        def gen(shard):
            per = synthetic // nshard
            yield (synthetic_urls(per, ndomain, shard),)
        lines = bs.ReaderGen(nshard, gen, [bs.string])

    # URL column -> host -> 64-bit dictionary id; on a GPU executor all three steps
    # are string kernels over a CSR column in HBM (bigslice_b200/text.py), and the
    # count then runs on the int64 hash-aggregation + fused exchange path.
    from bigslice_b200 import text
    tokens = text.Tokenize(text.ToDevice(lines), b" ,\t\r\n")
    domains = text.UrlDomains(tokens)
    counts = bs.Reduce(text.StringIds(domains, DOMAINS), bs.ops.add)

    def topk(k, v):
        kk = min(top, v.numel())
        idx = torch.topk(v, kk).indices
        return k[idx], v[idx]
    return bs.Flatmap(counts, topk, out=[bs.int64, bs.int64])


def main(argv=None) -> int:
    sess, rest = sliceconfig.parse(argv)
    ap = argparse.ArgumentParser(prog="urls")
    ap.add_argument("--input", action="append", default=[])
    ap.add_argument("--synthetic", type=int, default=100000)
    ap.add_argument("--ndomain", type=int, default=5000)
    ap.add_argument("--nshard", type=int, default=8)
    ap.add_argument("--top", type=int, default=10)
    a = ap.parse_args(rest)
    try:
        res = sess.run(count_domains, a.nshard, a.input, a.synthetic, a.ndomain, a.top)
        g = getattr(sess.executor, "gather", None)
        rows = g(res) if g is not None else res.rows()
    finally:
        sess.shutdown()
    rows.sort(key=lambda r: -r[1])
    rows = rows[:a.top]
    names = DOMAINS.lookup_global([k for k, _ in rows])
    for k, v in rows:
        print("%s\t%d" % (names.get(k) or "<%d>" % k, v))
    return 0


if __name__ == "__main__":
    sys.exit(main())
