"""Small instances of every sm_100a kernel, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python scripts/sanitizer_smoke.py
    compute-sanitizer --tool racecheck python scripts/sanitizer_smoke.py   (shared-memory hazards)
    compute-sanitizer --tool synccheck python scripts/sanitizer_smoke.py

The GPU analogue of the reference's `go test -race` tier.  Sizes are
tiny (the tools slow kernels down 10-100x) but chosen to take every code path:
the TMA ring (>= 4 tiles), partial tiles, both table-exchange kernels, flushes.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bigslice_b200 as bs  # noqa: E402
from bigslice_b200.frame import Frame  # noqa: E402
from bigslice_b200 import slicetype  # noqa: E402
from bigslice_b200.ops import gpu  # noqa: E402

only = set(sys.argv[1:])


def want(name):
    return not only or name in only


g = torch.Generator(device="cuda").manual_seed(1)
n = 20011
keys = torch.randint(0, 3000, (n,), device="cuda", generator=g)
vals = torch.randint(1, 9, (n,), device="cuda", generator=g)
ref = torch.zeros(3000, dtype=torch.int64, device="cuda").index_add_(0, keys, vals)

if want("hash"):
    for dt in (torch.int64, torch.int32, torch.float32, torch.int8):
        gpu.hash_column(keys.to(dt))
if want("partition"):
    f = Frame([keys, vals], slicetype.Type([torch.int64, torch.int64], 1))
    parts = gpu.partition_frame(f, keys % 5, 5)
    assert sum(len(p) for p in parts) == n
if want("combine"):
    for warm in (None, 1 << 10, 0):
        t = gpu.HashTable(8000, torch.int64, "add", "cuda", warm=warm)
        t.track(3)
        t.combine(keys, vals)                      # TMA ring: 6 tiles incl. a partial one
        t.combine(keys[1:301], vals[1:301])        # unaligned + tiny: simple kernel
        k, v = t.compact(clear=True)
        got = torch.zeros(3000, dtype=torch.int64, device="cuda").index_add_(0, k, v)
        exp = ref.clone().index_add_(0, keys[1:301], vals[1:301])
        assert torch.equal(got, exp), "combine mismatch"
    for dt, op in ((torch.float32, "max"), (torch.int32, "min"), (torch.float64, "add")):
        gpu.hash_combine(keys, vals.to(dt), op)
if want("sort"):
    sk, sv = gpu.sort_pairs_i64(keys * 1000003 - 5, None)
    assert bool((sk[1:] >= sk[:-1]).all())
if want("misc"):
    m = keys % 3 == 0
    assert torch.equal(gpu.compact_indices(m), torch.nonzero(m).flatten())
    sk, perm = gpu.sort_pairs_i64(keys, None)
    st = torch.ones(n, dtype=torch.bool, device="cuda")
    st[1:] = sk[1:] != sk[:-1]
    gpu.segmented_reduce(vals[perm], st, int(st.sum()), "add")
    offs = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), keys[:500] % 4]), 0)
    gpu.expand_indices(offs)
    gpu.crc32(keys)
    gpu.topk(vals * 1000 + keys, 17)
if want("jit"):
    @bs.Func
    def pipe(nshard: int):
        s = bs.Const(nshard, keys.cpu(), vals.cpu())
        s = bs.FilterExpr(s, "c0 % 2 == 0")
        return bs.MapExpr(s, ["c0 / 2", "c1 * 3"], out=[torch.int64, torch.int64])
if want("exchange") or want("jit"):
    with bs.start(bs.Spmd(device="cuda", arena_bytes=64 << 20), bs.Parallelism(2)) as sess:
        ex = sess.executor
        if want("jit"):
            rows = list(sess.run(pipe, 2).rows())
            assert len(rows) == int((keys % 2 == 0).sum())
        if want("exchange"):
            for nparts in (3, 12):
                got = ex.peers.exchange([keys, vals.to(torch.float32), vals.to(torch.int32)], nparts, nkeys=1)
                assert sum(int(c[0].numel()) for c in got.values()) == n
                for tracked in (False, True):
                    t = gpu.HashTable(8000, torch.int64, "add", "cuda", warm=0)
                    if tracked:
                        t.track(nparts)
                    t.combine(keys, vals)
                    got = ex.peers.exchange_table(t, nparts, clear=True)
                    tot = torch.zeros(3000, dtype=torch.int64, device="cuda")
                    for d, (rk, rv) in got.items():
                        tot.index_add_(0, rk, rv)
                    assert torch.equal(tot, ref), "table exchange mismatch"
torch.cuda.synchronize()
print("sanitizer smoke: ok")
