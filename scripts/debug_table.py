"""Diagnose duplicate / missing keys in the two-level table (GPU debug aid)."""
import sys
import torch
sys.path.insert(0, ".")
from bigslice_b200.ops import gpu

g = torch.Generator(device="cuda").manual_seed(11)
nk = 5_000_001
base = torch.randperm(1 << 23, device="cuda", generator=g)[:nk] * 0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF
dup = base[torch.rand(nk, device="cuda", generator=g) < 0.6]
keys = torch.cat([base, dup])
keys = keys[torch.randperm(keys.numel(), device="cuda", generator=g)]
vals = torch.randint(1, 9_000_000, (keys.numel(),), device="cuda", generator=g)
uk, inv = torch.unique(keys, return_inverse=True)
want = torch.zeros(len(uk), dtype=torch.int64, device="cuda").index_add_(0, inv, vals)
for warm, simple in ((None, False), (None, True), (0, False), (0, True), (1 << 16, False)):
    t = gpu.HashTable(2 * len(uk), torch.int64, "add", "cuda", warm=warm)
    t.combine(keys, vals, force_simple=simple)
    torch.cuda.synchronize()
    st = t.status.tolist()
    slots = t.slots
    occ = slots[:, 0] != -(1 << 63)
    tk = slots[:, 0][occ]
    tv = slots[:, 1][occ]
    pos = torch.nonzero(occ).flatten()
    sk, order = torch.sort(tk)
    dupmask = sk[1:] == sk[:-1]
    ndup = int(dupmask.sum())
    missing = len(uk) - len(torch.unique(tk))
    print("warm=%s simple=%s cap=%d warmcap=%d status=%s entries=%d unique=%d dups=%d missing=%d sum_ok=%s"
          % (warm, simple, t.cap, t.warm, st, len(tk), len(uk), ndup, missing,
             int(tv.sum()) == int(vals.sum())))
    if ndup:
        i = torch.nonzero(dupmask).flatten()[:5]
        for a in i.tolist():
            p0, p1 = int(pos[order[a]]), int(pos[order[a + 1]])
            print("   dup key %d at slots %d (%s) and %d (%s) vals %d %d"
                  % (int(sk[a]), p0, "warm" if p0 < t.warm else "main", p1, "warm" if p1 < t.warm else "main",
                     int(tv[order[a]]), int(tv[order[a + 1]])))
    k, v = t.compact()
    print("   compact: %d rows, sum_ok=%s" % (len(k), int(v.sum()) == int(vals.sum())))
