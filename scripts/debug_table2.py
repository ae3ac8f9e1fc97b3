"""GPU debug aid: run the TMA combine of an alternative build (BIGSLICE_CUDA_LIB) and
report which input rows went missing."""
import os
import sys
import torch
sys.path.insert(0, ".")
from bigslice_b200.ops import gpu

mode = os.environ.get("DBG", "")
g = torch.Generator(device="cuda").manual_seed(11)
nk = 5_000_001
base = torch.randperm(1 << 23, device="cuda", generator=g)[:nk] * 0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF
dup = base[torch.rand(nk, device="cuda", generator=g) < 0.6]
keys = torch.cat([base, dup])
keys = keys[torch.randperm(keys.numel(), device="cuda", generator=g)]
vals = torch.randint(1, 9_000_000, (keys.numel(),), device="cuda", generator=g)
uk, inv = torch.unique(keys, return_inverse=True)
want = torch.zeros(len(uk), dtype=torch.int64, device="cuda").index_add_(0, inv, vals)
n = keys.numel()
TILE = 3968
for rep in range(3):
    t = gpu.HashTable(2 * len(uk), torch.int64, "add", "cuda", warm=0)
    t.combine(keys, vals)
    torch.cuda.synchronize()
    st = t.status.tolist()
    if mode == "SUMKEYS":
        ks = int(keys.sum().item()) & ((1 << 64) - 1)
        print("mode=%s rep=%d n=%d rows_seen=%d keysum_ok=%s" % (mode, rep, n, st[3], (st[2] & ((1 << 64) - 1)) == ks))
        continue
    slots = t.slots
    occ = slots[:, 0] != -(1 << 63)
    tk, tv = slots[:, 0][occ], slots[:, 1][occ]
    stk, order = torch.sort(tk)
    stv = tv[order]
    idx = torch.searchsorted(stk, uk).clamp(max=len(stk) - 1)
    found = stk[idx] == uk
    got = torch.where(found, stv[idx], torch.zeros_like(want))
    missing_keys = uk[~found]
    wrong = found & (got != want)
    print("mode=%s rep=%d status=%s entries=%d unique=%d missing=%d wrong_sums=%d total_delta=%d"
          % (mode, rep, st, len(tk), len(uk), len(missing_keys), int(wrong.sum()),
             int(tv.sum()) - int(vals.sum())))
    if len(missing_keys):
        mpos = torch.nonzero(torch.isin(keys, missing_keys)).flatten()
        tile = (mpos // TILE).tolist()
        within = (mpos % TILE)
        print("   missing rows=%d tiles(first 20)=%s" % (len(mpos), tile[:20]))
        print("   distinct tiles=%d  cta(first20)=%s  it(first20)=%s" % (
            len(set(tile)), [x % 148 for x in tile[:20]], [x // 148 for x in tile[:20]]))
        print("   thread(first20)=%s j=%s" % ((within % 992).tolist()[:20], (within // 992).tolist()[:20]))
    if int(wrong.sum()):
        w = torch.nonzero(wrong).flatten()[:10]
        print("   wrong: got-want=%s" % ((got[w] - want[w]).tolist(),))
