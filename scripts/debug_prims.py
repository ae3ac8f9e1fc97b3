"""GPU debug aid: randomised single- and multi-threaded checks of the kernel primitives."""
import sys
import threading
import torch
sys.path.insert(0, ".")
from bigslice_b200.ops import gpu
from bigslice_b200.frame import Frame
from bigslice_b200 import slicetype

errs = []


def check(tid, iters, seed):
    torch.cuda.set_device(0)
    g = torch.Generator(device="cuda").manual_seed(seed)
    for it in range(iters):
        n = int(torch.randint(1, 30000, (1,), device="cuda", generator=g))
        kmax = [500, 7, 100000][it % 3]
        keys = torch.randint(0, kmax, (n,), device="cuda", generator=g)
        vals = torch.randint(1, 100, (n,), device="cuda", generator=g)
        # sort
        sk, sv = gpu.sort_pairs_i64(keys, keys * 7 + 1)
        rk, _ = torch.sort(keys)
        if not torch.equal(sk, rk) or not torch.equal(sv, sk * 7 + 1):
            errs.append(("sort", tid, it, n, kmax))
        # hash combine
        hk, hv = gpu.hash_combine(keys, vals, "add")
        want = torch.zeros(kmax, dtype=torch.int64, device="cuda").index_add_(0, keys, vals)
        got = torch.zeros(kmax, dtype=torch.int64, device="cuda").index_add_(0, hk, hv)
        if not torch.equal(want, got) or len(hk) != int((torch.bincount(keys, minlength=kmax) > 0).sum()):
            errs.append(("combine", tid, it, n, kmax, int((want - got).abs().sum())))
        # partition
        ids = keys % 4
        f = Frame([keys, vals], slicetype.Type([torch.int64, torch.int64], 1))
        parts = gpu.partition_frame(f, ids, 4)
        tot = 0
        for p, fr in enumerate(parts):
            tot += len(fr)
            if len(fr) and not bool((fr.col(0) % 4 == p).all()):
                errs.append(("partition-dest", tid, it, n, p))
        allk = torch.cat([fr.col(0) for fr in parts])
        allv = torch.cat([fr.col(1) for fr in parts])
        got = torch.zeros(kmax, dtype=torch.int64, device="cuda").index_add_(0, allk, allv)
        if tot != n or not torch.equal(want, got):
            errs.append(("partition", tid, it, n, kmax))
        # segmented reduce over the sorted keys
        sk2, perm = gpu.sort_pairs_i64(keys, None)
        svals = vals[perm]
        starts = torch.ones(n, dtype=torch.bool, device="cuda")
        starts[1:] = sk2[1:] != sk2[:-1]
        nseg = int(starts.sum())
        red = gpu.segmented_reduce(svals, starts, nseg, "add")
        uk = sk2[starts]
        got = torch.zeros(kmax, dtype=torch.int64, device="cuda").index_add_(0, uk, red)
        if not torch.equal(want, got):
            errs.append(("segreduce", tid, it, n, kmax))
        # compaction
        m = keys % 3 == 0
        ci = gpu.compact_indices(m)
        if not torch.equal(ci, torch.nonzero(m).flatten()):
            errs.append(("compact", tid, it, n))


check(0, 60, 1)
print("single-thread errors:", len(errs), errs[:10])
errs.clear()
ths = [threading.Thread(target=check, args=(t, 60, 100 + t)) for t in range(4)]
[t.start() for t in ths]
[t.join() for t in ths]
from collections import Counter
print("4-thread errors:", len(errs), Counter(e[0] for e in errs), errs[:10])
