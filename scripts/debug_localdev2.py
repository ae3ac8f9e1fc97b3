"""GPU debug aid: run the LocalDevice Reduce with invariant checks wrapped around the
building blocks, to find the first one that produces a wrong frame."""
import sys
import torch
sys.path.insert(0, ".")
import bigslice_b200 as bs
from bigslice_b200.frame import frame as F
from bigslice_b200.exec import combiner as CB, local as L
from bigslice_b200.sortio import reader as SR, sort as SS
from bigslice_b200.ops import gpu

flags = []


def ksum(f):
    if len(f) == 0:
        return {}
    k, v = f.col(0).cpu(), f.col(f.num_out() - 1).cpu()
    out = {}
    for a, b in zip(k.tolist(), v.tolist()):
        out[a] = out.get(a, 0) + b
    return out


orig_sorted = F.Frame.sorted
def sorted_chk(self):
    out = orig_sorted(self)
    if len(self) > 1 and isinstance(self.col(0), torch.Tensor):
        k = out.col(0)
        if not bool((k[1:] >= k[:-1]).all()) or ksum(out) != ksum(self):
            flags.append(("sorted", len(self), str(self.col(0).device), self.col(0).is_contiguous(),
                          self.col(0).data_ptr() % 16))
    return out
F.Frame.sorted = sorted_chk

orig_cf = CB.combine_frame
def cf_chk(f, func, ctx=None):
    out = orig_cf(f, func, ctx)
    if ksum(out) != ksum(f) or len(set(out.col(0).tolist())) != len(out):
        flags.append(("combine_frame", len(f), len(out), f.col(0).data_ptr() % 16, f.col(1).data_ptr() % 16))
    return out
CB.combine_frame = cf_chk

orig_cs = SR.combine_sorted
def cs_chk(f, func, ctx=None):
    out = orig_cs(f, func, ctx)
    if ksum(out) != ksum(f):
        flags.append(("combine_sorted", len(f), len(out)))
    return out
SR.combine_sorted = cs_chk

orig_pf = gpu.partition_frame
def pf_chk(f, ids, npart):
    out = orig_pf(f, ids, npart)
    tot = {}
    for p in out:
        for k, v in ksum(p).items():
            tot[k] = tot.get(k, 0) + v
    if tot != ksum(f) or sum(len(p) for p in out) != len(f):
        flags.append(("partition_frame", len(f), [len(p) for p in out]))
    return out
gpu.partition_frame = pf_chk

orig_step = SS.MergeReader._step
def step_chk(self):
    before = {}
    live_in = []
    for b in self.bufs:
        if b.frame is not None:
            live_in.append(b.frame)
    out = orig_step(self)
    if out is not None and len(out) > 1:
        k = out.col(0)
        if not bool((k[1:] >= k[:-1]).all()):
            flags.append(("merge_step_unsorted", len(out)))
    return out
SS.MergeReader._step = step_chk

n = 50000
g = torch.Generator().manual_seed(5)
keys = torch.randint(0, 500, (n,), generator=g)
vals = torch.ones(n, dtype=torch.int64)


@bs.Func
def plain(nshard: int):
    return bs.Reduce(bs.Const(nshard, keys, vals), bs.ops.add)


want2 = {}
for k in keys.tolist():
    want2[k] = want2.get(k, 0) + 1
with bs.start(bs.LocalDevice("cuda"), bs.Parallelism(1)) as sess:
    for it in range(10):
        flags.clear()
        res = sess.run(plain, 4)
        got = {}
        per_shard = []
        for f in res.frames():
            per_shard.append(len(f))
            for k, v in ksum(f).items():
                got[k] = got.get(k, 0) + v
        wrong = [(k, got.get(k), w) for k, w in want2.items() if got.get(k) != w]
        print("it=%d wrong=%s flags=%s shards=%s" % (it, wrong[:6], flags[:8], per_shard))
