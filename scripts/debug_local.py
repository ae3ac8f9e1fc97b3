"""Stage-by-stage GPU-vs-CPU comparison of the LocalDevice pipeline (debug aid)."""
import torch
import bigslice_b200 as bs
from bigslice_b200 import slicefunc, sliceio, sortio
from bigslice_b200.frame import Frame
from bigslice_b200.ops import gpu
from bigslice_b200.exec.combiner import Combiner, combine_frame
from bigslice_b200.exec.local import split_by_partition

g = torch.Generator().manual_seed(5)
n = 50000
keys = torch.randint(0, 500, (n,), generator=g)
vals = torch.ones(n, dtype=torch.int64)
func = slicefunc.of(bs.ops.add)


def stage(dev):
    out = {}
    f = Frame.of(keys.to(dev), vals.to(dev))
    m = f.col(0) % 2 == 0
    if dev == "cuda":
        idx = gpu.compact_indices(m)
        out["filter_idx"] = idx.cpu()
        f = f.take(idx)
    else:
        out["filter_idx"] = torch.nonzero(m).flatten()
        f = f.mask(m)
    f = Frame([f.col(0) // 2, f.col(1) * 3], f.typ)
    out["mapped"] = sorted(f.to("cpu").rows())
    ids = f.hash(0) % 4
    out["ids"] = ids.cpu()
    parts = split_by_partition(f, ids, 4)
    out["parts"] = [sorted(p.to("cpu").rows()) for p in parts]
    comb = []
    for p in parts:
        c = combine_frame(p, func)
        comb.append(sorted(c.to("cpu").rows()))
    out["combined"] = comb
    c = Combiner(f.typ, func)
    for p in parts:
        c.combine(p)
    out["combiner"] = sliceio.read_all(c.reader(), f.typ).to("cpu").rows()
    srt = [combine_frame(p, func).sorted() for p in parts]
    rr = sortio.reduce_reader(f.typ, [sliceio.FrameReader(s) for s in srt], func)
    out["reduce_reader"] = sliceio.read_all(rr, f.typ).to("cpu").rows()
    return out


a, b = stage("cpu"), stage("cuda")
for k in a:
    same = (torch.equal(a[k], b[k]) if isinstance(a[k], torch.Tensor) else a[k] == b[k])
    print(k, "OK" if same else "MISMATCH")
    if not same and not isinstance(a[k], torch.Tensor):
        if isinstance(a[k][0], list):
            for i, (x, y) in enumerate(zip(a[k], b[k])):
                if x != y:
                    print("  part", i, "cpu", len(x), "gpu", len(y), [r for r in y if r not in x][:5], [r for r in x if r not in y][:5])
        else:
            print("  cpu", len(a[k]), "gpu", len(b[k]), [r for r in b[k] if r not in a[k]][:5])
