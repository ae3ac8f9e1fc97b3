"""GPU debug aid: repeat the LocalDevice end-to-end pipeline and report diffs."""
import sys
import torch
sys.path.insert(0, ".")
import bigslice_b200 as bs

n = 50000
g = torch.Generator().manual_seed(5)
keys = torch.randint(0, 500, (n,), generator=g)
vals = torch.ones(n, dtype=torch.int64)


@bs.Func
def pipeline(nshard: int):
    s = bs.Const(nshard, keys, vals)
    s = bs.Filter(s, lambda k, v: k % 2 == 0)
    s = bs.Map(s, lambda k, v: (k // 2, v * 3))
    return bs.Reduce(s, bs.ops.add)


@bs.Func
def plain(nshard: int):
    return bs.Reduce(bs.Const(nshard, keys, vals), bs.ops.add)


want = {}
want2 = {}
for k in keys.tolist():
    want2[k] = want2.get(k, 0) + 1
    if k % 2 == 0:
        want[k // 2] = want.get(k // 2, 0) + 3
bad = 0
for par in (1, 2, 4):
    with bs.start(bs.LocalDevice("cuda"), bs.Parallelism(par)) as sess:
        for it in range(15):
            for fn, w in ((pipeline, want), (plain, want2)):
                got = dict(sess.run(fn, 4).rows())
                if got != w:
                    bad += 1
                    miss = [k for k in w if k not in got]
                    extra = [k for k in got if k not in w]
                    wrong = [(k, got[k], w[k]) for k in w if k in got and got[k] != w[k]]
                    print("par=%d it=%d %s MISMATCH missing=%d extra=%d wrong=%d total got=%d want=%d %s"
                          % (par, it, fn.__name__, len(miss), len(extra), len(wrong),
                             sum(got.values()), sum(w.values()), wrong[:5]))
print("bad runs:", bad)
