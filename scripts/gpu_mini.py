"""Smallest possible GPU check of the SPMD failure path and the device executor."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bigslice_b200 as bs  # noqa: E402
import spmd_worker as w  # noqa: E402  (registers its Funcs)


@bs.Func
def dev_reduce(nshard: int):
    g = torch.Generator().manual_seed(5)
    keys = torch.randint(0, 500, (200000,), generator=g)
    return bs.Reduce(bs.Const(nshard, keys, torch.ones(200000, dtype=torch.int64)), bs.ops.add)


with bs.start(bs.LocalDevice("cuda"), bs.Parallelism(2)) as sess:
    g = torch.Generator().manual_seed(5)
    keys = torch.randint(0, 500, (200000,), generator=g)
    want = torch.bincount(keys, minlength=500)
    got = dict(sess.run(dev_reduce, 6).rows())
    assert got == {k: int(want[k]) for k in range(500)}, "LocalDevice reduce mismatch"
    print("local device reduce ok")
sess = bs.start(bs.Spmd(device="cuda", arena_bytes=64 << 20), bs.Parallelism(2))
ex = sess.executor
print("failure_ok", w.failure_check(sess, ex, ex.world, ex.rank))
sess.shutdown()
