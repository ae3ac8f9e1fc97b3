import os, sys, torch
sys.path.insert(0, '/root/repo')
from bigslice_b200.ops import gpu
from bigslice_b200.models import wordcount as wc
n = 1 << 27
keys = wc.zipf_keys(n, 1 << 24, 1.0, seed=1234, device='cuda')
vals = torch.ones(n, dtype=torch.int64, device='cuda')
t = gpu.HashTable(gpu.table_capacity(int(11e6 * 1.15)), torch.int64, 'add', 'cuda')
def run():
    t.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t.combine(keys, vals); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for _ in range(3): run()
print(os.environ.get('BIGSLICE_HC_DEBUG', '0'), 'warm', t.warm, min(run() for _ in range(5)))
