import os, sys, torch
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29611")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
sys.path.insert(0, os.getcwd())
import torch.distributed as dist
torch.cuda.set_device(0)
dist.init_process_group("gloo")
from bigslice_b200.parallel.peer import PeerGroup, ExchangeVoteError
from bigslice_b200.frame import hashing
pg = PeerGroup(dist.group.WORLD, 64 << 20, "cuda:0")
k = torch.randint(-1000, 1000, (200000,), device="cuda")
v = torch.randn(200000, device="cuda")
kinds = []
for vote in ((1, 5), (1, 5), (0, 9)):
    try:
        desc = [(k.data_ptr(), 8, 8, torch.int64, ()), (v.data_ptr(), 4, 4, torch.float32, ())]
        got = pg.exchange_raw(desc, k.numel(), 3, vote=vote)
        kinds.append(sum(int(c[0].numel()) for c in got.values()))
    except ExchangeVoteError as e:
        kinds.append(-e.kind)
got = pg.exchange([k, v], 3)
dest = hashing.hash_tensor(k.cpu(), 0) % 3
ok = all(sorted(got[p][0].cpu().tolist()) == sorted(k.cpu()[dest == p].tolist()) for p in range(3))
print("QUICK", kinds, ok)
