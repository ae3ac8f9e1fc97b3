"""Worker program for multi-process SPMD tests (launched with torch.distributed.run)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bigslice_b200 as bs  # noqa: E402
from bigslice_b200.models import wordcount as _wc  # noqa: E402,F401  (registers its Funcs before the session)

N = 4000
rng = np.random.default_rng(7)
KEYS = torch.from_numpy(rng.integers(0, 97, N))
WORDS = np.array(["w%d" % (i % 13) for i in range(N)], dtype=object)
scanned = []


@bs.Func
def const_map_scan(nshard: int):
    s = bs.Const(nshard, torch.arange(100))
    s = bs.Map(s, lambda x: x * 2)

    def sink(shard, scanner):
        for (v,) in scanner:
            scanned.append(v)
    return bs.Scan(s, sink)


@bs.Func
def wordcount(nshard: int):
    s = bs.Const(nshard, KEYS, torch.ones(N, dtype=torch.int64))
    return bs.Reduce(s, bs.ops.add)


@bs.Func
def strcount(nshard: int):
    s = bs.Const(nshard, WORDS, torch.ones(N, dtype=torch.int64))
    return bs.Reduce(s, lambda a, b: a + b)


from bigslice_b200 import text as _text  # noqa: E402
WORD_DICT = _text.IdDictionary()


@bs.Func
def textcount(nshard: int):
    """String keys on the device path: lines -> tokens -> 64-bit ids -> Reduce(add)."""
    lines = np.array([" ".join(WORDS[i:i + 5].tolist()) for i in range(0, N, 5)], dtype=object)
    s = _text.ToDevice(bs.Const(nshard, lines))
    return bs.Reduce(_text.StringIds(_text.Tokenize(s), WORD_DICT), bs.ops.add)


@bs.Func
def join(nshard: int):
    a = bs.Const(nshard, KEYS, torch.arange(N, dtype=torch.float32))
    b = bs.Const(nshard, torch.arange(0, 97, 2), torch.arange(0, 97, 2) * 10)
    return bs.Cogroup(a, b)


PA = (torch.arange(N) % 11).to(torch.int32) - 5
PB = ((torch.arange(N) % 7).to(torch.float32) - 3) / 2
P8 = ((torch.arange(N) * 37) % 200 - 100).to(torch.int8)


@bs.Func
def pairsum(nshard: int):
    """Two-column key (int32, float32): packs into one 64-bit word on the device path."""
    s = bs.Prefixed(bs.Const(nshard, PA, PB, torch.ones(N, dtype=torch.int64)), 2)
    return bs.Reduce(s, bs.ops.add)


@bs.Func
def bytemax(nshard: int):
    """int8 key, float32 value, max."""
    return bs.Reduce(bs.Const(nshard, P8, (torch.arange(N) % 1000).to(torch.float32)), bs.ops.max)


@bs.Func
def bjoin(nshard: int):
    large = bs.Reshuffle(bs.Const(nshard, KEYS, torch.arange(N, dtype=torch.float32)))
    small = bs.Const(nshard, torch.arange(0, 97, 2), torch.arange(0, 97, 2) * 10)
    return bs.BroadcastJoin(large, small)


@bs.Func
def reshard(nshard: int, out: int):
    return bs.Reshard(bs.Const(nshard, KEYS, torch.arange(N)), out)


@bs.Func
def chain(prev: bs.Slice):
    return bs.Reduce(bs.Map(prev, lambda k, v: (k % 5, v)), bs.ops.add)


@bs.Func
def failing(nshard: int):
    """Shard 1 fails fatally in its Map; the others produce rows."""
    per = 8
    keys = torch.arange(per * nshard)
    keys[per] = -7777  # lands in shard 1 (Const splits rows evenly, in order)

    def check(k, v):
        if bool((k == -7777).any()):
            raise bs.errors.fatal("disk on fire")
        return k % 5, v
    return bs.Reduce(bs.Map(bs.Const(nshard, keys, torch.ones(per * nshard, dtype=torch.int64)), check),
                     bs.ops.add)


def failure_check(sess, ex, world, rank):
    """A fatal producer error on one rank fails the invocation on every rank
    (status vote: gloo on CPU, carried by the exchange on GPUs) and the session
    keeps working afterwards."""
    from bigslice_b200 import errors
    try:
        sess.run(failing, world + 1)
        return False
    except errors.Error as e:
        if not errors.is_fatal(e):
            return False
    rows = ex.gather(sess.run(wordcount, world))
    want = {}
    for k in KEYS.tolist():
        want[k] = want.get(k, 0) + 1
    return dict(rows) == want


def exchange_content_check(ex, world, rank):
    """Rows pushed through the fused exchange must arrive exactly once, intact,
    at the owner of hash(key) % nparts — checked column by column."""
    import torch.distributed as dist
    from bigslice_b200.frame import hashing
    ok = True
    for n, nparts in ((300007, 2 * world + 1), (2500000, world)):
        g = torch.Generator().manual_seed(100 + rank)
        k = torch.randint(-(1 << 40), 1 << 40, (n,), generator=g)
        a = torch.randn(n, generator=g)
        b = torch.randint(0, 1 << 20, (n,), generator=g, dtype=torch.int32)
        for zero_copy in (True, False):
            got = ex.peers.exchange([k.cuda(), a.cuda(), b.cuda()], nparts, nkeys=1, zero_copy=zero_copy)
            mine = {d: [c.cpu().clone() for c in cols] for d, cols in got.items()}
            everything = [None] * world
            dist.all_gather_object(everything, (k, a, b), group=ex.ctrl_group)
            K = torch.cat([e[0] for e in everything])
            A = torch.cat([e[1] for e in everything])
            B = torch.cat([e[2] for e in everything])
            dest = hashing.hash64(K, 0) % nparts
            for d in range(rank, nparts, world):
                sel = dest == d
                wk, order = torch.sort(K[sel], stable=True)
                wa, wb = A[sel][order], B[sel][order]
                rk, rorder = torch.sort(mine[d][0], stable=True)
                ra, rb = mine[d][1][rorder], mine[d][2][rorder]
                if not (torch.equal(wk, rk)):
                    ok = False
                    continue
                # rows with equal keys may arrive in any order: compare per-key multisets
                pa = torch.stack([wk.double(), wa.double(), wb.double()], 1)
                pb = torch.stack([rk.double(), ra.double(), rb.double()], 1)
                if not torch.equal(torch.unique(pa, dim=0), torch.unique(pb, dim=0)):
                    ok = False
    return ok


def table_exchange_check(ex, world, rank):
    """The specialised in-place table exchange: every <key, value> of every rank's
    table arrives exactly once at the owner of hash(key) % nparts, and the clearing
    scatter leaves the source table empty (reusable without an init pass)."""
    import torch.distributed as dist
    from bigslice_b200.frame import hashing
    from bigslice_b200.ops import gpu
    ok = True
    for vdtype, nparts, warm in ((torch.int64, 2 * world + 1, None), (torch.int32, world, 1 << 12),
                                 (torch.float32, 3 * world, 0), (torch.int64, 12 * world + 1, None),
                                 (torch.int32, 1, None)):
        g = torch.Generator().manual_seed(500 + rank)
        n = 400000
        k = torch.randint(-(1 << 40), 1 << 40, (n,), generator=g)
        v = torch.randint(1, 1000, (n,), generator=g).to(vdtype)
        t = gpu.HashTable(2 * n, vdtype, "add", "cuda", warm=warm)
        for rep in range(2):  # second round reuses the cleared table ...
            if rep == 1:
                t.track(nparts)  # ... with destinations counted while combining
            t.combine(k.cuda(), v.cuda())
            got = ex.peers.exchange_table(t, nparts, clear=True)
            mine = {d: [c.cpu().clone() for c in cols] for d, cols in got.items()}
            if int((t.slots[:, 0] != -(1 << 63)).sum()) != 0:
                ok = False  # not cleared
            if int((t.slots[:, 1] != 0).sum()) != 0:
                ok = False  # values not reset to the identity
            t.reset_status()
            uk, inv = torch.unique(k, return_inverse=True)
            uv = torch.zeros(len(uk), dtype=torch.float64).index_add_(0, inv, v.double())
            everything = [None] * world
            dist.all_gather_object(everything, (uk, uv), group=ex.ctrl_group)
            for d in range(rank, nparts, world):
                want = []
                for ek, ev in everything:
                    sel = hashing.hash64(ek, 0) % nparts == d
                    want.append(torch.stack([ek[sel].double(), ev[sel]], 1))
                want = torch.cat(want)
                have = torch.stack([mine[d][0].double(), mine[d][1].double()], 1)
                if want.shape != have.shape:
                    ok = False
                    continue
                w = want[torch.argsort(want[:, 0] * 1e6 + want[:, 1])]
                h = have[torch.argsort(have[:, 0] * 1e6 + have[:, 1])]
                if not torch.equal(torch.unique(w, dim=0), torch.unique(h, dim=0)):
                    ok = False
    return ok


def exchange_vote_check(ex, world, rank):
    """The vote that travels with the counts: one rank reporting failure (or a
    different fingerprint) makes the exchange raise on every rank, and the next
    exchange works again."""
    from bigslice_b200.ops import gpu
    from bigslice_b200.parallel.peer import ExchangeVoteError
    t = gpu.HashTable(4096, torch.int64, "add", "cuda")
    k = torch.arange(1000, device="cuda")
    kinds = []
    for vote in ((1 if rank == world - 1 else 0, 5), (0, 100 + rank), (0, 9)):
        t.combine(k, k)
        try:
            got = ex.peers.exchange_table(t, world, clear=True, vote=vote)
            kinds.append(0)
            n = sum(int(c[0].numel()) for c in got.values())
        except ExchangeVoteError as e:
            kinds.append(e.kind)
            t.reset()
            continue
        t.reset_status()
    want = [3, 4 if world > 1 else 0, 0]
    return kinds == want


def big_wordcount_check(sess, ex, world, rank):
    """A few steps of a 6M-row/rank Zipf wordcount through the table path (warm
    + main levels, in-place table exchange, combine on arrival): totals and the
    per-key counts must be exact on every step."""
    import torch.distributed as dist
    from bigslice_b200.models import wordcount as wc
    from bigslice_b200.ops import gpu
    gpu.HashTable.WARM_SLOTS = 1 << 16  # force the two-level path at this scale
    n = 6000000
    keys = wc.zipf_keys(n, 1 << 21, 1.0, seed=77 + rank, device="cuda")
    vals = torch.ones(n, dtype=torch.int64, device="cuda")
    empty = (torch.empty(0, dtype=torch.int64, device="cuda"),) * 2
    # two shards per rank (s and s + world): the exchange + combine-on-arrival path
    # runs even on a single GPU
    half = n // 2
    mine = {rank: (keys[:half], vals[:half]), rank + world: (keys[half:], vals[half:])}
    wc.register_source("big", wc.DeviceRows([mine.get(s, empty) for s in range(2 * world)]))
    allk = [None] * world
    dist.all_gather_object(allk, keys.cpu(), group=ex.ctrl_group)
    uk, cnt = torch.unique(torch.cat(allk), return_counts=True)
    ok = True
    for step in range(3):
        res = sess.run(wc.wordcount, "big")
        f = res.to_frame()
        local = dict(zip(f.col(0).tolist(), f.col(1).tolist())) if len(f) else {}
        everyone = [None] * world
        dist.all_gather_object(everyone, local, group=ex.ctrl_group)
        merged = {}
        for d in everyone:
            for k, v in d.items():
                if k in merged:
                    ok = False  # a key must live on exactly one rank
                merged[k] = v
        if len(merged) != len(uk) or sum(merged.values()) != n * world:
            ok = False
        else:
            mk = torch.tensor(sorted(merged))
            mv = torch.tensor([merged[k] for k in mk.tolist()])
            ok = ok and torch.equal(mk, uk) and torch.equal(mv, cnt)
        res.discard()
    gpu.HashTable.WARM_SLOTS = 1 << 22
    return ok


FLAKY = {"left": 0}


@bs.Func
def flaky_count(nshard: int):
    """Shard 0's Map fails once with a *non-fatal* error: the task is LOST and
    retried by the evaluator; the invocation must still succeed."""
    keys = torch.arange(64 * nshard) % 37

    def once(k, v):
        if FLAKY["left"] > 0 and bool((k == 0).any()):
            FLAKY["left"] -= 1
            raise bs.errors.Error("transient: connection reset by peer", bs.errors.Kind.Net)
        return k, v
    return bs.Reduce(bs.Map(bs.Const(nshard, keys, torch.ones(64 * nshard, dtype=torch.int64)), once,
                            out=[bs.int64, bs.int64]), bs.ops.add)


def recovery_check(sess, ex, world, rank):
    FLAKY["left"] = 1 if rank == 0 else 0   # only rank 0's producer stumbles
    rows = dict(ex.gather(sess.run(flaky_count, world)))
    want = {}
    for k in (torch.arange(64 * world) % 37).tolist():
        want[k] = want.get(k, 0) + 1
    return rows == want and FLAKY["left"] == 0


def async_table_checks(sess, ex, world, rank):
    """The asynchronous Reduce exchange: sentinel key on one rank only, more
    shards than ranks with all-distinct keys (table sizing), cardinality hints
    reused on the second run, landing-region overflow recovered by growing the
    arena, CUDA-graph replays."""
    import torch.distributed as dist
    from bigslice_b200.models import wordcount as wc
    out = {}
    dev = "cuda"
    empty = (torch.empty(0, dtype=torch.int64, device=dev),) * 2
    # (a) the reserved EMPTY key (INT64_MIN) present on rank 0 only
    n = 50000
    k = torch.arange(n, device=dev) * 7 + rank
    if rank == 0:
        k[:5] = -(1 << 63)
    v = torch.ones(n, dtype=torch.int64, device=dev)
    wc.register_source("sentinel", wc.DeviceRows([(k, v) if s == rank else empty for s in range(world)]))
    got = dict(ex.gather(sess.run(wc.wordcount, "sentinel")))
    allk = [None] * world
    dist.all_gather_object(allk, k.cpu(), group=ex.ctrl_group)
    uk, cnt = torch.unique(torch.cat(allk), return_counts=True)
    out["sentinel_ok"] = got == dict(zip(uk.tolist(), cnt.tolist())) and got.get(-(1 << 63)) == 5
    # (b) 4 shards per rank, every key distinct: the shared table must hold them all
    S = 4
    n = 200000
    base = torch.arange(n, device=dev) * (world * S)
    shards = []
    for s in range(world * S):
        shards.append((base + s, torch.ones(n, dtype=torch.int64, device=dev)) if s % world == rank else empty)
    wc.register_source("distinct", wc.DeviceRows(shards))
    ok = True
    for step in range(3):  # first run sizes exactly, later runs use the hints (and graphs)
        res = sess.run(wc.wordcount, "distinct")
        f = res.to_frame()
        tot = torch.tensor([len(f), int(f.col(1).sum()) if len(f) else 0], dtype=torch.int64)
        dist.all_reduce(tot, group=ex.ctrl_group)
        ok = ok and tot.tolist() == [n * world * S, n * world * S]
        res.discard()
    out["distinct_ok"] = ok
    out["hints_ok"] = any(k[0] == "devicerows" or True for k in ex.card_hints) and len(ex.card_hints) > 0
    out["graph_launches"] = ex.rx.stats["graph_launches"] if ex.rx is not None else 0
    # (c) landing overflow: shrink the arena so the regions are too small, run, and
    # expect the lost exchange to be repeated with a grown arena
    if ex.rx is not None:
        small = ex.rx.bytes_for(world * S, 1000)
        ex._want_arena = 0
        ex.rx.reset(small)
        before = ex.rx.arena_bytes
        res = sess.run(wc.wordcount, "distinct")
        f = res.to_frame()
        tot = torch.tensor([len(f)], dtype=torch.int64)
        dist.all_reduce(tot, group=ex.ctrl_group)
        out["overflow_recovered_ok"] = int(tot[0]) == n * world * S and ex.rx.arena_bytes > before
        res.discard()
    # (d) a destination overflow on ONE rank only: partition 1 (owned by rank 1 % world)
    # receives three times the distinct keys the hints of this call site promise.  The
    # exchange's closing agreement gives every rank the same verdict, all of them repeat
    # the exchange (exactly sized this time) — nobody retries on its own.
    if ex.rx is not None:
        from bigslice_b200.ops import gpu
        P = world * S
        ex._want_arena = 0
        ex.rx.reset(ex.rx.bytes_for(P, 4 * n))     # landing regions are not the limit here
        cand = torch.arange(0, 8 * n * P, device=dev) * 3 + 1000003
        cand = cand[(gpu.hash_column(cand) % P) == 1][:3 * n]
        assert cand.numel() == 3 * n
        mine = cand[rank::world].contiguous()
        shards = [(mine, torch.ones(mine.numel(), dtype=torch.int64, device=dev)) if s == rank else empty
                  for s in range(P)]
        wc.register_source("skew", wc.DeviceRows(shards))
        before = len(ex.exchange_log)
        res = sess.run(wc.wordcount, "skew")
        f = res.to_frame()
        tot = torch.tensor([len(f), int(f.col(1).sum()) if len(f) else 0], dtype=torch.int64)
        dist.all_reduce(tot, group=ex.ctrl_group)
        attempts = torch.tensor([len(ex.exchange_log) - before], dtype=torch.int64)
        lo, hi = attempts.clone(), attempts.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=ex.ctrl_group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=ex.ctrl_group)
        out["dest_overflow_attempts"] = [int(lo[0]), int(hi[0])]
        out["dest_overflow_recovered_ok"] = tot.tolist() == [3 * n, 3 * n] and int(lo[0]) == int(hi[0]) == 2
        res.discard()
    return out


def main():
    import faulthandler
    # a hang prints every thread's stack and exits instead of running into the
    # harness timeout with no information
    faulthandler.dump_traceback_later(int(os.environ.get("BS_TEST_WATCHDOG", "420")), exit=True)
    device = os.environ.get("BS_TEST_DEVICE", "cpu")
    shuffle = os.environ.get("BS_TEST_SHUFFLE", "auto")
    sess = bs.start(bs.Spmd(device=None if device == "cpu" else "cuda", shuffle=shuffle,
                            arena_bytes=64 << 20), bs.Parallelism(2))
    ex = sess.executor
    world, rank = ex.world, ex.rank
    out = {"rank": rank, "world": world}
    sess.run(const_map_scan, 2 * world)
    out["scan_local"] = sorted(scanned)
    r = sess.run(wordcount, world)
    rows = ex.gather(r)
    want = {}
    for k in KEYS.tolist():
        want[k] = want.get(k, 0) + 1
    out["wordcount_ok"] = dict(rows) == want
    out["wordcount_local_keys"] = sorted(k for k, _ in r.rows())
    r2 = sess.run(chain, r)
    want2 = {}
    for k, v in want.items():
        want2[k % 5] = want2.get(k % 5, 0) + v
    out["chain_ok"] = dict(ex.gather(r2)) == want2
    rows = ex.gather(sess.run(strcount, world))
    wants = {}
    for w in WORDS.tolist():
        wants[w] = wants.get(w, 0) + 1
    out["strcount_ok"] = dict(rows) == wants
    rows = ex.gather(sess.run(textcount, 2 * world))
    names = WORD_DICT.lookup_global([k for k, _ in rows], group=ex.ctrl_group)
    out["textcount_ok"] = {names[k]: v for k, v in rows} == wants
    out["textcount_mode"] = ex.exchange_log[-1]["mode"] if ex.exchange_log else ""
    rows = ex.gather(sess.run(join, 2 * world))
    got = {k: (sorted(a), sorted(b)) for k, a, b in rows}
    wantj = {}
    for i, k in enumerate(KEYS.tolist()):
        wantj.setdefault(k, ([], []))[0].append(float(i))
    for k in range(0, 97, 2):
        wantj.setdefault(k, ([], []))[1].append(k * 10)
    out["join_ok"] = got == {k: (sorted(a), sorted(b)) for k, (a, b) in wantj.items()}
    rows = ex.gather(sess.run(pairsum, world))
    wantp = {}
    for a, b in zip(PA.tolist(), PB.tolist()):
        wantp[(a, b)] = wantp.get((a, b), 0) + 1
    out["pairsum_ok"] = {(a, b): c for a, b, c in rows} == wantp and len(rows) == len(wantp)
    out["pairsum_mode"] = ex.exchange_log[-1]["mode"] if ex.exchange_log else ""
    rows = ex.gather(sess.run(bytemax, 2 * world))
    wantm = {}
    for k, v in zip(P8.tolist(), (torch.arange(N) % 1000).tolist()):
        wantm[k] = max(wantm.get(k, -1.0), float(v))
    out["bytemax_ok"] = dict(rows) == wantm
    out["bytemax_mode"] = ex.exchange_log[-1]["mode"] if ex.exchange_log else ""
    rows = ex.gather(sess.run(bjoin, 2 * world))
    gotb = {k: (sorted(a), sorted(b)) for k, a, b in rows}
    out["bjoin_ok"] = (len(rows) == len(gotb) and
                       gotb == {k: (sorted(a), sorted(b)) for k, (a, b) in wantj.items() if a})
    res = sess.run(reshard, world, 2 * world + 1)
    rows = ex.gather(res)
    out["reshard_ok"] = sorted(rows) == sorted(zip(KEYS.tolist(), range(N)))
    local = res.rows()
    from bigslice_b200.frame import hashing
    out["reshard_placement_ok"] = all(
        int(hashing.hash64(torch.tensor([k]), 0)[0]) % (2 * world + 1) % world == rank
        for k, _ in local)
    if device != "cpu" and ex.peers is not None:
        out["exchange_content_ok"] = exchange_content_check(ex, world, rank)
        out["table_exchange_ok"] = table_exchange_check(ex, world, rank)
        out["exchange_vote_ok"] = exchange_vote_check(ex, world, rank)
        out["bigcount_ok"] = big_wordcount_check(sess, ex, world, rank)
        out.update(async_table_checks(sess, ex, world, rank))
    out["recovery_ok"] = recovery_check(sess, ex, world, rank)
    out["failure_ok"] = failure_check(sess, ex, world, rank)
    out["modes"] = sorted(set(e["mode"] for e in ex.exchange_log))
    try:
        from bigslice_b200.ops import gpu
        out["launches"] = gpu.launches() if device != "cpu" else 0
    except Exception:
        out["launches"] = 0
    sess.shutdown()
    faulthandler.cancel_dump_traceback_later()
    path = os.environ["BS_TEST_OUT"] + ".%d" % rank
    with open(path, "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
