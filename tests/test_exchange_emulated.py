"""The multi-rank Reduce exchange protocol, run on the CPU.

``csrc/reduce_exchange.cu`` and ``csrc/hash_combine.cu`` are compiled unmodified
against the SIMT emulator (tests/emu).  The emulator's state is thread-local, so
every *host thread is one GPU*: two (or three) ranks run their kernels
concurrently in one process, with each other's control blocks and landing arenas
mapped as ordinary memory — the same pointers a rank gets from CUDA IPC on a real
box.  ``rx_begin / rx_scatter / rx_wait / rx_combine / table_compact / rx_end``
then exchange data and flags exactly as they do over NVLink:

* rows arrive at the right rank, in the right shard (and table region), combined;
* parity halves and acknowledgements carry several epochs;
* every device-detected failure is the *same verdict on every rank* — a landing
  overflow, a vote mismatch, a failed producer, and in particular a destination
  overflow that only one rank can see (the closing agreement of ``rx_end``);
* a rank that never shows up is a bounded timeout on the others, not a hang.
"""
import ctypes
import os
import shutil
import subprocess
import sys
import threading

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emu  # noqa: E402

from bigslice_b200.frame import hashing  # noqa: E402
from bigslice_b200.ops import gpu  # noqa: E402
from bigslice_b200.parallel import rx  # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
c_void_p, c_int, c_ll, c_ull = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong
EMPTY = -(1 << 63)


class Lib:
    def __init__(self):
        asan = bool(os.environ.get("BS_EMU_ASAN"))
        self.libs = [ctypes.CDLL(emu.build(f, asan=asan)) for f in ("reduce_exchange", "hash_combine")]

    def __getattr__(self, name):
        for lib in self.libs:
            try:
                return getattr(lib, name)
            except AttributeError:
                continue
        raise AttributeError(name)


@pytest.fixture
def lib(monkeypatch):
    lib = Lib()
    monkeypatch.setattr(gpu, "require", lambda: lib)
    monkeypatch.setattr(gpu, "_lib", lib)
    monkeypatch.setattr(gpu, "_stream", lambda: c_void_p(0))
    return lib


def ptr(t):
    return c_void_p(t.data_ptr())


class Rank:
    """One emulated GPU: control block, landing arena, device state, report ring."""

    def __init__(self, lib, world, rank, arena_bytes, timeout_s=20.0):
        self.lib, self.world, self.rank = lib, world, rank
        sizes = [c_int() for _ in range(4)]
        lib.bs_rx_sizes(*[ctypes.byref(s) for s in sizes])
        self.ctrl_bytes, self.state_bytes, self.report_bytes, peers_b = [s.value for s in sizes]
        assert peers_b == ctypes.sizeof(rx.RxPeers)
        offs = [c_ll() for _ in range(3)]
        lib.bs_rx_report_offsets(*[ctypes.byref(o) for o in offs])
        self.report_offs = tuple(o.value for o in offs)
        self.ctrl = torch.zeros(self.ctrl_bytes, dtype=torch.uint8)
        self.arena = torch.zeros(arena_bytes, dtype=torch.uint8)
        self.state = torch.zeros(self.state_bytes, dtype=torch.uint8)
        self.ring = torch.zeros(self.report_bytes * rx.NSLOTS // 8, dtype=torch.int64)
        self.arena_bytes = arena_bytes
        self.timeout_ns = int(timeout_s * 1e9)
        self.epoch = 0
        self.peers = None

    def connect(self, ranks):
        p = rx.RxPeers()
        p.world, p.rank = self.world, self.rank
        for r, other in enumerate(ranks):
            p.ctrl[r], p.arena[r] = other.ctrl.data_ptr(), other.arena.data_ptr()
        self.peers = p
        assert self.lib.bs_rx_state_init(ptr(self.state), c_ull(self.timeout_ns), c_void_p(0)) == 0

    def geometry(self, nparts, regions):
        pl = (nparts + self.world - 1) // self.world
        half = (self.arena_bytes // 2) & ~0xFF
        p = self.peers
        p.shards_per_rank, p.regions = pl * regions, regions
        p.region_rows = half // (pl * regions * self.world * 16)
        p.half_bytes = half
        return pl

    def exchange(self, keys, vals, nparts, op="add", regions=1, vote=(0, 0), dest_cap=None, out_cap=None,
                 two_part=False, keymode=0):
        """One full exchange of this rank's rows; returns (report, {shard: (keys, vals)}).
        ``two_part``: the first-run shape of the host code — scatter + wait, a report
        without acknowledgement (the host sizes the destinations from it), then combine,
        compaction and the closing ``rx_end``."""
        lib, pe = self.lib, self.peers
        pl = self.geometry(nparts, regions)
        vdt = vals.dtype
        vt, opc = c_int(gpu.VTYPE[vdt]), c_int(gpu.OPS[op])
        table = gpu.HashTable(max(2048, 4 * keys.numel()), vdt, op, "cpu", warm=0)
        if keys.numel():
            table.combine(keys, vals)
        bits = c_ull(table.identity_bits())
        st = ptr(self.state)
        assert lib.bs_rx_set_vote(st, c_ull(vote[0]), c_ull(vote[1]), c_void_p(0)) == 0
        assert lib.bs_rx_scatter(ctypes.byref(pe), st, ptr(table.slots), c_ll(table.total), c_int(nparts), c_int(1),
                                 bits, ptr(table.status), c_int(keymode), c_void_p(0)) == 0
        assert lib.bs_rx_wait(ctypes.byref(pe), st, c_int(nparts), c_void_p(0)) == 0
        locals_ = [j for j in range(pl) if self.rank + j * self.world < nparts]
        if two_part:
            none, zero = (c_void_p * pl)(), (c_ll * pl)()
            assert lib.bs_rx_end(ctypes.byref(pe), st, c_int(pl), none, zero, c_ull(0), ptr(self.ring),
                                 c_int(rx.NSLOTS), c_int(0), c_void_p(0)) == 0
            words = self.ring.numpy().reshape(rx.NSLOTS, self.report_bytes // 8)[(self.epoch + 1) % rx.NSLOTS].copy()
            probe = rx.Report(words, self.report_offs)
            assert probe.ticket == self.epoch + 1
            self.probe_rows = probe.recv_rows[:pl]
        stp, caps = (c_void_p * max(len(locals_), 1))(), (c_ll * max(len(locals_), 1))()
        dests = []
        max_rows = int(pe.region_rows) * self.world
        for i, j in enumerate(locals_):
            dt = gpu.HashTable(dest_cap or 8192, vdt, op, "cpu", warm=0)
            cap = out_cap or dt.total
            ok, ov = torch.full((cap,), EMPTY), torch.zeros(cap, dtype=vdt)
            rcap = dt.total // regions
            for r in range(regions):
                assert lib.bs_rx_combine(ctypes.byref(pe), st, c_int(j * regions + r), vt, opc,
                                         c_void_p(dt.slots.data_ptr() + r * rcap * 16), c_ll(rcap), ptr(dt.status),
                                         c_ll(max_rows), c_void_p(0)) == 0
            assert lib.bs_table_compact(ptr(dt.slots), c_ll(dt.total), vt, ptr(ok), ptr(ov), ptr(dt.status), c_int(1),
                                        bits, c_ll(cap), c_void_p(0)) == 0
            stp[i], caps[i] = dt.status.data_ptr(), cap
            dests.append((j, dt, ok, ov))
        assert lib.bs_rx_end(ctypes.byref(pe), st, c_int(len(locals_)), stp, caps, bits, ptr(self.ring),
                             c_int(rx.NSLOTS), c_int(1), c_void_p(0)) == 0
        self.epoch += 1
        words = self.ring.numpy().reshape(rx.NSLOTS, self.report_bytes // 8)[self.epoch % rx.NSLOTS].copy()
        rep = rx.Report(words, self.report_offs)
        assert rep.ticket == self.epoch
        out = {}
        for i, (j, dt, ok, ov) in enumerate(dests):
            n = rep.out_rows[i]
            out[self.rank + j * self.world] = (ok[:n].clone(), ov[:n].clone())
        return rep, out


def run_ranks(world, fn, arena_bytes=1 << 20, timeout_s=20.0, lib=None, absent=()):
    """Run ``fn(rank_object)`` on one thread per rank; returns the per-rank results."""
    ranks = [Rank(lib, world, r, arena_bytes, timeout_s) for r in range(world)]
    for r in ranks:
        r.connect(ranks)
    res, errs = [None] * world, []

    def body(r):
        try:
            res[r.rank] = fn(r)
        except BaseException as e:  # noqa: BLE001
            errs.append((r.rank, e))
    ts = [threading.Thread(target=body, args=(r,)) for r in ranks if r.rank not in absent]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert not any(t.is_alive() for t in ts), "an emulated rank hangs"
    assert not errs, errs
    return res


def rows_of(rank, world, n=6000, seed=0):
    g = torch.Generator().manual_seed(100 * seed + rank)
    k = torch.randint(0, 900, (n,), generator=g) * 104729 - 77
    k[::131] = EMPTY
    v = torch.randint(1, 5, (n,), generator=g)
    return k, v


def reference(world, nparts, op="add", seed=0, n=6000):
    ks, vs = zip(*[rows_of(r, world, n, seed) for r in range(world)])
    k, v = torch.cat(ks), torch.cat(vs)
    uk, inv = torch.unique(k, return_inverse=True)
    red = {"add": "sum", "min": "amin", "max": "amax"}[op]
    acc = torch.zeros(uk.numel(), dtype=v.dtype).scatter_reduce_(0, inv, v, red, include_self=False)
    dest = hashing.hash_tensor(uk, 0) % nparts
    return {p: dict(zip(uk[dest == p].tolist(), acc[dest == p].tolist())) for p in range(nparts)}


@pytest.mark.parametrize("world,nparts,regions", [(1, 3, 1), (2, 4, 1), (2, 2, 2), (3, 5, 1)])
def test_rows_reach_their_shard_combined_over_several_epochs(lib, world, nparts, regions):
    steps = 4   # parity halves alternate; from the third epoch on rx_begin waits for acknowledgements

    def body(r):
        got = []
        for s in range(steps):
            k, v = rows_of(r.rank, world, seed=s)
            rep, out = r.exchange(k, v, nparts, regions=regions)
            assert rep.error == 0, rep.error
            got.append(out)
        return got
    res = run_ranks(world, body, lib=lib)
    for s in range(steps):
        want = reference(world, nparts, seed=s)
        seen = {}
        for r in range(world):
            for p, (k, v) in res[r][s].items():
                assert p % world == r                      # shard p lives on rank p % world
                seen[p] = dict(zip(k.tolist(), v.tolist()))
        assert seen == want


def test_first_run_and_steady_state_paths_interoperate(lib):
    """One rank takes the two-part first-run path (probe report, then combine) while the
    other runs the whole sequence at once, with the destination tables split into two
    regions: both see the same epochs, flags and landing layout."""
    def body(r):
        got = []
        for s in range(3):
            k, v = rows_of(r.rank, 2, seed=s)
            rep, out = r.exchange(k, v, 4, regions=2, two_part=(r.rank + s) % 2 == 0)
            assert rep.error == 0
            if (r.rank + s) % 2 == 0:   # the probe said how many rows arrived for every local shard
                assert r.probe_rows[:2] == rep.recv_rows[:2] and sum(r.probe_rows) > 0
            got.append(out)
        return got
    res = run_ranks(2, body, lib=lib)
    for s in range(3):
        want = reference(2, 4, seed=s)
        seen = {p: dict(zip(k.tolist(), v.tolist())) for r in range(2) for p, (k, v) in res[r][s].items()}
        assert seen == want


def test_packed_keys_are_placed_like_the_frame_hash_of_their_columns(lib):
    """Narrow and two-column key prefixes ride the tables packed into one word
    (ops/keypack.py); the scatter kernel's key modes must send every packed key to the
    shard ``Frame.hash`` assigns to its *columns*."""
    from bigslice_b200 import slicetype
    from bigslice_b200.frame import Frame
    from bigslice_b200.ops import keypack
    g = torch.Generator().manual_seed(4)
    n, nparts = 4000, 5
    cases = {
        "int32": [torch.randint(-500, 500, (n,), generator=g).to(torch.int32)],
        "int16+uint8": [torch.randint(-300, 300, (n,), generator=g).to(torch.int16),
                        torch.randint(0, 200, (n,), generator=g).to(torch.uint8)],
        "float32+int32": [(torch.randint(-9, 9, (n,), generator=g).to(torch.float32) * 0.25),
                          torch.randint(-9, 9, (n,), generator=g).to(torch.int32)],
        "float64": [torch.randint(-99, 99, (n,), generator=g).to(torch.float64) * 0.5],
    }
    for name, cols in cases.items():
        typ = slicetype.Type([slicetype.dtype_of(c.dtype) for c in cols] + [slicetype.int64], len(cols))
        lay = keypack.layout_for(typ)
        assert lay is not None, name
        word = keypack.pack(cols, lay)
        ones = torch.ones(n, dtype=torch.int64)
        want_dest = Frame(cols + [ones], typ).hash(0) % nparts

        def body(r, word=word, lay=lay):
            return r.exchange(word[r.rank::2], ones[r.rank::2], nparts, keymode=lay.mode)
        res = run_ranks(2, body, lib=lib)
        got = {}
        for rep, out in res:
            assert rep.error == 0
            for p, (k, v) in out.items():
                for kk, vv in zip(k.tolist(), v.tolist()):
                    got[kk] = (p, vv)
        uw, inv = torch.unique(word, return_inverse=True)
        cnt = torch.bincount(inv)
        first = torch.zeros(uw.numel(), dtype=torch.int64).scatter_(0, inv, torch.arange(n))
        assert got == {int(w): (int(want_dest[i]), int(c)) for w, i, c in zip(uw.tolist(), first.tolist(), cnt.tolist())}, name
        # and the packed word unpacks to the original key columns
        back = keypack.unpack(word, lay)
        assert all(torch.equal(a, b) for a, b in zip(back, cols)), name


def test_min_and_float_values(lib):
    def body(r):
        k, v = rows_of(r.rank, 2)
        return r.exchange(k, v.to(torch.float32) * 0.5, 2, op="min")
    res = run_ranks(2, body, lib=lib)
    ks, vs = zip(*[rows_of(r, 2) for r in range(2)])
    k, v = torch.cat(ks), torch.cat(vs).to(torch.float32) * 0.5
    uk, inv = torch.unique(k, return_inverse=True)
    acc = torch.zeros(uk.numel()).scatter_reduce_(0, inv, v, "amin", include_self=False)
    want = dict(zip(uk.tolist(), acc.tolist()))
    got = {}
    for rep, out in res:
        assert rep.error == 0
        for p, (kk, vv) in out.items():
            got.update(zip(kk.tolist(), vv.tolist()))
    assert got == want


def test_every_rank_reports_the_same_verdict(lib):
    # (a) a destination overflow only rank 1 can see: its result buffer holds 16 rows,
    # rank 0's own destination is fine — without rx_end's closing agreement rank 0
    # would report success and the two ranks would part ways
    def dest_overflow(r):
        k, v = rows_of(r.rank, 2)
        rep, _ = r.exchange(k, v, 2, out_cap=16 if r.rank == 1 else None)
        return rep.error
    assert run_ranks(2, dest_overflow, lib=lib) == [rx.DEST_OVERFLOW, rx.DEST_OVERFLOW]
    # (b) the ranks disagree about which invocation this is
    def fingerprints(r):
        k, v = rows_of(r.rank, 2)
        return r.exchange(k, v, 2, vote=(0, 1000 + r.rank))[0].error
    assert run_ranks(2, fingerprints, lib=lib) == [rx.FINGERPRINT, rx.FINGERPRINT]
    # (c) one rank's producers failed
    def producer_failed(r):
        k, v = rows_of(r.rank, 3)
        return r.exchange(k, v, 3, vote=(1 if r.rank == 2 else 0, 7))[0].error
    assert run_ranks(3, producer_failed, lib=lib) == [rx.PRODUCER_FAILED] * 3
    # (d) landing regions too small: everybody learns how many rows would have been needed
    def landing(r):
        k, v = rows_of(r.rank, 2)
        rep, _ = r.exchange(k, v, 2)
        return rep.error, rep.need_rows
    res = run_ranks(2, landing, arena_bytes=16 << 10, lib=lib)    # 128 rows per region
    assert [e for e, _ in res] == [rx.LANDING_OVERFLOW] * 2
    assert res[0][1] == res[1][1] and res[0][1] > 128
    # ... and the exchange after a failed one works (flags, parity and tables were left consistent)
    def fail_then_succeed(r):
        k, v = rows_of(r.rank, 2)
        first, _ = r.exchange(k, v, 2, vote=(0, r.rank))          # fingerprint mismatch
        second, out = r.exchange(k, v, 2, vote=(0, 5))
        return first.error, second.error, out
    res = run_ranks(2, fail_then_succeed, lib=lib)
    assert [(a, b) for a, b, _ in res] == [(rx.FINGERPRINT, 0)] * 2
    want = reference(2, 2)
    assert {p: dict(zip(k.tolist(), v.tolist())) for _, _, out in res for p, (k, v) in out.items()} == want


def test_a_rank_that_never_arrives_is_a_bounded_timeout(lib):
    def body(r):
        k, v = rows_of(r.rank, 2)
        return r.exchange(k, v, 2)[0].error
    res = run_ranks(2, body, timeout_s=0.5, lib=lib, absent=(1,))
    assert res[0] == rx.TIMEOUT


@pytest.mark.skipif(not os.path.exists(emu.asan_preload().split(":")[0]), reason="needs libasan")
def test_the_protocol_is_clean_under_address_sanitizer():
    """The data-path scenarios once more with the kernels built with -fsanitize=address:
    landing regions, staging buffers, tables and reports are never accessed out of bounds."""
    env = dict(os.environ, BS_EMU_ASAN="1", LD_PRELOAD=emu.asan_preload(),
               ASAN_OPTIONS="detect_leaks=0:abort_on_error=1:detect_stack_use_after_return=0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", __file__,
                        "-k", "several_epochs or same_verdict"], env=env, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "passed" in r.stdout and "AddressSanitizer" not in r.stderr


# ---------------------------------------------------------------------------
# the generic (any columns) exchange of csrc/exchange.cu, through parallel/peer.py

def _peer_groups(xlib, world, arena_bytes, monkeypatch):
    """PeerGroup objects wired to each other without CUDA IPC / torch.distributed:
    what ``PeerGroup.__init__`` sets up, over ordinary memory."""
    from bigslice_b200.parallel import peer
    monkeypatch.setattr(gpu, "require", lambda: xlib)
    monkeypatch.setattr(gpu, "_stream", lambda: c_void_p(0))
    sizes = [c_int() for _ in range(4)]
    xlib.bs_xc_sizes(*[ctypes.byref(s) for s in sizes])
    ctrl_bytes, plan_bytes, peers_b, spec_b = [s.value for s in sizes]
    assert peers_b == ctypes.sizeof(peer.XcPeers) and spec_b == ctypes.sizeof(peer.XcSpec)
    offs = [c_ll() for _ in range(5)]
    xlib.bs_xc_plan_offsets(*[ctypes.byref(o) for o in offs])
    ctrls = [torch.zeros(ctrl_bytes, dtype=torch.uint8) for _ in range(world)]
    arenas = [torch.zeros(arena_bytes, dtype=torch.uint8) for _ in range(world)]
    groups = []
    for r in range(world):
        pg = object.__new__(peer.PeerGroup)
        pg.lib, pg.group, pg.world, pg.rank, pg.device = xlib, None, world, r, torch.device("cpu")
        pg.ctrl_bytes, pg.plan_bytes = ctrl_bytes, plan_bytes
        (pg.off_recv_total, pg.off_recv_off, pg.off_arena_end, pg.off_send_count, pg.off_error) = [o.value for o in offs]
        pg.arena_bytes = arena_bytes
        pg._ctrl, pg._arena = c_void_p(ctrls[r].data_ptr()), c_void_p(arenas[r].data_ptr())
        pg.peers = peer.XcPeers()
        pg.peers.world, pg.peers.rank, pg.peers.arena_bytes = world, r, arena_bytes
        for q in range(world):
            pg.peers.ctrl[q], pg.peers.arena[q] = ctrls[q].data_ptr(), arenas[q].data_ptr()
        pg._opened, pg.epoch, pg.last_sent, pg.barrier_epoch = [], 0, 0, 0
        pg.plan = torch.zeros(plan_bytes, dtype=torch.uint8)
        pg.err = torch.zeros(2, dtype=torch.int32)
        pg.arena_view = arenas[r]
        pg.stats = {"exchanges": 0, "bytes_sent": 0, "rows_sent": 0}
        pg._keep = (ctrls, arenas)
        groups.append(pg)
    return groups


def _threads(world, fn):
    res, errs = [None] * world, []

    def body(r):
        try:
            res[r] = fn(r)
        except BaseException as e:  # noqa: BLE001
            errs.append((r, repr(e)))
    ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert not any(t.is_alive() for t in ts), "an emulated rank hangs"
    assert not errs, errs
    return res


@pytest.mark.parametrize("world,nparts", [(1, 3), (2, 5), (3, 3)])
def test_generic_exchange_moves_every_row_to_the_owner_of_its_shard(world, nparts, monkeypatch):
    """Reshard/Reshuffle/Cogroup shuffles: hash -> count -> plan (count all-gather by
    peer stores) -> scatter -> wait, for a frame of mixed-width columns keyed on two
    columns, then with explicit destination ids; two epochs each."""
    xlib = ctypes.CDLL(emu.build("exchange", asan=bool(os.environ.get("BS_EMU_ASAN"))))
    groups = _peer_groups(xlib, world, 1 << 20, monkeypatch)

    def cols_of(r, seed):
        g = torch.Generator().manual_seed(1000 * seed + r)
        n = 3000 + 17 * r
        return [torch.randint(-50, 50, (n,), generator=g),                       # int64 key
                torch.randint(0, 9, (n,), generator=g).to(torch.int32),          # int32 key
                torch.randn(n, generator=g),                                     # float32 payload
                torch.randn(n, 3, generator=g, dtype=torch.float64)]             # 24-byte rows

    def body(r):
        out = []
        for seed in range(2):
            cols = cols_of(r, seed)
            out.append(groups[r].exchange(cols, nparts, nkeys=2))
        ids = (torch.arange(cols[0].numel()) * 7 + r) % nparts
        out.append(groups[r].exchange(cols, nparts, nkeys=2, ids=ids))
        return out
    res = _threads(world, body)

    def rows(cols):
        return sorted(zip(cols[0].tolist(), cols[1].tolist(), cols[2].tolist(), map(tuple, cols[3].tolist())))
    for step in range(3):
        seed = min(step, 1)
        every = [cols_of(r, seed) for r in range(world)]
        for p in range(nparts):
            want = []
            for r, cols in enumerate(every):
                if step < 2:
                    dest = (hashing.hash_tensor(cols[0], 0) ^ hashing.hash_tensor(cols[1], 0)) % nparts
                else:
                    dest = (torch.arange(cols[0].numel()) * 7 + r) % nparts
                sel = dest == p
                want += rows([c[sel] for c in cols])
            got = res[p % world][step][p]
            assert got[3].shape[1:] == (3,) and got[1].dtype == torch.int32
            assert rows(got) == sorted(want), (step, p)
    # no other shard shows up on a rank
    for r in range(world):
        assert set(res[r][0]) == set(range(r, nparts, world))


def test_generic_exchange_vote_and_arena_overflow_fail_on_every_rank(monkeypatch):
    from bigslice_b200.parallel import peer
    xlib = ctypes.CDLL(emu.build("exchange"))
    groups = _peer_groups(xlib, 2, 64 << 10, monkeypatch)
    k = torch.arange(1000)

    def voted(r):
        desc = [(k.data_ptr(), 8, 8, torch.int64, ())]
        kinds = []
        # a failed producer, twice in a row, then different fingerprints, then a clean exchange:
        # a failed vote must leave nothing behind (in particular no abort flag — the plan
        # kernel once set it by reading the vote verdict as the flag-wait verdict)
        for vote in ((1 if r == 1 else 0, 5), (1 if r == 1 else 0, 5), (0, 100 + r), (0, 9)):
            try:
                got = groups[r].exchange_raw(desc, 1000, 2, vote=vote)
                kinds.append(sum(int(c[0].numel()) for c in got.values()))
            except peer.ExchangeVoteError as e:
                kinds.append(-e.kind)
        return kinds
    res = _threads(2, voted)
    assert [x[:3] for x in res] == [[-3, -3, -4]] * 2
    assert res[0][3] + res[1][3] == 2000          # every row of both ranks arrived somewhere

    def too_big(r):
        big = torch.arange(20000)
        try:
            groups[r].exchange([big, big.clone()], 2)
        except peer.ExchangeError as e:
            return "overflow" in str(e)
    assert _threads(2, too_big) == [True, True]


# ---------------------------------------------------------------------------
# parallel/rx.py (the host side of the Reduce exchange) over the emulated library

class _FakeEvent:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass

    def query(self):
        return True


class _FakeStream:
    def __init__(self, *a, **k):
        self.cuda_stream = 0

    def wait_event(self, ev):
        pass


class _Null:
    def __init__(self, *a):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def test_reduce_exchange_host_object_single_rank(monkeypatch):
    """``ReduceExchange`` itself — allocation through the IPC entry points, the
    direct-launch path, first-run probe/finish, reports through the pinned ring,
    purely local publishes, overflow verdicts, a collective reset that grows the
    arena — with the CUDA stream/event objects stubbed and the kernels emulated."""
    class Lib3(Lib):
        def __init__(self):
            self.libs = [ctypes.CDLL(emu.build(f)) for f in ("reduce_exchange", "hash_combine", "exchange")]
    lib = Lib3()
    monkeypatch.setattr(gpu, "require", lambda: lib)
    monkeypatch.setattr(gpu, "_lib", lib)
    monkeypatch.setattr(gpu, "_stream", lambda: c_void_p(0))
    monkeypatch.setattr(gpu, "_pool", gpu.TablePool())
    monkeypatch.setenv("BIGSLICE_NO_GRAPHS", "1")
    for name, fake in (("Stream", _FakeStream), ("Event", _FakeEvent), ("stream", _Null)):
        monkeypatch.setattr(torch.cuda, name, fake)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)

    ex = rx.ReduceExchange(None, 1 << 20, "cpu", timeout_s=5.0)
    try:
        assert ex.world == 1 and ex.arena_bytes == 1 << 20
        nparts = 3
        k, v = rows_of(0, 1, n=5000)
        want = reference(1, nparts, n=5000)

        def table():
            t = gpu.HashTable(16384, torch.int64, "add", "cpu", warm=0)
            t.combine(k, v)
            return t

        def dests(cap_rows):
            out = []
            for j in range(nparts):
                dt = gpu.HashTable(gpu.table_capacity(cap_rows), torch.int64, "add", "cpu", warm=0)
                out.append((j, dt, torch.empty(cap_rows, dtype=torch.int64), torch.empty(cap_rows, dtype=torch.int64),
                            cap_rows))
            return out

        def check(pend, ds):
            got = {}
            for j, dt, ok, ov, cap in ds:
                n = pend.rows(j)
                got[j] = dict(zip(ok[:n].tolist(), ov[:n].tolist()))
            assert got == want
        # first run: probe (rows per shard) -> exactly sized destinations -> finish
        t = table()
        probe = ex.exchange_probe(t, nparts, vote=(0, 7))
        rep = probe.resolve()
        assert rep.error == 0 and rep.recv_rows[:nparts] == [len(want[p]) for p in range(nparts)]
        ds = dests(max(rep.recv_rows[:nparts]) + 16)
        seen = []
        check(ex.exchange_finish(t, nparts, ds, on_report=seen.append), ds)
        assert len(seen) == 1 and seen[0].sent_rows == sum(len(w) for w in want.values())
        # steady state: the whole sequence in one call, several epochs, reports harvested
        for _ in range(3):
            ds = dests(1024)
            check(ex.exchange(table(), nparts, ds, vote=(0, 7)), ds)
        ex.harvest()
        assert not ex.pending and ex.stats["exchanges"] == 4
        # destination buffers too small: a retriable verdict that says what happened
        ds = dests(64)
        pend = ex.exchange(table(), nparts, ds)
        with pytest.raises(rx.ExchangeFailed) as ei:
            pend.rows(0)
        assert ei.value.code == rx.DEST_OVERFLOW and ei.value.retriable
        # landing regions too small (1 MiB arena = 10922 rows per region; 50 000 distinct keys)
        # -> need_rows -> grow the arena with the collective reset -> the same exchange works
        bigk = torch.arange(50000) * 7919 + 13
        bigv = torch.ones(50000, dtype=torch.int64)

        def big_table():
            t = gpu.HashTable(131072, torch.int64, "add", "cpu", warm=0)
            t.combine(bigk, bigv)
            return t
        attempts = 0
        while True:   # the executor's loop: lost -> grow by what the report asked for -> again
            attempts += 1
            ds = dests(32768)
            pend = ex.exchange(big_table(), nparts, ds)
            try:
                assert sum(pend.rows(j) for j in range(nparts)) == 50000
                break
            except rx.ExchangeFailed as e:
                # need_rows is exact (every CTA keeps reserving after an overflow was raised)
                assert e.code == rx.LANDING_OVERFLOW and e.retriable and e.need_rows > ex.geometry(nparts)[1]
                before = ex.arena_bytes
                ex.reset(ex.bytes_for(nparts, e.need_rows + 8))
                assert ex.epoch == 0 and ex.arena_bytes > before
        assert attempts == 2
        ds = dests(1024)
        check(ex.exchange(table(), nparts, ds), ds)
        # a purely local combine -> compact publishes its row count the same way
        t = table()
        ok, ov = torch.empty(4096, dtype=torch.int64), torch.empty(4096, dtype=torch.int64)
        t.compact_into(ok, ov, clear=True)
        pub = ex.publish(t.status, 4096)
        assert pub.rows(0) == sum(len(w) for w in want.values())
    finally:
        ex.close()


@pytest.mark.parametrize("world,nparts", [(1, 3), (2, 4), (3, 5)])
def test_row_record_exchange_variant(world, nparts, monkeypatch):
    """The opt-in row-record mover (rows packed into aligned records, write-combined per
    destination, bulk stores, unpack on arrival) delivers the same rows as the column mover."""
    xlib = ctypes.CDLL(emu.build("exchange", asan=bool(os.environ.get("BS_EMU_ASAN"))))
    groups = _peer_groups(xlib, world, 4 << 20, monkeypatch)

    def cols_of(r):
        g = torch.Generator().manual_seed(r)
        n = 5000 + 13 * r
        return [torch.randint(-50, 50, (n,), generator=g), torch.randn(n, generator=g),
                torch.randn(n, 3, generator=g, dtype=torch.float64)]
    res = _threads(world, lambda r: [groups[r].exchange_rows(cols_of(r), nparts, nkeys=1) for _ in range(2)])
    for p in range(nparts):
        want = []
        for r in range(world):
            c = cols_of(r)
            sel = hashing.hash_tensor(c[0], 0) % nparts == p
            want += list(zip(c[0][sel].tolist(), c[1][sel].tolist(), map(tuple, c[2][sel].tolist())))
        got = res[p % world][1][p]
        assert sorted(zip(got[0].tolist(), got[1].tolist(), map(tuple, got[2].tolist()))) == sorted(want)


def test_table_exchange_of_the_generic_path(monkeypatch):
    """``PeerGroup.exchange_table`` (the in-place table shuffle of csrc/exchange.cu, kept as
    the library-baseline's companion): rows by shard, scan-and-clear, votes, and a clean
    exchange after failed votes — two ranks."""
    from bigslice_b200.parallel import peer

    class Both(Lib):
        def __init__(self):
            self.libs = [ctypes.CDLL(emu.build(f)) for f in ("exchange", "hash_combine")]
    xlib = Both()
    monkeypatch.setattr(gpu, "_lib", xlib)
    groups = _peer_groups(xlib, 2, 1 << 20, monkeypatch)
    nparts = 4

    def body(r):
        k, v = rows_of(r, 2)
        kinds, out = [], None
        for vote in ((1 if r == 1 else 0, 5), (0, 100 + r), (0, 9)):
            t = gpu.HashTable(16384, torch.int64, "add", "cpu", warm=0)
            t.track(nparts)
            t.combine(k, v)
            try:
                got = groups[r].exchange_table(t, nparts, clear=True, vote=vote)
                kinds.append(0)
                out = {p: (c[0].clone(), c[1].clone()) for p, c in got.items()}
                assert int((t.slots[:, 0] != EMPTY).sum()) == 0          # scanned and cleared
                # the reserved key does not travel with this kernel: it stays in the status
                # block (has_special, value) for the caller to route
                assert int(t.status[1]) == 1 and int(t.status[2]) == int(v[k == EMPTY].sum())
            except peer.ExchangeVoteError as e:
                kinds.append(e.kind)
        return kinds, out
    res = _threads(2, body)
    assert [k for k, _ in res] == [[3, 4, 0]] * 2
    # arrivals are per-source partial aggregates of each shard: combine them and compare
    want = {p: {a: b for a, b in d.items() if a != EMPTY} for p, d in reference(2, nparts).items()}
    got = {}
    for _, out in res:
        for p, (kk, vv) in out.items():
            d = got.setdefault(p, {})
            for a, b in zip(kk.tolist(), vv.tolist()):
                d[a] = d.get(a, 0) + b
    assert got == want


class _FakeGroup:
    """A two-thread stand-in for a torch.distributed group: rank = which thread asks."""

    def __init__(self, world):
        self.world = world
        self.ranks = {}
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world

    def rank(self):
        return self.ranks[threading.get_ident()]


def test_reduce_exchange_host_objects_two_ranks(monkeypatch):
    """Two ``ReduceExchange`` objects (one per thread) find each other through the IPC
    entry points, run graph-less exchanges over several epochs, fail together when only
    rank 1's result buffers are too small, and grow their arenas with the collective reset."""
    import torch.distributed as dist

    class Lib3(Lib):
        def __init__(self):
            self.libs = [ctypes.CDLL(emu.build(f)) for f in ("reduce_exchange", "hash_combine", "exchange")]
    lib = Lib3()
    monkeypatch.setattr(gpu, "require", lambda: lib)
    monkeypatch.setattr(gpu, "_lib", lib)
    monkeypatch.setattr(gpu, "_stream", lambda: c_void_p(0))
    monkeypatch.setattr(gpu, "_pool", gpu.TablePool())
    monkeypatch.setenv("BIGSLICE_NO_GRAPHS", "1")
    for name, fake in (("Stream", _FakeStream), ("Event", _FakeEvent), ("stream", _Null)):
        monkeypatch.setattr(torch.cuda, name, fake)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    grp = _FakeGroup(2)
    monkeypatch.setattr(dist, "get_world_size", lambda g=None: g.world)
    monkeypatch.setattr(dist, "get_rank", lambda g=None: g.rank())

    def all_gather_object(out, obj, group=None):
        group.slots[group.rank()] = obj
        group.barrier.wait()
        out[:] = list(group.slots)
        group.barrier.wait()
    monkeypatch.setattr(dist, "all_gather_object", all_gather_object)
    monkeypatch.setattr(dist, "barrier", lambda group=None: group.barrier.wait())
    nparts = 4

    def dests(ex, cap_rows):
        pl, _ = ex.geometry(nparts)
        return [(j, gpu.HashTable(gpu.table_capacity(cap_rows), torch.int64, "add", "cpu", warm=0),
                 torch.empty(cap_rows, dtype=torch.int64), torch.empty(cap_rows, dtype=torch.int64), cap_rows)
                for j in range(pl) if ex.rank + j * ex.world < nparts]

    def body(r):
        grp.ranks[threading.get_ident()] = r
        grp.barrier.wait()
        ex = rx.ReduceExchange(grp, 1 << 20, "cpu", timeout_s=20.0)
        try:
            assert (ex.world, ex.rank) == (2, r)
            outs = []
            for seed in range(3):
                k, v = rows_of(r, 2, seed=seed)
                t = gpu.HashTable(16384, torch.int64, "add", "cpu", warm=0)
                t.combine(k, v)
                ds = dests(ex, 2048)
                pend = ex.exchange(t, nparts, ds, vote=(0, 3))
                outs.append({ex.rank + j * ex.world: dict(zip(ok[:pend.rows(i)].tolist(), ov[:pend.rows(i)].tolist()))
                             for i, (j, dt, ok, ov, cap) in enumerate(ds)})
            # only rank 1's buffers are too small: both ranks fail, together
            k, v = rows_of(r, 2)
            t = gpu.HashTable(16384, torch.int64, "add", "cpu", warm=0)
            t.combine(k, v)
            pend = ex.exchange(t, nparts, dests(ex, 2048 if r == 0 else 32))
            try:
                pend.rows(0)
                verdict = 0
            except rx.ExchangeFailed as e:
                verdict = e.code
            # collective reset with a larger arena, then a clean exchange at epoch 1 again
            ex.reset(4 << 20)
            t = gpu.HashTable(16384, torch.int64, "add", "cpu", warm=0)
            t.combine(k, v)
            ds = dests(ex, 2048)
            pend = ex.exchange(t, nparts, ds)
            last = {ex.rank + j * ex.world: dict(zip(ok[:pend.rows(i)].tolist(), ov[:pend.rows(i)].tolist()))
                    for i, (j, dt, ok, ov, cap) in enumerate(ds)}
            return outs, verdict, last, ex.arena_bytes, ex.epoch
        finally:
            grp.barrier.wait()
            ex.close()
    res = _threads(2, body)
    for seed in range(3):
        got = {}
        for outs, *_ in res:
            got.update(outs[seed])
        assert got == reference(2, nparts, seed=seed)
    assert [r[1] for r in res] == [rx.DEST_OVERFLOW] * 2
    last = {}
    for r in res:
        last.update(r[2])
    assert last == reference(2, nparts)
    assert all(r[3] == 4 << 20 and r[4] == 1 for r in res)


def test_exchange_random_shapes(lib):
    """Random worlds, shard counts (fewer shards than ranks, odd counts), table regions,
    ops, value types and row counts (including ranks with no rows at all)."""
    import random
    rnd = random.Random(99)
    for trial in range(12):
        world = rnd.choice([1, 2, 3, 4])
        regions = rnd.choice([1, 1, 2, 4])
        nparts = rnd.choice([1, 2, 3, 5, 8])
        if nparts * regions > 64:
            continue
        op = rnd.choice(["add", "min", "max"])
        dt = rnd.choice([torch.int64, torch.int32, torch.float32, torch.float64])
        sizes = [rnd.choice([0, 1, 37, 2500, 7000]) for _ in range(world)]
        nkeys = rnd.choice([3, 200, 5000])

        def rows(r):
            g = torch.Generator().manual_seed(1000 * trial + r)
            n = sizes[r]
            k = torch.randint(0, nkeys, (n,), generator=g) * 1000003 - 5
            if n > 10:
                k[::7] = EMPTY
            return k, torch.randint(-9, 9, (n,), generator=g).to(dt)

        def body(r):
            k, v = rows(r.rank)
            return r.exchange(k, v, nparts, op=op, regions=regions, dest_cap=16384)
        res = run_ranks(world, body, arena_bytes=4 << 20, lib=lib)
        ks, vs = zip(*[rows(r) for r in range(world)])
        k, v = torch.cat(ks), torch.cat(vs)
        want = {}
        if k.numel():
            uk, inv = torch.unique(k, return_inverse=True)
            red = {"add": "sum", "min": "amin", "max": "amax"}[op]
            acc = torch.zeros(uk.numel(), dtype=dt).scatter_reduce_(0, inv, v, red, include_self=False)
            dest = hashing.hash_tensor(uk, 0) % nparts
            want = {(int(p), a): b for p, a, b in zip(dest.tolist(), uk.tolist(), acc.tolist())}
        got = {}
        for r, (rep, out) in enumerate(res):
            assert rep.error == 0, (trial, world, nparts, regions, rep.error)
            for p, (kk, vv) in out.items():
                assert p % world == r
                for a, b in zip(kk.tolist(), vv.tolist()):
                    got[(p, a)] = b
        assert got == want, (trial, world, nparts, regions, op, dt, sizes)


def test_replicator_peer_store_broadcast(monkeypatch):
    """``parallel/broadcast.Replicator`` in its peer-store mode (the fallback of the NVSwitch
    multicast path, same kernel shape): every rank's rows land in every rank's buffer at the
    right offsets, for several columns of different widths."""
    import torch.distributed as dist
    from bigslice_b200.parallel.broadcast import Replicator
    xlib = ctypes.CDLL(emu.build("cogroup"))
    monkeypatch.setattr(gpu, "require", lambda: xlib)
    monkeypatch.setattr(gpu, "_stream", lambda: c_void_p(0))
    world = 3
    grp = _FakeGroup(world)

    def all_gather_object(out, obj, group=None):
        group.slots[group.rank()] = obj
        group.barrier.wait()
        out[:] = list(group.slots)
        group.barrier.wait()
    monkeypatch.setattr(dist, "all_gather_object", all_gather_object)

    class Handle:
        def barrier(self):
            grp.barrier.wait()
    nbytes = 1 << 20
    bufs = [torch.zeros(nbytes, dtype=torch.uint8) for _ in range(world)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64)

    def cols_of(r):
        g = torch.Generator().manual_seed(r)
        n = [400, 0, 1002][r]
        return [torch.randint(-9, 9, (n,), generator=g), torch.randn(n, 2, generator=g, dtype=torch.float64),
                torch.randn(n, 4, generator=g)]

    def body(r):
        grp.ranks[threading.get_ident()] = r
        grp.barrier.wait()
        rep = object.__new__(Replicator)
        rep.group = rep.ctrl_group = grp
        rep.world, rep.rank, rep.device, rep.nbytes = world, r, torch.device("cpu"), nbytes
        rep.mode, rep.buf, rep.handle, rep.mc_ptr, rep.peer_ptrs, rep.why = "peer", bufs[r], Handle(), 0, ptrs, ""
        return [rep.all_gather(cols_of(r)) for _ in range(2)][-1]
    res = _threads(world, body)
    want = [torch.cat([cols_of(r)[j] for r in range(world)]) for j in range(3)]
    for got in res:
        assert all(torch.equal(a, b) for a, b in zip(got, want))


def test_long_sequences_with_failures_of_every_kind(lib):
    """Fifteen epochs on three ranks where a random epoch fails in a random way (vote
    mismatch, failed producer, destination overflow on one rank, landing overflow): every
    rank reports the same verdict every time, and the epochs around a failure are exact —
    parity halves, acknowledgements, closing flags and tables stay consistent."""
    import random
    rnd = random.Random(7)
    world, nparts, epochs = 3, 5, 15
    plan = []
    for e in range(epochs):
        kind = rnd.choice(["ok", "ok", "ok", "fp", "producer", "dest", "landing"])
        plan.append((kind, rnd.randrange(world)))

    def rows(r, e, big):
        g = torch.Generator().manual_seed(100 * e + r)
        n = 60000 if big else 3000
        k = torch.randint(0, 40000 if big else 700, (n,), generator=g) * 613 + 1
        return k, torch.ones(n, dtype=torch.int64)

    def body(r):
        out = []
        for e, (kind, who) in enumerate(plan):
            k, v = rows(r.rank, e, kind == "landing")
            vote = (0, 11)
            kw = {}
            if kind == "fp" and r.rank == who:
                vote = (0, 12)
            if kind == "producer" and r.rank == who:
                vote = (1, 11)
            if kind == "dest" and r.rank == who:
                kw["out_cap"] = 8
            rep, got = r.exchange(k, v, nparts, vote=vote, dest_cap=65536 if kind == "landing" else 8192, **kw)
            out.append((rep.error, {p: dict(zip(a.tolist(), b.tolist())) for p, (a, b) in got.items()}
                        if rep.error == 0 else None))
        return out
    # 1 MiB arena, 5 shards over 3 ranks: 2 local shards -> 5461 rows per region: the "landing" epochs overflow
    res = run_ranks(world, body, arena_bytes=1 << 20, lib=lib)
    expect = {"ok": 0, "fp": rx.FINGERPRINT, "producer": rx.PRODUCER_FAILED, "dest": rx.DEST_OVERFLOW,
              "landing": rx.LANDING_OVERFLOW}
    for e, (kind, who) in enumerate(plan):
        verdicts = [res[r][e][0] for r in range(world)]
        assert verdicts == [expect[kind]] * world, (e, kind, who, verdicts)
        if kind == "ok":
            ks, vs = zip(*[rows(r, e, False) for r in range(world)])
            uk, cnt = torch.unique(torch.cat(ks), return_counts=True)
            dest = hashing.hash_tensor(uk, 0) % nparts
            want = {p: dict(zip(uk[dest == p].tolist(), cnt[dest == p].tolist())) for p in range(nparts)}
            got = {}
            for r in range(world):
                got.update(res[r][e][1])
            assert got == want, e
    assert {k for k, _ in plan} >= {"ok", "fp", "producer", "dest", "landing"}
