"""Multi-process SPMD executor tests.  CPU: gloo, world_size 2 (BASELINE config
1: Const->Map->Scan plumbing) and every shuffle-bearing combinator.  GPU: the
same worker with the fused NVLink exchange (1 GPU self-exchange; 2 GPUs when
available) and the NCCL baseline path."""
import json
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(nproc, tmp_path, device="cpu", shuffle="auto", timeout=600, extra_env=None):
    out = str(tmp_path / "out.json")
    env = dict(os.environ, BS_TEST_OUT=out, BS_TEST_DEVICE=device, BS_TEST_SHUFFLE=shuffle)
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node=%d" % nproc, "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(HERE, "spmd_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    dump = os.environ.get("BS_TEST_DUMP")   # keep what the workers said (debugging on a remote GPU box)
    if dump:
        os.makedirs(dump, exist_ok=True)
        tag = "%s_w%d_%s" % (device, nproc, os.path.basename(str(tmp_path)))
        with open(os.path.join(dump, tag + ".stderr"), "w") as fh:
            fh.write(r.stderr[-20000:])
        for i in range(nproc):
            if os.path.exists(out + ".%d" % i):
                with open(out + ".%d" % i) as src, open(os.path.join(dump, "%s.%d.json" % (tag, i)), "w") as dst:
                    o = json.load(src)
                    o.pop("scan_local", None)
                    o.pop("wordcount_local_keys", None)
                    json.dump(o, dst)
    assert r.returncode == 0, r.stderr[-4000:]
    return [json.load(open(out + ".%d" % i)) for i in range(nproc)]


def check(outs, world):
    scanned = sorted(v for o in outs for v in o["scan_local"])
    assert scanned == [2 * i for i in range(100)]
    for o in outs:
        assert o["world"] == world
        for k in ("wordcount_ok", "chain_ok", "strcount_ok", "textcount_ok", "join_ok", "bjoin_ok", "pairsum_ok",
                  "bytemax_ok",
                  "reshard_ok",
                  "reshard_placement_ok", "failure_ok"):
            assert o[k], (k, o)
    for o in outs:
        for k in ("exchange_content_ok", "table_exchange_ok", "exchange_vote_ok", "bigcount_ok",
                  "sentinel_ok", "distinct_ok", "hints_ok", "overflow_recovered_ok", "dest_overflow_recovered_ok",
                  "recovery_ok"):
            if k in o:
                assert o[k], (k, o["rank"])
    if world > 1:
        allk = [set(o["wordcount_local_keys"]) for o in outs]
        assert not (allk[0] & allk[1])  # each key lives on exactly one rank


@pytest.mark.parametrize("world", [1, 2])
def test_spmd_cpu_gloo(world, tmp_path):
    outs = launch(world, tmp_path)
    check(outs, world)
    if world > 1:
        # tensor frames ride the gloo all-to-all, string/list frames the object path
        assert set(outs[0]["modes"]) == {"gloo", "object", "replicate-library"}


@pytest.mark.gpu
def test_spmd_single_gpu_fused_self_exchange(tmp_path):
    outs = launch(1, tmp_path, device="cuda", shuffle="fused")
    check(outs, 1)
    assert "fused" in outs[0]["modes"] and outs[0]["launches"] > 0
    assert outs[0]["exchange_content_ok"] and outs[0]["bigcount_ok"] and outs[0]["table_exchange_ok"]
    # packed two-column / narrow keys take the device table path too
    assert outs[0]["pairsum_mode"] in ("fused-table", "table-local") and outs[0]["bytemax_mode"] == "fused-table"
    # string keys: tokenised and hashed by the string kernels, counted on the table path
    assert outs[0]["textcount_mode"] == "fused-table"


@pytest.mark.gpu
@pytest.mark.multigpu
@pytest.mark.parametrize("shuffle", ["fused", "nccl"])
def test_spmd_two_gpus(shuffle, tmp_path):
    outs = launch(2, tmp_path, device="cuda", shuffle=shuffle)
    check(outs, 2)
    assert shuffle in outs[0]["modes"]


@pytest.mark.gpu
def test_spmd_two_ranks_sharing_one_gpu(tmp_path):
    """Two ranks on ONE device (gloo bootstrap, CUDA IPC between the processes):
    the whole multi-rank protocol — peer stores, flags, acknowledgements, the
    shared-memory vote, loss recovery — is exercised on a single-GPU box."""
    outs = launch(2, tmp_path, device="cuda", shuffle="fused", timeout=900,
                  extra_env={"BIGSLICE_DIST_BACKEND": "gloo", "BIGSLICE_SHARE_GPU": "1"})
    check(outs, 2)
    assert "fused-table" in outs[0]["modes"]
    assert outs[0]["sentinel_ok"] and outs[0]["distinct_ok"] and outs[0]["overflow_recovered_ok"]


def _spawn_ranks(script, world, tmp_path, env_extra, timeout=120):
    """Plain subprocesses (not torchrun: its agent would kill the survivors as
    soon as one rank dies)."""
    out = str(tmp_path / "out.json")
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), BS_TEST_OUT=out)
        env.update(env_extra)
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, script)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    res = []
    for p in procs:
        try:
            o, e = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            p.kill()
            o, e = p.communicate()
            e += "\n[killed: timeout]"
        res.append((p.returncode, o, e))
    return out, res


def _check_rank_death(out, res):
    assert res[1][0] == 17, res[1][2][-2000:]          # rank 1 died as scripted
    assert os.path.exists(out + ".0"), res[0][2][-3000:]
    o = json.load(open(out + ".0"))
    assert o["ok_first"]
    assert o["error"] is not None and o["lost_classified"], o
    assert o["seconds"] < 35, o


def test_rank_death_is_a_lost_task_not_a_hang_cpu(tmp_path):
    out, res = _spawn_ranks("kill_worker.py", 2, tmp_path, {"BS_TEST_DEVICE": "cpu", "BS_TEST_TIMEOUT": "3"})
    _check_rank_death(out, res)


@pytest.mark.gpu
def test_rank_death_is_a_lost_task_not_a_hang_gpu(tmp_path):
    out, res = _spawn_ranks("kill_worker.py", 2, tmp_path,
                            {"BS_TEST_DEVICE": "cuda", "BS_TEST_TIMEOUT": "3", "BIGSLICE_DIST_BACKEND": "gloo",
                             "BIGSLICE_SHARE_GPU": "1", "LOCAL_RANK": "0"}, timeout=180)
    _check_rank_death(out, res)
