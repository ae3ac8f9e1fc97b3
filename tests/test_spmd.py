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
    assert r.returncode == 0, r.stderr[-4000:]
    return [json.load(open(out + ".%d" % i)) for i in range(nproc)]


def check(outs, world):
    scanned = sorted(v for o in outs for v in o["scan_local"])
    assert scanned == [2 * i for i in range(100)]
    for o in outs:
        assert o["world"] == world
        for k in ("wordcount_ok", "chain_ok", "strcount_ok", "join_ok", "reshard_ok",
                  "reshard_placement_ok", "failure_ok"):
            assert o[k], (k, o)
    for o in outs:
        for k in ("exchange_content_ok", "table_exchange_ok", "exchange_vote_ok", "bigcount_ok",
                  "sentinel_ok", "distinct_ok", "hints_ok", "overflow_recovered_ok", "recovery_ok"):
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
        assert set(outs[0]["modes"]) == {"gloo", "object"}


@pytest.mark.gpu
def test_spmd_single_gpu_fused_self_exchange(tmp_path):
    outs = launch(1, tmp_path, device="cuda", shuffle="fused")
    check(outs, 1)
    assert "fused" in outs[0]["modes"] and outs[0]["launches"] > 0
    assert outs[0]["exchange_content_ok"] and outs[0]["bigcount_ok"] and outs[0]["table_exchange_ok"]


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
