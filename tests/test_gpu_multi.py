"""Multi-GPU tests proper (NCCL, one rank per GPU): skipped on a box with fewer than 2 GPUs.  They launch
torchrun on 2 GPUs and check the FUSED exchange -- the accept kernel's peer stores into every rank's
symmetric-memory replica, the release signal in its last CTA and the acquire wait at the head of the next
proposal kernel -- and the NCCL all_gather fallback against a single-rank replay, bit for bit."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def torchrun(script, *args, nproc=2, port=29761, env=None, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), script] + list(args)
    e = dict(os.environ, OMP_NUM_THREADS="1")
    e.update(env or {})
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e)


@pytest.mark.skipif(ngpus() < 2, reason="needs 2 GPUs")
def test_fused_nvlink_exchange_is_bitwise_the_single_rank_run():
    r = torchrun(os.path.join(ROOT, "tests", "nccl_check.py"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("bitwise equal to the single-process run") == 4      # 2 ranks x (symm, nccl)


@pytest.mark.skipif(ngpus() < 2, reason="needs 2 GPUs")
def test_multimem_store_branch():
    """COSMOSLIK_B200_MULTICAST=1: positions are published with one multimem.st through the NVSwitch multicast
    address instead of one store per peer; same bits (falls back to peer stores where the fabric has no multicast)"""
    r = torchrun(os.path.join(ROOT, "tests", "nccl_check.py"), port=29762, env={"COSMOSLIK_B200_MULTICAST": "1"})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MISMATCH" not in r.stdout


@pytest.mark.skipif(ngpus() < 2, reason="needs 2 GPUs")
def test_bench_line_on_two_gpus():
    """the bench contract at N = 2: one JSON line, literal + weak records, exchange parity, per-rank times"""
    import json
    r = torchrun(os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "4", "--warmup", "3", "--walkers", "8192",
                 "--secondary", "gauss30_mh", "--no-cpu-baseline", port=29763)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and d["exchange_parity"] == "bitwise"
    assert len(d["ms_per_rank"]) == 2 and d["weak"]["walkers_total"] == 2 * 8192
    assert d["config"]["walkers_total"] == 8192 and d["secondary"][0]["name"] == "gauss30_mh"
    assert "error" not in d["secondary"][0] and d["secondary"][0]["config"]["adaptations"] > 0
