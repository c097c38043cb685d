"""N > 1 on real GPUs: torchrun with one rank per GPU over NCCL (skipped on a single-GPU box;
the same logic runs on CPU with gloo in test_mc3_gloo.py / test_shard_gloo.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_two_gpus_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=800)
    assert r.returncode == 0 and "MULTIGPU OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_all_gpus_nccl():
    """the same worker on every GPU of an 8-GPU box: sharded evaluation, the split ladder and ONE chain advanced by
    eight ranks (80 CTAs: the split-reduction path), bit-identical to the single-GPU chain and in lockstep with the oracle"""
    import torch
    n = torch.cuda.device_count()
    if n < 8:
        pytest.skip("needs eight GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=8", "--master-addr", "127.0.0.1",
           "--master-port", str(29300 + os.getpid() % 300), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=800)
    assert r.returncode == 0 and "MULTIGPU OK world=8" in r.stdout and "SHARDED STEPPING OK world=8" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
