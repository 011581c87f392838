"""N > 1 path: candidates / particles sharded by contiguous index range, map
replicated, per-rank bests and integer weight sums all-gathered.

* CPU (`-m "not gpu"`): two gloo ranks check the host-side logic against the unsharded oracle.
* GPU (`-m gpu`, needs >= 2 devices): two ranks, one GPU each, through the C ABI's own NCCL
  communicator; winner, hit counts, weights and resampled ancestors must equal the oracle's.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mr_worker.py")


def _launch(mode: str, nproc: int, port: int, timeout: int = 600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, "--mode", mode]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert p.returncode == 0, f"{mode} worker failed:\n{p.stdout[-3000:]}\n{p.stderr[-3000:]}"
    assert f"multirank {mode} ok on {nproc} ranks" in p.stdout


def test_two_rank_host_logic_gloo(b200slam, oracle):
    _launch("cpu", 2, 29533)


def test_three_rank_host_logic_gloo(b200slam, oracle):
    _launch("cpu", 3, 29534)


@pytest.mark.gpu
def test_two_gpu_allgather_matches_oracle(b200slam, oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    _launch("gpu", 2, 29535)
