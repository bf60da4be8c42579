"""Two GPUs: the decomposed march over the real NCCL transports is bit-identical to the undecomposed one.  Skipped on
a one-GPU box (there ``bench.py --gpus N`` runs the same check as its multi-rank parity leg before timing)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_rank_march_over_nccl_is_bit_identical(gpu):
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(HERE, "nccl_worker.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "'bit_equal': True" in res.stdout
