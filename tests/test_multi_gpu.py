"""GPU tier, two ranks (skips on a one-GPU box): the fused all-gather / in-kernel barrier / NVLink-multicast stores, ragged and empty
shards, a non-register model (planar push), CUDA-graph replay of fused steps, the sharded gradient bundle, the sharded rocket step
and the host-facing sharded sweep — each bit-identical to the single-GPU result (tests/multi_gpu_worker.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_rank_paths_are_bitwise_equal_to_one_gpu(built):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`)")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29655",
           os.path.join(here, "multi_gpu_worker.py")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(res.stdout[-6000:])
    assert res.returncode == 0
