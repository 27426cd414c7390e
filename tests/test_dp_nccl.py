"""On-hardware data-parallel equivalence (VERDICT r1 parity gap 3): 2 NCCL ranks x B/2 reproduce the 1-GPU flat
gradient buffers and post-step weights of OTTrainStep.  Needs two GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_match_single_process(cuda_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dp_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "DP_NCCL_OK" in r.stdout, r.stderr[-3000:]
