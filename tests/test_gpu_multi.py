"""N>1 on real GPUs: sharded scan + NCCL merge equals the single-GPU result (tests/multi_gpu_check.py
under torchrun). Skipped on boxes with one GPU; the host-side N>1 logic is covered on CPU by
tests/test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_merge_equals_single_gpu(built_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(ROOT, "tests", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
