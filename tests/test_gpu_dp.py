"""Multi-GPU data-parallel parity (needs >= 2 GPUs; skipped otherwise): tools/dp_parity.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("overlap", ("0", "1"))
def test_two_rank_nccl_matches_single_process(overlap):
    """overlap = 1: the world-model gradient all-reduce in two buckets, the head networks' bucket running next to the
    recurrent weight gradients + encoder backward (dv3_dp_piece 5 / 6); 0: one collective after the backward pass."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "dp_parity.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, DV3_DP_OVERLAP=overlap))
    assert out.returncode == 0 and "DP PARITY OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
