"""Multi-process sharding over >= 2 real GPUs (skipped on a 1-GPU box): launches tests/mp_shard_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_ipc_remap_and_reductions():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tests", "mp_shard_check.py"), "16"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-3000:])
    assert "MP_SHARD_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
