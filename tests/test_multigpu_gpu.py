"""World-size-2 NCCL run of the sharded ensemble on a box with >= 2 GPUs (skipped otherwise):
tools/check_multigpu.py compares the sharded logits with the single-GPU logits."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_sharding_matches_single_gpu():
    env = dict(os.environ, CHECK_BATCH="8")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        "29531", str(ROOT / "tools" / "check_multigpu.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "argmax-equal=True" in r.stdout
