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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_model_on_a_device_that_is_not_current():
    """the C side launches on the CURRENT device; a model / input on cuda:1 while
    cuda:0 is current must still run on cuda:1's stream (the wrappers make the operand's device
    current) and give the cuda:0 result bit for bit."""
    from devit_b200 import models, synth  # noqa: F401  (models registers the entrypoints)
    from devit_b200.registry import create_model
    torch.cuda.set_device(0)
    x = synth.images(3)
    outs = []
    for d in (0, 1):
        m = create_model('dedeit', num_classes=25)
        m.load_state_dict(synth.dedeit_state_dict(0, num_classes=25))
        m = m.to(f'cuda:{d}').eval().set_precision('bf16')
        assert torch.cuda.current_device() == 0
        outs.append(m(x.to(f'cuda:{d}')).cpu())
        # layer-wise module path (plain wrappers) on the same device
        blk = m.blocks[0]
        y = blk(torch.randn(2, 198, 384, device=f'cuda:{d}'))['output']
        assert y.device.index == d
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    assert torch.equal(outs[0], outs[1])
