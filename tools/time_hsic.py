"""HSIC neuron ranking of ONE layer on the GPU: the batched evaluation of devit_b200/shrink.py
against the reference's per-unit loop (core/imp_rank.py:30-34; timed on 48 units through the
oracle restatement, which issues the same torch ops + .item() per unit, and extrapolated).
  gpurun -- 'python tools/time_hsic.py > gpurun_out/time_hsic.txt'"""
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import shrink  # noqa: E402
from oracle import hsic_oracle as HO  # noqa: E402

B, N, FEAT, C = 64, 198, 1536, 100
g = torch.Generator(device='cuda').manual_seed(0)
no = F.gelu(torch.randn(B, N, FEAT, device='cuda', generator=g))
logits = torch.randn(B, C, device='cuda', generator=g)
prob = F.softmax(logits, -1)
shrink.neuron_scores(no, logits)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    shrink.neuron_scores(no, logits)
torch.cuda.synchronize()
batched = (time.perf_counter() - t0) / 3
HO.hsic(no[:, :, 0], prob, 'linear', True).item()
t0 = time.perf_counter()
for f in range(48):
    HO.hsic(no[:, :, f], prob, 'linear', True).item()
loop = (time.perf_counter() - t0) / 48 * FEAT
print(f"one Mlp layer, B={B} N={N} F={FEAT}: batched {batched * 1e3:.1f} ms | per-unit loop "
      f"{loop * 1e3:.0f} ms (48 units timed, x{FEAT // 48}) | x{loop / batched:.0f}")
