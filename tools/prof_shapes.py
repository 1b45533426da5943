"""Micro-driver for ncu: launches each hot kernel at its bs-256 dedeit shape a few times.
  ncu --set full -k regex:gemm_kernel ... python tools/prof_shapes.py gemm
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import _lib as L  # noqa: E402

M, D = 256 * 198, 384
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*s, dt=torch.bfloat16, scale=1.0):
    return (torch.randn(*s, device=dev, generator=g) * scale).to(dt)


y = rnd(M, D)
x = rnd(M, D, dt=torch.float32)
if which in ("gemm", "all"):
    w_qkv, b_qkv = rnd(1152, D, scale=.05), rnd(1152, dt=torch.float32)
    w_proj, b_proj = rnd(D, D, scale=.05), rnd(D, dt=torch.float32)
    w1, b1 = rnd(1536, D, scale=.05), rnd(1536, dt=torch.float32)
    w2, b2 = rnd(D, 1536, scale=.05), rnd(D, dt=torch.float32)
    qkv = torch.empty(M, 1152, device=dev, dtype=torch.bfloat16)
    hid = torch.empty(M, 1536, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        L.gemm(y, w_qkv, bias=b_qkv, out=qkv, out_kind=L.OUT_BF16, tag=2)
        L.gemm(y, w_proj, bias=b_proj, resid=x, out=x, out_kind=L.OUT_F32, tag=3)
        L.gemm(y, w1, bias=b1, act=L.ACT_GELU_ERF, out=hid, out_kind=L.OUT_BF16, tag=4)
        L.gemm(hid, w2, bias=b2, resid=x, out=x, out_kind=L.OUT_F32, tag=5)
if which in ("attn", "all"):
    qkv = rnd(M, 1152)
    for _ in range(reps):
        L.attention(qkv, 256, 198, 6, 0.125)
if which in ("ln", "all"):
    gam, bet = rnd(D, dt=torch.float32), rnd(D, dt=torch.float32)
    for _ in range(reps):
        L.layernorm(x, gam, bet, 1e-6, L.OUT_BF16)
torch.cuda.synchronize()
print("done")
