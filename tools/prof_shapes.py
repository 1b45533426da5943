"""Micro-driver for ncu: launches each hot kernel at its bs-256 dedeit shape (shrunk widths:
4 kept heads, 928 kept neurons; LayerNorm-folded epilogues as the forward uses them).
  ncu --set full -k regex:gemm_kernel ... python tools/prof_shapes.py gemm
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import _lib as L  # noqa: E402

M, D, H, F = 256 * 198, 384, 4, 928
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*s, dt=torch.bfloat16, scale=1.0):
    return (torch.randn(*s, device=dev, generator=g) * scale).to(dt)


x = rnd(M, D, dt=torch.float32)
xb, stats1 = L.rowstats(x)
if which in ("gemm", "all"):
    w_qkv, c1q, c2q = rnd(192 * H, D, scale=.05), rnd(192 * H, dt=torch.float32), rnd(192 * H, dt=torch.float32)
    w_proj, b_proj = rnd(D, 64 * H, scale=.05), rnd(D, dt=torch.float32)
    w1, c11, c21 = rnd(F, D, scale=.05), rnd(F, dt=torch.float32), rnd(F, dt=torch.float32)
    w2, b2 = rnd(D, F, scale=.05), rnd(D, dt=torch.float32)
    qkv = torch.empty(M, 192 * H, device=dev, dtype=torch.bfloat16)
    o = rnd(M, 64 * H)
    hid = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
    stats = torch.empty(6, M, 2, device=dev)
    for _ in range(reps):
        L.gemm(xb, w_qkv, bias=c2q, out=qkv, out_kind=L.OUT_BF16, tag=2, ln_stats=stats1,
               ln_colsum=c1q, ln_dim=D, ln_eps=1e-6)
        L.gemm(o, w_proj, bias=b_proj, resid=x, out=x, out_kind=L.OUT_F32, tag=3, out_bf16=xb,
               stats_out=stats)
        L.gemm(xb, w1, bias=c21, act=L.ACT_GELU_ERF, out=hid, out_kind=L.OUT_BF16, tag=4,
               ln_stats=stats, ln_colsum=c11, ln_dim=D, ln_eps=1e-6)
        L.gemm(hid, w2, bias=b2, resid=x, out=x, out_kind=L.OUT_F32, tag=5, out_bf16=xb,
               stats_out=stats)
if which in ("mlp", "all"):
    for F_ in (F, 1536):
        w1, c11, c21 = rnd(F_, D, scale=.05), rnd(F_, dt=torch.float32), rnd(F_, dt=torch.float32)
        w2, b2 = rnd(D, F_, scale=.05), rnd(D, dt=torch.float32)
        stats6 = torch.zeros(6, M, 2, device=dev)
        stats6[0] = stats1[0]
        so = torch.empty(4, M, 2, device=dev)
        for _ in range(reps):
            L.mlp_fused(x, xb, stats6, w1, c11, c21, w2, b2, 1e-6, xb_out=xb, stats_out=so)
if which in ("projmlp", "all"):
    # the projection fused in front of the MLP (devit_mlp_args.o): the per-layer tail kernel
    for F_ in (F, 1536):
        w1, c11, c21 = rnd(F_, D, scale=.05), rnd(F_, dt=torch.float32), rnd(F_, dt=torch.float32)
        w2, b2 = rnd(D, F_, scale=.05), rnd(D, dt=torch.float32)
        hh = H if F_ == F else 6
        o, w_proj, b_proj = rnd(M, 64 * hh), rnd(D, 64 * hh, scale=.05), rnd(D, dt=torch.float32)
        so = torch.empty(4, M, 2, device=dev)
        lo = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
        for _ in range(reps):
            # the steady state inside a model: residual stream as hi (xb) + lo planes in and out
            L.mlp_fused(x, xb, None, w1, c11, c21, w2, b2, 1e-6, xb_out=xb, stats_out=so,
                        o=o, w_proj=w_proj, b_proj=b_proj, x_lo_in=lo, x_lo_out=lo)
if which in ("attn", "all"):
    qkv = rnd(M, 192 * H)
    for _ in range(reps):
        L.attention(qkv, 256, 198, H, 0.125)
if which in ("rows", "all"):
    gam, bet = rnd(D, dt=torch.float32), rnd(D, dt=torch.float32)
    for _ in range(reps):
        L.rowstats(x)
        L.layernorm(x, gam, bet, 1e-6, L.OUT_BF16)
torch.cuda.synchronize()
print("done")
