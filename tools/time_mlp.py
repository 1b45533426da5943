"""Times devit_mlp_fused at the bs-256 shape (TIME_MLP_BATCH=<images> for another batch):
   python tools/time_mlp.py [hidden ...]
TIME_MLP_PROJ=<heads>: the variant with the attention-output projection fused in front."""
import os
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import _lib as L  # noqa: E402
M, D = int(os.environ.get('TIME_MLP_BATCH', '256')) * 198, 384
H = int(os.environ.get('TIME_MLP_PROJ', '0'))
SPLIT = os.environ.get('TIME_MLP_SPLIT', '0') == '1'  # residual as bf16 hi + lo planes
g = torch.Generator(device="cuda").manual_seed(0)
for F in [int(a) for a in sys.argv[1:]] or [928, 1536]:
    x = torch.randn(M, D, device="cuda", generator=g)
    xb, stats = L.rowstats(x)
    w1 = (torch.randn(F, D, device="cuda", generator=g) * .05).bfloat16()
    w2 = (torch.randn(D, F, device="cuda", generator=g) * .05).bfloat16()
    c1, c2, b2 = (torch.randn(n, device="cuda", generator=g) * .1 for n in (F, F, D))
    so = torch.empty(4, M, 2, device="cuda")
    if H:
        o = torch.randn(M, 64 * H, device="cuda", generator=g).bfloat16()
        wp = (torch.randn(D, 64 * H, device="cuda", generator=g) * .05).bfloat16()
        bp = torch.randn(D, device="cuda", generator=g) * .1
        lo = torch.zeros(M, D, device="cuda", dtype=torch.bfloat16)
    def run():
        if H and SPLIT:
            L.mlp_fused(x, xb, None, w1, c1, c2, w2, b2, 1e-6, xb_out=xb, stats_out=so,
                        o=o, w_proj=wp, b_proj=bp, x_lo_in=lo, x_lo_out=lo)
        elif H:
            L.mlp_fused(x, None, None, w1, c1, c2, w2, b2, 1e-6, xb_out=xb, stats_out=so,
                        o=o, w_proj=wp, b_proj=bp)
        else:
            L.mlp_fused(x, xb, stats, w1, c1, c2, w2, b2, 1e-6, xb_out=xb, stats_out=so)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 4.0 * M * D * F + 2.0 * M * D * 64 * H
    print(f"F={F} proj_heads={H} split={int(SPLIT)}: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s")
