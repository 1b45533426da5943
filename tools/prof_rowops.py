"""HBM-bound kernels at the headline shapes (bs 256): CUDA-event time per launch, algorithmic
bytes and GB/s against the measured copy bandwidth in MEASURED_PEAKS.json.
    python tools/prof_rowops.py            # timing table (two buffer sets alternate: > 2 x L2)
    ncu --set full -k regex:'im2col|token_init|rowstats|ln_kernel|gather_ln|eval_' ... python tools/prof_rowops.py --once
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from devit_b200 import _lib as L  # noqa: E402
from devit_b200 import synth  # noqa: E402
from devit_b200.models import IMAGENET_DEFAULT_MEAN as MEAN, IMAGENET_DEFAULT_STD as STD  # noqa: E402

once = '--once' in sys.argv
B, T, D = 256, 198, 384
M = B * T
peak = 6550.1
try:
    pk = json.loads((ROOT / 'MEASURED_PEAKS.json').read_text())
    peak = float(pk.get('hbm_gbs', peak))
except Exception:  # noqa: BLE001
    pass
dev = 'cuda'
sets = 1 if once else 2
img = [torch.randn(B, 3, 224, 224, device=dev) for _ in range(sets)]
u8 = [synth.images_u8(B).to(dev) for _ in range(sets)]
u8h = [t.permute(0, 2, 3, 1).contiguous() for t in u8]
x = [torch.randn(M, D, device=dev) for _ in range(sets)]
g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
prefix, pos, bias = torch.randn(2, D, device=dev), torch.randn(T, D, device=dev), torch.randn(D, device=dev)
logits = torch.randn(B, 100, device=dev)
target = torch.randint(0, 100, (B,), device=dev)
acc = torch.zeros(5, device=dev, dtype=torch.float64)
lib = L.load()


def token_init(i):
    L.check(lib.devit_token_init(x[i].data_ptr(), prefix.data_ptr(), pos.data_ptr(),
                                 bias.data_ptr(), B, T, D, 2, L.stream_ptr()))


MB = 1e6
cases = [
    ('im2col_tokens fp32 -> bf16', lambda i: L.im2col_tokens(img[i], 2), (B * 3 * 224 * 224 * 4 + M * 768 * 2) / MB),
    ('im2col_tokens_u8 NCHW -> bf16', lambda i: L.im2col_tokens_u8(u8[i], MEAN, STD, 2), (B * 3 * 224 * 224 + M * 768 * 2) / MB),
    ('im2col_tokens_u8 NHWC -> bf16', lambda i: L.im2col_tokens_u8(u8h[i], MEAN, STD, 2, layout=L.LAYOUT_NHWC), (B * 3 * 224 * 224 + M * 768 * 2) / MB),
    ('token_init', token_init, M * D * 4 / MB),
    ('rowstats (bf16 copy + row sums)', lambda i: L.rowstats(x[i]), (M * D * 6 + M * 8) / MB),
    ('layernorm fp32 -> bf16', lambda i: L.layernorm(x[i], g, b, 1e-6), M * D * 6 / MB),
    ('eval_tail (2 launches)', lambda i: L.eval_tail(logits, target, acc), (B * 100 * 4 + B * 8) / MB),
]
print(f"peak (measured copy bandwidth): {peak:.1f} GB/s")
print(f"{'kernel':36s} {'MB':>8s} {'us':>8s} {'GB/s':>8s} {'% peak':>7s}")
for name, fn, mb in cases:
    n = 1 if once else 20
    for i in range(0 if once else 3):
        fn(i % sets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i % sets)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"{name:36s} {mb:8.1f} {us:8.1f} {mb / us * 1e3:8.0f} {100 * mb / us * 1e3 / peak:6.1f}%")
