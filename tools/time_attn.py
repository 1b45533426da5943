"""Times devit_attention at the bs-256 shapes: python tools/time_attn.py"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import _lib as L  # noqa: E402
B, N = 256, 198
for heads in (3, 4, 5, 6):
    qkv = torch.randn(B * N, 3 * heads * 64, device="cuda").bfloat16()
    for _ in range(3):
        L.attention(qkv, B, N, heads, 0.125)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        L.attention(qkv, B, N, heads, 0.125)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 4.0 * heads * N * N * 64 * B
    print(f"heads={heads}: {us:7.1f} us  {fl / us / 1e6:6.1f} TFLOP/s  "
          f"{(B * N * heads * 64 * 2 * 4) / us / 1e3:6.1f} GB/s algorithmic")
