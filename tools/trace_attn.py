"""clock64 trace of the persistent attention kernel, CTA 0 (needs the -DDEVIT_GEMM_TRACE build):
   DEVIT_B200_LIB=devit_b200/lib/libdevit_b200_trace.so python tools/trace_attn.py [heads]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import _lib as L  # noqa: E402
heads = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B, N = 256, 198
qkv = torch.randn(B * N, 3 * heads * 64, device="cuda").bfloat16()
for _ in range(3):
    L.attention(qkv, B, N, heads, 0.125)
torch.cuda.synchronize()
buf = torch.zeros(20, 512, device="cuda", dtype=torch.int64)
L.load().devit_debug_set_trace(buf.data_ptr())
L.attention(qkv, B, N, heads, 0.125)
torch.cuda.synchronize()
L.load().devit_debug_set_trace(None)
t = buf.cpu()
t0 = int(t[0, 0])
names = ["c:top", "c:qk landed", "c:S0 issued", "c:S1 issued", "c:P0 ready", "c:P1 ready",
         "c:PV0 issued", "c:PV1 issued",
         "w0:wait S", "w0:S ready", "w0:max done", "w0:P done", "w0:O ready", "w0:stored",
         "w4:wait S", "w4:S ready", "w4:max done", "w4:P done", "w4:O ready", "w4:stored"]
for k in range(7):
    print(f"item {k}: " + "  ".join(f"{names[i]}={int(t[i, k]) - t0}" for i in range(20)))
