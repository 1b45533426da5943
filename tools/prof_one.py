"""One GEMM shape a few times (for ncu): python tools/prof_one.py qkv|proj|fc1|fc2 [bn] [cl]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import _lib as L  # noqa: E402
M, D = 256 * 198, 384
name = sys.argv[1]
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sh = {"qkv": (D, 1152, L.OUT_BF16, False, L.ACT_NONE), "proj": (D, D, L.OUT_F32, True, L.ACT_NONE),
      "fc1": (D, 1536, L.OUT_BF16, False, L.ACT_GELU_ERF), "fc2": (1536, D, L.OUT_F32, True, L.ACT_NONE)}[name]
k, n, ok, resid, act = sh
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(M, k, device="cuda", generator=g).bfloat16()
w = (torch.randn(n, k, device="cuda", generator=g) * .05).bfloat16()
b = torch.randn(n, device="cuda", generator=g)
x = torch.randn(M, D, device="cuda", generator=g)
out = x if resid else torch.empty(M, n, device="cuda", dtype=torch.bfloat16)
for _ in range(int(sys.argv[4]) if len(sys.argv) > 4 else 3):
    L.gemm(a, w, bias=b, resid=x if resid else None, out=out, out_kind=ok, act=act, block_n=bn, cluster_m=cl)
torch.cuda.synchronize()
