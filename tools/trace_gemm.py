"""clock64 trace of GEMM CTA 0: python tools/trace_gemm.py qkv [bn] [cl]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import _lib as L  # noqa: E402
M, D = 256 * 198, 384
name = sys.argv[1]
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 192
cl = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sh = {"qkv": (D, 1152, L.OUT_BF16, False, L.ACT_NONE), "proj": (D, D, L.OUT_F32, True, L.ACT_NONE),
      "fc1": (D, 1536, L.OUT_BF16, False, L.ACT_GELU_ERF), "fc2": (1536, D, L.OUT_F32, True, L.ACT_NONE)}[name]
k, n, ok, resid, act = sh
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(M, k, device="cuda", generator=g).bfloat16()
w = (torch.randn(n, k, device="cuda", generator=g) * .05).bfloat16()
b = torch.randn(n, device="cuda", generator=g)
x = torch.randn(M, D, device="cuda", generator=g)
out = x if resid else torch.empty(M, n, device="cuda", dtype=torch.bfloat16)
def run():
    L.gemm(a, w, bias=b, resid=x if resid else None, out=out, out_kind=ok, act=act, block_n=bn, cluster_m=cl)
for _ in range(3):
    run()
torch.cuda.synchronize()
buf = torch.zeros(20, 512, device="cuda", dtype=torch.int64)
L.load().devit_debug_set_trace(buf.data_ptr())
run()
torch.cuda.synchronize()
L.load().devit_debug_set_trace(None)
t = buf.cpu()
t0 = int(t[0, 0])
def rel(r, i): return int(t[r, i]) - t0
print("producer: k-block i: [start, after wait(empty), after TMA issue]")
for i in list(range(0, 14)) + list(range(40, 52)):
    print(f"  P{i:3d}: {rel(0,i):7d} {rel(1,i):7d} {rel(2,i):7d}   | M{i:3d}: start {rel(3,i):7d} waited(full) {rel(4,i):7d} mma-issued {rel(5,i):7d} committed {rel(6,i):7d}")
print("tiles: MMA [start wait tmem_empty, got]  EPI warp0 [start wait tmem_full, got, done]")
for i in range(0, 16):
    print(f"  T{i:2d}: {rel(7,i):7d} {rel(8,i):7d}  | {rel(9,i):7d} {rel(10,i):7d} {rel(11,i):7d}")
print("epilogue warp0 chunk0: [start, tmem loaded, math done, store drained, sts done, fenced, store issued] (relative to tmem_full got)")
for i in range(1, 10):
    b0 = int(t[10, i])
    print("  T%2d:" % i, [int(t[r, i]) - b0 for r in range(12, 19)])
nk = 96
print("per-k-block MMA loop period (cycles):", (rel(6, nk-1) - rel(6, 5)) / (nk - 6))
