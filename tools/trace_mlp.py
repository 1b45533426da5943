"""clock64 trace of the fused MLP kernel, CTA 0 (needs the -DDEVIT_GEMM_TRACE build):
   DEVIT_B200_LIB=devit_b200/lib/libdevit_b200_trace.so python tools/trace_mlp.py [hidden]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import _lib as L  # noqa: E402
import os
F = int(sys.argv[1]) if len(sys.argv) > 1 else 928
H = int(os.environ.get('TIME_MLP_PROJ', '0'))  # > 0: the variant with the projection in front
M, D = int(os.environ.get('TRACE_MLP_BATCH', '256')) * 198, 384
g = torch.Generator(device="cuda").manual_seed(0)
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
def run():
    if H:
        L.mlp_fused(x, None, None, w1, c1, c2, w2, b2, 1e-6, xb_out=xb, stats_out=so,
                    o=o, w_proj=wp, b_proj=bp)
    else:
        L.mlp_fused(x, xb, stats, w1, c1, c2, w2, b2, 1e-6, xb_out=xb, stats_out=so)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print(f"F={F}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch")
buf = torch.zeros(20, 512, device="cuda", dtype=torch.int64)
L.load().devit_debug_set_trace(buf.data_ptr())
run()
torch.cuda.synchronize()
L.load().devit_debug_set_trace(None)
t = buf.cpu()
t0 = int(t[7, 0])
NC = (F + 63) // 64
def r(s, i): return int(t[s, i]) - t0
for it in range(3):
    if H:
        print(f"tile {it}: mma wait O {r(7,it)} got {r(8,it)} GEMM0 issued {r(18,it)} y_ready {r(19,it)} | "
              f"epi0 wait p_full {r(12,it)} got {r(13,it)} done {r(14,it)} | final: acc2_full {r(16,it*4+3)} "
              f"chunks done {[r(17, it * 4 + j) for j in range(3)]} stores drained {r(15,it*4+3)}")
        print(f"   GEMM0: x_loaded seen {r(0, 400 + it * 8 + 7)}  Wp chunk landed {[r(1, 400 + it * 8 + a) for a in range(H)]}")
    else:
        print(f"tile {it}: mma wait Y {r(7,it)} got {r(8,it)} | epi wait acc2 {r(12,it)} got {r(13,it)} final done {r(14,it)}")
        print("   final chunks [before resid wait, resid landed, chunk done]: " + "  ".join(
            f"j{j}: {r(15, it * 4 + j)} {r(16, it * 4 + j)} {r(17, it * 4 + j)}" for j in range(3)))
    for c in range(it * NC, it * NC + NC):
        print(f"   c{c - it * NC:2d}: G1 [wait {r(0,c)} got {r(1,c)} issued {r(2,c)}]  G2 [wait {r(3,c)} h {r(4,c)} w2 {r(5,c)} issued {r(6,c)}]  epi [wait {r(9,c)} got {r(10,c)} done {r(11,c)}]")
