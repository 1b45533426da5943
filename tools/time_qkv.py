"""QKV GEMM at the bs-256 shapes, plain vs LayerNorm-folded epilogue: python tools/time_qkv.py"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import _lib as L  # noqa: E402
M, D = 256 * 198, 384
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, D, device="cuda", generator=g)
xb, st1 = L.rowstats(x)
st6 = torch.randn(6, M, 2, device="cuda", generator=g).abs()
def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for heads in (3, 4, 5, 6):
    N = 192 * heads
    w = (torch.randn(N, D, device="cuda", generator=g) * .05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    c1 = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    res = []
    for bn in (0, 192, 256):
        t0 = timeit(lambda: L.gemm(xb, w, bias=b, out=out, out_kind=L.OUT_BF16, block_n=bn))
        t1 = timeit(lambda: L.gemm(xb, w, bias=b, out=out, out_kind=L.OUT_BF16, block_n=bn,
                                   ln_stats=st6, ln_colsum=c1, ln_dim=D, ln_eps=1e-6))
        res.append(f"bn={bn}: plain {t0:.1f} fold {t1:.1f}")
    print(f"N={N}: " + " | ".join(res))
