"""Times devit_gemm at the bs-256 dedeit shapes for every (block_n, cluster_m) pair.
   python tools/sweep_gemm.py [dense|shrunk]"""
import itertools
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import _lib as L  # noqa: E402

M, D = 256 * 198, 384
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*s, dt=torch.bfloat16, scale=1.0):
    return (torch.randn(*s, device=dev, generator=g) * scale).to(dt)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


mode = sys.argv[1] if len(sys.argv) > 1 else "dense"
H, F = (6, 1536) if mode == "dense" else (4, 928)
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
shapes = {
    "qkv": dict(k=D, n=192 * H, out=L.OUT_BF16, resid=False, act=L.ACT_NONE),
    "proj": dict(k=64 * H, n=D, out=L.OUT_F32, resid=True, act=L.ACT_NONE),
    "fc1": dict(k=D, n=F, out=L.OUT_BF16, resid=False, act=L.ACT_GELU_ERF),
    "fc2": dict(k=F, n=D, out=L.OUT_F32, resid=True, act=L.ACT_NONE),
}
x = rnd(M, D, dt=torch.float32)
for name, sh in shapes.items():
    if only and name not in only:
        continue
    a = rnd(M, sh["k"])
    w = rnd(sh["n"], sh["k"], scale=0.05)
    b = rnd(sh["n"], dt=torch.float32)
    out = x if sh["resid"] else torch.empty(M, sh["n"], device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * sh["k"] * sh["n"]
    res = []
    for bn, cl in itertools.product((128, 192, 256), (1, 2)):
        t = timeit(lambda: L.gemm(a, w, bias=b, resid=x if sh["resid"] else None, out=out,
                                  out_kind=sh["out"], act=sh["act"], block_n=bn, cluster_m=cl))
        res.append((t, bn, cl))
    res.sort()
    best = res[0]
    print(f"{name:5s} K={sh['k']:5d} N={sh['n']:5d}  best {best[0]:7.1f} us bn={best[1]} cl={best[2]} "
          f"-> {fl / best[0] / 1e6:7.1f} TFLOP/s | " +
          " ".join(f"{bn}/{cl}:{t:.0f}" for t, bn, cl in sorted(res, key=lambda r: (r[1], r[2]))))
