"""BASELINE config C1: deit_base_distilled_patch16_224 teacher forward, synthetic bs 256, 1 GPU.
   python tools/bench_teacher.py [batch] [precision]
Prints one JSON line (images/sec, ms/step, TFLOP/s vs the measured bf16 peak, per-family times)."""
import json
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from devit_b200 import _lib as L, synth  # noqa: E402
from devit_b200.registry import create_model  # noqa: E402
import devit_b200.models  # noqa: E402,F401

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
t = create_model('deit_base_distilled_patch16_224', num_classes=100)
t.load_state_dict(synth.teacher_state_dict(100))
t = t.cuda().eval().set_precision(prec)
x = synth.images(B).cuda()
with torch.no_grad():
    for _ in range(3):
        out = t(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        t(x)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        gout = t(x)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(gout, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    L.profile_enable(True)
    for _ in range(3):
        t(x)
    torch.cuda.synchronize()
    fam = {k: round(v[0] / 3, 3) for k, v in L.profile_collect().items()}
    L.profile_enable(False)
D, H, F, N, P, depth = 768, 12, 3072, 198, 196, 12
fl = 2 * P * 768 * D + depth * (2 * N * D * 3 * D + 4 * H * N * N * 64 + 2 * N * D * D + 4 * N * D * F) \
    + 2 * 2 * D * 100
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() \
    else {"bf16_tflops_sustained": 1400.0}
tf = fl * B / (ms / 1e3) / 1e12
print(json.dumps({"workload": f"deit_base_distilled_patch16_224 teacher forward, bs {B}, {prec}",
                  "images_per_sec": B / (ms / 1e3), "ms_per_step": ms, "tflops": tf,
                  "frac_of_sustained_bf16": tf / peaks["bf16_tflops_sustained"],
                  "gflop_per_image": fl / 1e9, "families_ms": fam}))
