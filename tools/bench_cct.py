"""BASELINE config C4: decomposed CCT ensemble (4 x decct_7_3x1 backbones + EnsembleCCT),
synthetic 32x32 batch, 1 GPU (sub-models sequential).   python tools/bench_cct.py [batch] [3x1|3x2]
Prints one JSON line (images/sec, ms/step, TFLOP/s, per-family device times)."""
import json
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from devit_b200 import _lib as L, cct, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cfg = sys.argv[2] if len(sys.argv) > 2 else "3x1"
n_conv, tokens = (1, 256) if cfg == "3x1" else (2, 64)
n_sub = 4
multi = cct.MultiCCT(f'decct_7_{cfg}', num_classes_list=[25] * n_sub, num_sub_models=n_sub, input_size=32)
for s in range(n_sub):
    multi.models[s].load_state_dict(synth.cct_state_dict(s, n_conv=n_conv, tokens=tokens, backbone=True))
fuse = cct.EnsembleCCT(sub_size=256, teacher_size=None, num_sub_models=n_sub, num_classes=100)
fuse.load_state_dict(synth.ensemble_cct_state_dict(n_sub, 256, None, 100))
multi, fuse = multi.cuda().eval().set_precision('bf16'), fuse.cuda().eval().set_precision('bf16')
x = synth.cifar_images(B).cuda()
def step():
    return fuse(multi(x))
with torch.no_grad():
    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        gout = step()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(gout, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    L.profile_enable(True)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    fam = {k: round(v[0] / 2, 3) for k, v in L.profile_collect().items()}
    L.profile_enable(False)
D, F, depth, H = 256, 512, 7, 4
side = 32
conv_fl = 0
cin = 3
for i in range(n_conv):
    cout = 256 if i == n_conv - 1 else 64
    conv_fl += 2 * side * side * 9 * cin * cout
    cin, side = cout, side // 2
blk = depth * (2 * tokens * D * 3 * D + 4 * H * tokens * tokens * 64 + 2 * tokens * D * D + 4 * tokens * D * F)
fl = n_sub * (conv_fl + blk) + 2 * n_sub * D * 100
print(json.dumps({"workload": f"4-way decct_7_{cfg} ensemble (EnsembleCCT 100 classes), 32x32, bs {B}, bf16, 1 GPU",
                  "images_per_sec": B / (ms / 1e3), "ms_per_step": ms,
                  "tflops": fl * B / (ms / 1e3) / 1e12, "gflop_per_image": fl / 1e9,
                  "families_ms": fam}))
