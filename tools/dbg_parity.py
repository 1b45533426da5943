"""Prints the bf16-mode relative errors of the single / stress / teacher goldens (debug aid)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from devit_b200 import synth  # noqa: E402
from devit_b200.registry import create_model  # noqa: E402
import devit_b200.models  # noqa: E402,F401
import test_model_gpu as T  # noqa: E402

x = synth.images(T.B).cuda()
for precision in ("bf16",):
    m = T.make_sub(0, precision)
    print("single", T.rel(m(x), T.G['single_logits']))
    st = T.make_sub(7, precision, qkv_gain=3.0)
    print("stress", T.rel(st(x), T.G['stress_logits']))
    t = create_model('deit_base_distilled_patch16_224', num_classes=100)
    t.load_state_dict(synth.teacher_state_dict(100))
    t = t.cuda().eval().set_precision(precision)
    print("teacher", T.rel(t(x[:2]), T.G['teacher_logits']))
