"""Runs the UNMODIFIED reference modules (models/ensemble_models.py MultiViT + EnsMLP over
models/de_vit.py `dedeit`, the models/deit_vit.py teacher, the models/cct.py decomposed CCT) on
the host CPU, on the seeded synthetic weights / gates of devit_b200/synth.py.  The modules come
from /root/reference or from the byte copy under oracle/_ref (oracle/make_ref.py) through the
import shims of oracle/ref_shim.py.

Test infrastructure: this is the `cpu_baseline.kind = "reference"` arm of bench.py and the
reference side of the oracle cross-checks in tests/.  Nothing under devit_b200/ imports it.
"""
from __future__ import annotations

import sys

import torch

from devit_b200 import synth
from . import ref_shim


def available() -> bool:
    return ref_shim.reference_root() is not None


def _imp_rank():
    ref_shim.install()
    from core import imp_rank  # reference, unmodified (core/imp_rank.py)
    return imp_rank


def ensemble(n_sub=4, num_class=100, shrunk=True):
    """-> callable x[B,3,224,224] -> logits[B,num_class]; the reference's own
    MultiViT.forward (models/ensemble_models.py:32-40) + EnsMLP.forward (:65-90), gates set by
    the reference's mlp_neuron_shrink / attn_head_shrink (core/imp_rank.py:65-71,147-153)."""
    _, _, ens, _ = ref_shim.load_reference()
    multi = ens.MultiViT(model='dedeit', drop=0, drop_path=0.1,
                         num_classes_list=[num_class // n_sub] * n_sub, num_div=n_sub)
    fuse = ens.EnsMLP(model='dedeit', num_class=num_class, sub_size=384,
                      num_classes_list=[num_class // n_sub] * n_sub, teacher_size=768)
    for s in range(n_sub):
        multi.backbones[s].load_state_dict(synth.dedeit_state_dict(s, with_heads=False))
    fuse.load_state_dict(synth.ensmlp_state_dict(n_sub, num_class=num_class))
    if shrunk:
        ir = _imp_rank()
        for s in range(n_sub):
            ng, hg = synth.shrink_gates(s)
            ir.mlp_neuron_shrink(multi.backbones[s], ng)
            ir.attn_head_shrink(multi.backbones[s], hg)
    multi.eval(), fuse.eval()

    def run(x):
        with torch.no_grad():
            return fuse(multi(x))
    return run


def teacher(num_class=100):
    """deit_base_distilled_patch16_224 of models/deit_vit.py:477-485, eval forward (:251-296)."""
    _, _, _, create_model = ref_shim.load_reference()
    m = create_model('deit_base_distilled_patch16_224', num_classes=num_class)
    m.load_state_dict(synth.teacher_state_dict(num_class))
    m.eval()

    def run(x):
        with torch.no_grad():
            return m(x)
    return run
