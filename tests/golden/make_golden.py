"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on the
seeded synthetic weights / inputs / gates of devit_b200/synth.py.  Run in the build container
(the reference is not present on the GPU box):   python tests/golden/make_golden.py
The oracle (oracle/devit_oracle.py) and the CUDA path are both checked against these files.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_shim  # noqa: E402
from devit_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    torch.manual_seed(0)
    de_vit, deit_vit, ens, create_model = ref_shim.load_reference()
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    from core import imp_rank, compute_metric  # reference, unmodified

    n_sub, B = 4, 4
    x = synth.images(B)
    multi = ens.MultiViT(model='dedeit', drop=0, drop_path=0.1, num_classes_list=[25] * n_sub,
                         num_div=n_sub)
    fuse = ens.EnsMLP(model='dedeit', num_class=100, sub_size=384, num_classes_list=[25] * n_sub,
                      teacher_size=768)
    for s in range(n_sub):
        multi.backbones[s].load_state_dict(synth.dedeit_state_dict(s, with_heads=False))
    fuse.load_state_dict(synth.ensmlp_state_dict(n_sub))
    multi.eval(), fuse.eval()
    out = {}
    with torch.no_grad():
        # dense gates (all ones)
        cls, dist = multi(x)
        out['dense_logits'] = fuse((cls, dist)).numpy()
        out['dense_cls'] = torch.stack(cls).numpy()
        out['dense_dist'] = torch.stack(dist).numpy()
        # shrunk gates through the reference's own mask/shrink functions
        kept_h, kept_n = [], []
        for s in range(n_sub):
            rng = np.random.RandomState(4321 + s)
            from devit_b200 import shrink
            n_ratio, h_ratio = shrink.sample_policy(rng)
            n_rank = [rng.permutation(1536) for _ in range(12)]
            h_rank = [rng.permutation(6) for _ in range(12)]
            bb = multi.backbones[s]
            nm = imp_rank.mlp_neuron_mask(bb, n_ratio, n_rank)
            hm = imp_rank.attn_head_mask(bb, h_ratio, h_rank)
            imp_rank.mlp_neuron_shrink(bb, nm)
            imp_rank.attn_head_shrink(bb, hm)
            out[f'policy{s}_ratios'] = np.array(n_ratio + h_ratio)
            out[f'policy{s}_macs'] = np.array(compute_metric.cal_shrink_macs(
                n_ratio, h_ratio, emb=384, mlp_ratio=4, seq_length=197, head=6, layer=12))
            out[f'policy{s}_neuron_masks'] = torch.stack(nm).numpy().astype(np.uint8)
            out[f'policy{s}_head_masks'] = torch.stack(hm).numpy().astype(np.uint8)
        cls, dist = multi(x)
        out['shrunk_logits'] = fuse((cls, dist)).numpy()
        out['shrunk_cls'] = torch.stack(cls).numpy()
        out['shrunk_dist'] = torch.stack(dist).numpy()

        # per-block state of sub-model 0 (shrunk): a slice + checksums of the residual stream
        bb = multi.backbones[0]
        feats = bb.forward_features(x, output_emb=True, output_encoders=True)
        enc = feats['encoder']  # [emb, block0, ..., block11]
        out['sub0_block_absmean'] = np.array([e.abs().mean().item() for e in enc])
        out['sub0_block_slice'] = torch.stack([e[:, :4, :16] for e in enc]).numpy()

        # single sub-model eval logits with heads (engine.evaluate path)
        single = create_model('dedeit', num_classes=25, drop_rate=0, drop_path_rate=0.1,
                              drop_block_rate=None)
        single.load_state_dict(synth.dedeit_state_dict(0, num_classes=25))
        single.eval()
        out['single_logits'] = single(x).numpy()

        # sharper softmax (stress weights), dense
        stress = create_model('dedeit', num_classes=25)
        stress.load_state_dict(synth.dedeit_state_dict(7, num_classes=25, qkv_gain=3.0))
        stress.eval()
        out['stress_logits'] = stress(x).numpy()

        # teacher (models/deit_vit.py), 100 classes
        teacher = create_model('deit_base_distilled_patch16_224', num_classes=100)
        teacher.load_state_dict(synth.teacher_state_dict(100))
        teacher.eval()
        out['teacher_logits'] = teacher(x[:2]).numpy()
    np.savez_compressed(OUT / 'devit_golden.npz', **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))


if __name__ == '__main__':
    main()
