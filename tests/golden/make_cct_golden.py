"""Generates tests/golden/cct_golden.npz by running the UNMODIFIED reference CCT
(/root/reference/models/cct.py, models/ensemble_models.py) on the seeded synthetic weights /
inputs of devit_b200/synth.py.  Run in the build container:  python tests/golden/make_cct_golden.py"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_shim  # noqa: E402
from devit_b200 import synth  # noqa: E402


def main():
    ref_shim.install()
    import models.cct as cct  # reference
    import models.ensemble_models as ens  # reference
    out = {}
    x = synth.cifar_images(4)
    with torch.no_grad():
        for name, n_conv, tokens in (('3x1', 1, 256), ('3x2', 2, 64)):
            m = cct.get_decct(num_classes=100, kernel_size=3, n_conv_layers=n_conv, img_size=32)
            sd = synth.cct_state_dict(0, n_conv=n_conv, tokens=tokens, num_classes=100)
            assert list(m.state_dict().keys()) == list(sd.keys()), 'state_dict layout differs'
            m.load_state_dict(sd)
            m.eval()
            logits, pool = m(x, output_pool=True)
            out[f'logits_{name}'] = logits.numpy()
            out[f'pool_{name}'] = pool.numpy()
        # ensemble: 4 backbone sub-models + EnsembleCCT (teacher_size None and 512)
        n_sub = 4
        subs = []
        for s in range(n_sub):
            m = cct.get_decct(num_classes=25, kernel_size=3, n_conv_layers=1, img_size=32, backbone=True)
            sd = synth.cct_state_dict(s, n_conv=1, tokens=256, backbone=True)
            assert list(m.state_dict().keys()) == list(sd.keys())
            m.load_state_dict(sd)
            subs.append(m.eval())
        feats = [m(x) for m in subs]                      # full [B, 256] features
        out['ens_feats'] = torch.stack(feats).numpy()
        for tag, ts in (('plain', None), ('mlp', 512)):
            e = ens.EnsembleCCT(sub_size=256, teacher_size=ts, num_sub_models=n_sub, num_classes=100)
            e.load_state_dict(synth.ensemble_cct_state_dict(n_sub, 256, ts, 100))
            out[f'ens_logits_{tag}'] = e.eval()(feats).numpy()
    np.savez_compressed(Path(__file__).resolve().parent / 'cct_golden.npz', **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
