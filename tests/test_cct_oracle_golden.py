"""The CCT oracle (oracle/cct_oracle.py) against outputs of the reference CCT itself
(tests/golden/cct_golden.npz, made by tests/golden/make_cct_golden.py in the build container)."""
from pathlib import Path

import numpy as np
import torch

from devit_b200 import synth
from oracle import cct_oracle as CO

G = np.load(Path(__file__).resolve().parent / 'golden' / 'cct_golden.npz')


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def test_cct_single_models():
    x = synth.cifar_images(4)
    with torch.no_grad():
        for name, n_conv, tokens in (('3x1', 1, 256), ('3x2', 2, 64)):
            sd = synth.cct_state_dict(0, n_conv=n_conv, tokens=tokens, num_classes=100)
            pool = CO.pooled_features(sd, x, n_conv, 7, 4)
            logits = CO.logits(sd, x, n_conv)
            assert rel(pool.numpy(), G[f'pool_{name}']) < 2e-5
            assert rel(logits.numpy(), G[f'logits_{name}']) < 2e-5
            assert (logits.argmax(-1).numpy() == G[f'logits_{name}'].argmax(-1)).all()


def test_cct_ensemble():
    x = synth.cifar_images(4)
    sds = [synth.cct_state_dict(s, n_conv=1, tokens=256, backbone=True) for s in range(4)]
    with torch.no_grad():
        for tag, ts in (('plain', None), ('mlp', 512)):
            esd = synth.ensemble_cct_state_dict(4, 256, ts, 100)
            logits, feats = CO.ensemble_logits(sds, esd, x, 1)
            assert rel(torch.stack(feats).numpy(), G['ens_feats']) < 2e-5
            assert rel(logits.numpy(), G[f'ens_logits_{tag}']) < 2e-5
            assert (logits.argmax(-1).numpy() == G[f'ens_logits_{tag}'].argmax(-1)).all()
