"""Generates tests/golden/hsic_golden.npz: the reference's UNMODIFIED HSIC importance ranking
(core/imp_rank.py `mlp_neuron_rank` :16-47, `attn_head_rank` :93-129, `HSICLoss` :204-239) run on
CPU (mode='cpu') over a seeded stand-in model whose Mlp / Attention sub-modules expose the
observers the rank functions read (`neuron_output`, `head_output`).

Run in the build container (the reference is not present on the GPU box):
    python tests/golden/make_hsic_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_shim  # noqa: E402

OUT = Path(__file__).resolve().parent
B, N, FEAT, HEADS, HD, CLASSES, LAYERS = 12, 9, 20, 4, 8, 7, 2


class Mlp(nn.Module):  # found by `'Mlp' in str(m) and 'Attention' not in str(m)`
    neuron_output = None


class Attention(nn.Module):
    head_output = None


class Blk(nn.Module):  # str() holds both names: skipped by the reference's predicate
    def __init__(self):
        super().__init__()
        self.attn, self.mlp = Attention(), Mlp()


class StandIn(nn.Module):
    """model(data) fills the observers from seeded tensors and returns seeded logits."""

    def __init__(self, neuron_outputs, head_outputs, logits):
        super().__init__()
        self.blocks = nn.ModuleList([Blk() for _ in neuron_outputs])
        self._obs = (neuron_outputs, head_outputs, logits)

    def forward(self, data):
        for blk, no, ho in zip(self.blocks, self._obs[0], self._obs[1]):
            blk.mlp.neuron_output, blk.attn.head_output = no, ho
        return self._obs[2]


def observers(seed=77):
    g = torch.Generator().manual_seed(seed)
    # post-GELU-like activations with very different magnitudes per neuron / head
    # (scales shuffled so that the expected ranks are not the identity)
    no = [F.gelu(torch.randn(B, N, FEAT, generator=g) *
                 torch.linspace(0.2, 3.0, FEAT)[torch.randperm(FEAT, generator=g)])
          for _ in range(LAYERS)]
    ho = [torch.randn(B, N, HEADS, HD, generator=g) *
          torch.linspace(0.5, 2.0, HEADS)[torch.randperm(HEADS, generator=g)].view(1, 1, HEADS, 1)
          + torch.randn(B, N, 1, 1, generator=g) * 0.5 for _ in range(LAYERS)]
    logits = torch.randn(B, CLASSES, generator=g) * 2
    return no, ho, logits


def main():
    ref_shim.install()
    from core import imp_rank  # the reference, unmodified
    no, ho, logits = observers()
    model = StandIn(no, ho, logits)
    loader = [(torch.zeros(B, 3, 4, 4), torch.zeros(B, dtype=torch.long))] * 2  # only batch 0 is read
    n_rank = imp_rank.mlp_neuron_rank(model, loader, mode='cpu')
    h_rank = imp_rank.attn_head_rank(model, loader, mode='cpu')
    rel = imp_rank.HSICLoss(y_kernel='linear', mean_sub=True)
    red = imp_rank.HSICLoss(y_kernel='rbf', mean_sub=False)
    y = F.softmax(logits, dim=-1)
    out = {'logits': logits.numpy()}
    for l in range(LAYERS):
        out[f'neuron_output_{l}'] = no[l].numpy()
        out[f'head_output_{l}'] = ho[l].numpy()
        out[f'neuron_rank_{l}'] = np.asarray(n_rank[l])
        out[f'head_rank_{l}'] = np.asarray(h_rank[l])
        out[f'neuron_hsic_{l}'] = np.array([rel(no[l][:, :, f], y).item() for f in range(FEAT)])
        xm = ho[l].mean(-1)
        out[f'head_rel_{l}'] = np.array([rel(xm[:, :, h], y).item() for h in range(HEADS)])
        out[f'head_red_{l}'] = np.array([[red(xm[:, :, a], xm[:, :, b]).item() for b in range(HEADS)]
                                         for a in range(HEADS)])
    np.savez_compressed(OUT / 'hsic_golden.npz', **out)
    print('wrote', OUT / 'hsic_golden.npz', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
