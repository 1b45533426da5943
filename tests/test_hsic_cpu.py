"""HSIC importance ranking (core/imp_rank.py:16-47, :93-129, :176-239):
the oracle restatement against fixtures produced by the reference's own functions
(tests/golden/make_hsic_golden.py), the batched product implementation (devit_b200/shrink.py)
against both, and -- where the reference is importable -- against the reference itself on fresh
random observers."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F
from pathlib import Path

from devit_b200 import shrink
from oracle import hsic_oracle as HO

GOLD = np.load(Path(__file__).parent / 'golden' / 'hsic_golden.npz')
LAYERS = 2


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / np.abs(b).max()


def fixtures(l):
    t = lambda k: torch.from_numpy(GOLD[k])  # noqa: E731
    return t(f'neuron_output_{l}'), t(f'head_output_{l}'), t('logits')


class Mlp(nn.Module):  # the names the rank functions search for in str(m)
    neuron_output = None


class Attention(nn.Module):
    head_output = None


class Blk(nn.Module):
    def __init__(self):
        super().__init__()
        self.attn, self.mlp = Attention(), Mlp()


class StandIn(nn.Module):
    """model(data) fills the observers and returns the logits (what a forward of the real modules
    does, tests/test_model_gpu.py checks those observers against the oracle on the GPU)."""

    def __init__(self, neuron_outputs, head_outputs, logits):
        super().__init__()
        self.blocks = nn.ModuleList([Blk() for _ in neuron_outputs])
        self._obs = (neuron_outputs, head_outputs, logits)
        self.calls = 0

    def forward(self, data):
        self.calls += 1
        for blk, no, ho in zip(self.blocks, self._obs[0], self._obs[1]):
            blk.mlp.neuron_output, blk.attn.head_output = no, ho
        return self._obs[2]


def stand_in():
    no, ho = zip(*[fixtures(l)[:2] for l in range(LAYERS)])
    model = StandIn(list(no), list(ho), fixtures(0)[2])
    loader = [(torch.zeros(12, 3, 4, 4), torch.zeros(12, dtype=torch.long))] * 3
    return model, loader


@pytest.mark.parametrize('l', range(LAYERS))
def test_oracle_matches_the_reference_fixtures(l):
    no, ho, logits = fixtures(l)
    y = F.softmax(logits, dim=-1)
    hs = [HO.hsic(no[:, :, f], y, 'linear', True).item() for f in range(no.shape[-1])]
    assert rel(hs, GOLD[f'neuron_hsic_{l}']) < 1e-5
    xm = ho.mean(-1)
    H = xm.shape[-1]
    assert rel([HO.hsic(xm[:, :, h], y, 'linear', True).item() for h in range(H)],
               GOLD[f'head_rel_{l}']) < 1e-5
    red = [[HO.hsic(xm[:, :, a], xm[:, :, b], 'rbf', False).item() for b in range(H)]
           for a in range(H)]
    assert rel(red, GOLD[f'head_red_{l}']) < 1e-5
    # index selection is integer work: exact
    assert np.array_equal(np.argsort(HO.neuron_scores(no, logits)), GOLD[f'neuron_rank_{l}'])
    assert np.array_equal(np.argsort(HO.head_scores(ho, logits)), GOLD[f'head_rank_{l}'])


@pytest.mark.parametrize('l', range(LAYERS))
def test_batched_estimators_match_fixtures_and_oracle(l):
    no, ho, logits = fixtures(l)
    y = F.softmax(logits, dim=-1)
    hs = shrink.hsic_relevance(no.permute(2, 0, 1), y, chunk=7)  # ragged chunks on purpose
    assert rel(hs.numpy(), GOLD[f'neuron_hsic_{l}']) < 1e-4
    xm = ho.mean(-1).permute(2, 0, 1)
    assert rel(shrink.hsic_relevance(xm, y).numpy(), GOLD[f'head_rel_{l}']) < 1e-4
    assert rel(shrink.hsic_redundancy(xm).numpy(), GOLD[f'head_red_{l}']) < 1e-4
    assert rel(shrink.neuron_scores(no, logits), HO.neuron_scores(no, logits)) < 1e-4
    assert rel(shrink.head_scores(ho, logits), HO.head_scores(ho, logits)) < 1e-4


def test_rank_functions_match_the_reference_fixtures():
    """Drop-in signatures of core/imp_rank.py:16,93; only the loader's first batch is used."""
    model, loader = stand_in()
    n_rank = shrink.mlp_neuron_rank(model, loader, mode='cpu')
    assert model.calls == 1
    h_rank = shrink.attn_head_rank(model, loader, mode='cpu')
    assert len(n_rank) == LAYERS and len(h_rank) == LAYERS
    for l in range(LAYERS):
        assert np.array_equal(n_rank[l], GOLD[f'neuron_rank_{l}'])
        assert np.array_equal(h_rank[l], GOLD[f'head_rank_{l}'])
    # ... and they feed the mask functions like the reference's ranks do (core/imp_rank.py:50-62)
    for blk in model.blocks:
        blk.mlp.hidden_features, blk.attn.num_heads = 20, 4
    masks = shrink.mlp_neuron_mask(model, [0.5] * LAYERS, n_rank)
    assert [int(m.sum()) for m in masks] == [10] * LAYERS
    assert all(masks[l][GOLD[f'neuron_rank_{l}'][-1]] == 1 for l in range(LAYERS))


def test_float64_estimators_agree_with_float32():
    """The fp32 estimates sit ~1e-6 from their fp64 values: rank flips need closer ties than that."""
    no, ho, logits = fixtures(0)
    y = F.softmax(logits.double(), dim=-1)
    h64 = shrink.hsic_relevance(no.double().permute(2, 0, 1), y)
    h32 = shrink.hsic_relevance(no.permute(2, 0, 1), y.float())
    assert h64.dtype == torch.float64 and rel(h32.numpy(), h64.numpy()) < 1e-5


def test_against_the_reference_itself_on_fresh_observers():
    from oracle import ref_shim
    if ref_shim.reference_root() is None:
        pytest.skip("no /root/reference and no oracle/_ref copy on this machine")
    ref_shim.install()
    from core import imp_rank
    g = torch.Generator().manual_seed(5)
    B, N, Fh, H, hd, Cn = 10, 6, 33, 3, 4, 5
    no = [F.gelu(torch.randn(B, N, Fh, generator=g) * (1 + 2 * torch.rand(Fh, generator=g)))]
    ho = [torch.randn(B, N, H, hd, generator=g) * (1 + torch.rand(H, generator=g)).view(1, 1, H, 1)]
    logits = torch.randn(B, Cn, generator=g)
    model = StandIn(no, ho, logits)
    loader = [(torch.zeros(B, 3, 4, 4), torch.zeros(B, dtype=torch.long))]
    assert np.array_equal(shrink.mlp_neuron_rank(model, loader, mode='cpu')[0],
                          imp_rank.mlp_neuron_rank(model, loader, mode='cpu')[0])
    assert np.array_equal(shrink.attn_head_rank(model, loader, mode='cpu')[0],
                          imp_rank.attn_head_rank(model, loader, mode='cpu')[0])
