"""devit_pack_layer (C ABI, csrc/pack.cu: gate compaction + LayerNorm folding on the device) against
the same packing written with torch ops (devit_b200/packing.py, DEVIT_PACK_TORCH=1): kept-index
lists identical, bf16 / split-fp32 weight operands bit-identical, folded row sums and biases to
fp32 summation noise -- for shrunk gates, non-binary gate values and a layer with every head off."""
import numpy as np
import pytest
import torch

from devit_b200 import _lib as L
from devit_b200 import models, packing, shrink, synth  # noqa: F401
from devit_b200.registry import create_model

pytestmark = pytest.mark.gpu


def build(precision):
    m = create_model('dedeit', num_classes=25)
    m.load_state_dict(synth.dedeit_state_dict(3, num_classes=25))
    ng, hg = synth.shrink_gates(3)
    shrink.mlp_neuron_shrink(m, ng)
    shrink.attn_head_shrink(m, hg)
    # non-binary gate values, a layer with every head gated off, a layer with 5 kept neurons
    g = m.blocks[2].mlp.gate.clone()
    g[g != 0] = torch.linspace(0.5, 1.5, int((g != 0).sum()))
    m.blocks[2].mlp.gate = g
    m.blocks[4].attn.gate = torch.tensor([0.0, 2.0, 0.0, 0.5, 0.0, 1.0])
    m.blocks[5].attn.gate = torch.zeros(6)
    tiny = torch.zeros(1536)
    tiny[[3, 77, 500, 1000, 1535]] = 1.0
    m.blocks[6].mlp.gate = tiny
    return m.cuda().eval().set_precision(precision)


@pytest.mark.parametrize('precision', ['bf16', 'fp32'])
def test_pack_layer_matches_the_torch_packing(precision, monkeypatch):
    m = build(precision)
    prec = models._PREC[precision]
    dev = torch.device('cuda', torch.cuda.current_device())
    monkeypatch.setenv('DEVIT_PACK_TORCH', '1')
    ref = packing.PackedVit(m, prec, dev)
    monkeypatch.setenv('DEVIT_PACK_TORCH', '0')
    got = packing.PackedVit(m, prec, dev)
    for i in range(len(m.blocks)):
        assert torch.equal(got.kept_heads[i], ref.kept_heads[i]), i
        assert torch.equal(got.kept_neurons[i], ref.kept_neurons[i]), i
        dg, dr = got.layers[i], ref.layers[i]
        assert (dg.heads, dg.hidden, dg.hidden_ld) == (dr.heads, dr.hidden, dr.hidden_ld), i
        a, b = got.layer_arrays(i), ref.layer_arrays(i)
        for k in ('w_qkv', 'w_proj', 'w_fc1', 'w_fc2'):
            assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), (i, k)
        for k in ('b_qkv', 'b_fc1', 'cs_qkv', 'cs_fc1'):
            if b[k] is None:
                assert a[k] is None, (i, k)
                continue
            err = (a[k].double() - b[k].double()).abs().max().item()
            scale = max(b[k].double().abs().max().item(), 1e-6)
            assert err <= 2e-6 * scale + 1e-7, (i, k, err, scale)
    # and the forward through either pack agrees to that noise
    x = synth.images(2).cuda()
    m._packs = {(precision, str(dev)): (packing.module_version(m), ref)}
    out_ref = m(x)
    m._packs = {(precision, str(dev)): (packing.module_version(m), got)}
    out_got = m(x)
    r = ((out_got - out_ref).abs().max() / out_ref.abs().max()).item()
    # (bf16: 1e-6-level differences in c1 / c2 flip individual bf16 roundings down the 12 layers)
    assert r < (8e-3 if precision == 'bf16' else 1e-5), r


def test_pack_layer_argument_checks():
    w = L.BlockWeights()
    w.dim, w.num_heads, w.hidden = 384, 5, 1536  # 384 / 5 is not a head_dim of 64
    assert L.load().devit_pack_layer_bytes(w, L.DEVIT_BF16) == 0
    assert b'geometry' in L.load().devit_last_error() or b'head_dim' in L.load().devit_last_error()
