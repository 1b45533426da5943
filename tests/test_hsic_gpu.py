"""HSIC importance ranking on the device modules: the drop-ins of
core/imp_rank.py `mlp_neuron_rank` / `attn_head_rank` run on a `dedeit` on the GPU -- a plain
model(data), the lazily materialised observers, one batched evaluation per layer -- and their raw
observers are compared with the oracle's, their batched HSIC estimates with the per-unit oracle.
(The ranking arithmetic itself is pinned to the reference on the CPU: tests/test_hsic_cpu.py.)"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from devit_b200 import models  # noqa: F401  (registers the entrypoints)
from devit_b200 import shrink, synth
from devit_b200.registry import create_model
from oracle import devit_oracle as O
from oracle import hsic_oracle as HO

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    return np.abs(a - b).max() / np.abs(b).max()


def test_rank_functions_on_the_device_modules():
    Bn = 6
    sd = synth.dedeit_state_dict(0, num_classes=25)
    m = create_model('dedeit', num_classes=25)
    m.load_state_dict(sd)
    m = m.cuda().eval().set_precision('fp32')
    x = synth.images(Bn)
    loader = [(x, torch.zeros(Bn, dtype=torch.long))] * 2  # host batch, moved by mode='cuda'
    n_rank = shrink.mlp_neuron_rank(m, loader)
    h_rank = shrink.attn_head_rank(m, loader)
    assert len(n_rank) == 12 and len(h_rank) == 12
    assert all(sorted(r.tolist()) == list(range(1536)) for r in n_rank)
    assert all(sorted(r.tolist()) == list(range(6)) for r in h_rank)
    # the ranks drive the reference's mask -> gate protocol on the same modules
    shrink.mlp_neuron_shrink(m, shrink.mlp_neuron_mask(m, [0.5] * 12, n_rank))
    assert shrink.check_neuron_sparsity(m) == [0.5] * 12
    shrink.mlp_neuron_restore(m)

    # raw estimates of layer 0 against the per-unit oracle on the oracle's own observers
    with torch.no_grad():
        ref_logits = O.forward_logits(sd, x)
        xe = O.embed_tokens(sd, x)
        ln = F.layer_norm(xe, (384,), sd['blocks.0.norm1.weight'], sd['blocks.0.norm1.bias'], 1e-6)
        y, head_out = O.attention(sd, 'blocks.0.attn.', ln, 6)
        ln2 = F.layer_norm(xe + y, (384,), sd['blocks.0.norm2.weight'], sd['blocks.0.norm2.bias'],
                           1e-6)
        _, hidden = O.mlp(sd, 'blocks.0.mlp.', ln2)
    out = m(x.cuda())
    assert rel(out, ref_logits) < 1e-4
    no, ho = m.blocks[0].mlp.neuron_output, m.blocks[0].attn.head_output
    # the observers are those of THIS batch
    assert no.is_cuda and rel(no, hidden) < 1e-4 and rel(ho, head_out) < 1e-4
    # the batched estimators on the device against the per-unit oracle on the SAME observers.
    # Unit-variance scaling: with the std-0.02 synthetic weights the raw features lie so close
    # together that every Gaussian kernel is ~1 and the centred estimates are rounding noise.
    prob = F.softmax(out.float(), dim=-1)
    xn = (no.float() / no.float().std()).permute(2, 0, 1).contiguous()  # [1536, B, 198]
    got = shrink.hsic_relevance(xn, prob)
    assert got.is_cuda and got.shape == (1536,)
    xn_c, prob_c = xn.cpu(), prob.cpu()
    units = list(range(0, 1536, 48))  # 32 neurons spread over the layer
    want = [HO.hsic(xn_c[f], prob_c, 'linear', True).item() for f in units]
    assert rel(got[units], want) < 1e-3
    xm = ho.float().mean(-1)
    xm = (xm / xm.std()).permute(2, 0, 1).contiguous()  # [6, B, 198]
    xm_c = xm.cpu()
    assert rel(shrink.hsic_relevance(xm, prob),
               [HO.hsic(xm_c[h], prob_c, 'linear', True).item() for h in range(6)]) < 1e-3
    assert rel(shrink.hsic_redundancy(xm),
               [[HO.hsic(xm_c[a], xm_c[b], 'rbf', False).item() for b in range(6)]
                for a in range(6)]) < 1e-3
