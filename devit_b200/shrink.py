"""Host-side mirror of the reference's gate selection (core/imp_rank.py, core/compute_metric.py,
core/shrink_imp.py:66-82).  Integer index work only -- it configures the hot path, it is not on
it.  The functions keep the reference's names, argument meaning and module-discovery protocol
(``'Mlp' in str(m) and 'Attention' not in str(m)``), so the reference's own core/imp_rank.py
functions and these are interchangeable on devit_b200 modules.
"""
from __future__ import annotations

import numpy as np
import torch


def _is_mlp(m) -> bool:
    s = str(m)
    return 'Mlp' in s and 'Attention' not in s


def _is_attn(m) -> bool:
    s = str(m)
    return 'Attention' in s and 'Mlp' not in s


def keep_mask(width: int, ratio: float, rank) -> torch.Tensor:
    """core/imp_rank.py:55-58 / :137-140: keep the int(width*(1-ratio)) highest-ranked units
    (rank = ascending argsort of the importance scores)."""
    num_keep = int(width * (1 - ratio))
    kept = np.asarray(rank)[::-1][:num_keep]
    mask = torch.zeros(width)
    mask[kept.tolist()] = 1
    return mask


def mlp_neuron_mask(model, ratio, rank):
    """core/imp_rank.py:50-62."""
    out, idx = [], 0
    for m in model.modules():
        if _is_mlp(m):
            out.append(keep_mask(m.hidden_features, ratio[idx], rank[idx]))
            idx += 1
    return out


def attn_head_mask(model, ratio, rank):
    """core/imp_rank.py:132-144."""
    out, idx = [], 0
    for m in model.modules():
        if _is_attn(m):
            out.append(keep_mask(m.num_heads, ratio[idx], rank[idx]))
            idx += 1
    return out


def mlp_neuron_shrink(model, neuron_mask):
    """core/imp_rank.py:65-71."""
    idx = 0
    for m in model.modules():
        if _is_mlp(m):
            m.gate = neuron_mask[idx]
            idx += 1


def attn_head_shrink(model, head_mask):
    """core/imp_rank.py:147-153."""
    idx = 0
    for m in model.modules():
        if _is_attn(m):
            m.gate = head_mask[idx]
            idx += 1


def mlp_neuron_restore(model):
    """core/imp_rank.py:74-81."""
    for m in model.modules():
        if _is_mlp(m):
            m.gate = torch.ones(m.gate.shape[0])


def attn_head_restore(model):
    """core/imp_rank.py:156-163."""
    for m in model.modules():
        if _is_attn(m):
            m.gate = torch.ones(m.gate.shape[0])


def check_neuron_sparsity(model):
    """core/imp_rank.py:84-90."""
    return [torch.sum(m.gate == 0).item() / m.gate.shape[0] for m in model.modules() if _is_mlp(m)]


def check_head_sparsity(model):
    """core/imp_rank.py:166-172."""
    return [torch.sum(m.gate == 0).item() / m.gate.shape[0] for m in model.modules() if _is_attn(m)]


def cal_shrink_flops(neuron_sparsity, head_sparsity, emb=768, seq_length=197, mlp_ratio=4,
                     head=12, layer=12, num_class=1000):
    """core/compute_metric.py:31-64 (GFLOPs; softmax and norms neglected)."""
    assert len(head_sparsity) == layer
    channel, img_size = 3, 224
    head_dim = emb / head
    flops = 2 * channel * emb * img_size ** 2
    for n_s, h_s in zip(neuron_sparsity, head_sparsity):
        sa = 3 * 2 * seq_length * emb * head_dim + 2 * head_dim * seq_length ** 2 \
            + 2 * head_dim * seq_length ** 2
        shrink_head = int((1 - h_s) * head)
        mhsa = sa * shrink_head + seq_length * 2 * head_dim * shrink_head * emb
        hidden = int(mlp_ratio * (1 - n_s) * emb)
        mlp = seq_length * hidden * 2 * emb + seq_length * emb * 2 * hidden
        flops += mhsa + mlp
    flops += 2 * emb * num_class
    return flops / 1e9


def cal_shrink_macs(neuron_sparsity, head_sparsity, emb=768, seq_length=197, mlp_ratio=4,
                    head=12, layer=12, num_class=1000):
    """core/compute_metric.py:67-69."""
    return cal_shrink_flops(neuron_sparsity, head_sparsity, emb, seq_length, mlp_ratio, head,
                            layer, num_class) / 2


def sample_policy(rng: np.random.RandomState, shrink_ratio=0.3, lb=0.0, ub=0.5, layer=12,
                  max_tries=1_000_000):
    """One accepted candidate of core/shrink_imp.py:66-82 (`screen`): 2*layer ratios ~ U(lb, ub)
    whose MACs land within 2 % of shrink_ratio * 9.19 (core/shrink_imp.py:144).
    Returns (neuron_ratios[layer], head_ratios[layer])."""
    target = shrink_ratio * 9.19
    for _ in range(max_tries):
        ratio = rng.uniform(lb, ub, size=(1, 2 * layer))[0].tolist()
        macs = cal_shrink_macs(ratio[:layer], ratio[layer:], emb=384, mlp_ratio=4, seq_length=197,
                               head=6, layer=layer)
        if abs(macs - target) <= 0.02 * target:
            return ratio[:layer], ratio[layer:]
    raise RuntimeError("no policy within the MACs budget")
