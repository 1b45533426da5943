"""Host-side mirror of the reference's gate selection (core/imp_rank.py, core/compute_metric.py,
core/shrink_imp.py:66-82).  Integer index work only -- it configures the hot path, it is not on
it.  The functions keep the reference's names, argument meaning and module-discovery protocol
(``'Mlp' in str(m) and 'Attention' not in str(m)``), so the reference's own core/imp_rank.py
functions and these are interchangeable on devit_b200 modules.

Also here: the HSIC importance ranking that produces the `rank` lists (core/imp_rank.py
`mlp_neuron_rank` :16-47, `attn_head_rank` :93-129, `HSICLoss` :204-239).
The reference evaluates one unit at a time -- 12 x 1536 HSIC estimates, each a dozen small torch
ops followed by an ``.item()`` host sync.  Here every unit of a layer goes through ONE batched
evaluation on the device the observers live on, with a single device->host read per layer.  It
is offline work written with torch ops (batched matmul + elementwise), not a hand-written kernel.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


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



# ------------------------------------------------------------------------------------------
# HSIC importance ranking (core/imp_rank.py:16-47, :93-129, :176-239), batched over the units
# ------------------------------------------------------------------------------------------
HSIC_SIGMAS = (1.0, 2.0, 4.0, 8.0, 16.0)  # core/imp_rank.py:208-212


def _center(g):
    """core/imp_rank.py:176-180 on the last two dims."""
    return g - g.mean(-2, keepdim=True) - g.mean(-1, keepdim=True) + g.mean((-2, -1), keepdim=True)


def _multi_gaussian(x):
    """Mean of the five Gaussian kernels of [.., B, N] samples -> [.., B, B]
    (core/imp_rank.py:189-193, :229-230): squared distances from the Gram matrix."""
    inner = x @ x.transpose(-1, -2)
    norm = torch.diagonal(inner, dim1=-2, dim2=-1)
    dist_sq = norm.unsqueeze(-2) + norm.unsqueeze(-1) - 2 * inner
    return sum(torch.exp(-dist_sq / (2 * s ** 2)) for s in HSIC_SIGMAS) / len(HSIC_SIGMAS)


def _mean_sub(x):
    """core/imp_rank.py:226 AS WRITTEN: x - (mean / (std + 1e-12)) over the batch dimension (the
    reference's operator precedence, unbiased std); x is [.., B, N]."""
    return x - x.mean(-2, keepdim=True) / (x.std(-2, keepdim=True) + 1e-12)


def _trace_of_product(gx, gy):
    """trace(gx @ gy) for batches of square matrices without forming the product."""
    return (gx * gy.transpose(-1, -2)).sum((-2, -1))


def hsic_relevance(x, prob, chunk=256):
    """HSICLoss(y_kernel='linear', mean_sub=True)(x_u, prob) of core/imp_rank.py:220-239 for every
    unit u at once.  x: [U, B, N] (unit, sample, feature), prob: [B, C] -> [U] (x's dtype)."""
    y = prob - prob.mean(0)
    gy = _center(y @ y.t())
    out = []
    for u0 in range(0, x.shape[0], chunk):  # bounds the [chunk, B, B] temporaries
        gx = _center(_multi_gaussian(_mean_sub(x[u0:u0 + chunk].contiguous())))
        out.append(_trace_of_product(gx, gy))
    return torch.cat(out)


def hsic_redundancy(x):
    """HSICLoss(y_kernel='rbf', mean_sub=False)(x_a, x_b) for every pair of units:
    x [U, B, N] -> [U, U]."""
    g = _center(_multi_gaussian(x))
    return torch.einsum('aij,bji->ab', g, g)


def neuron_scores(neuron_output, output):
    """Importance of every neuron of one Mlp (core/imp_rank.py:30-40): 0.1 * min-max-normalised
    HSIC relevance to the softmax prediction + 0.9 * min-max-normalised sum of |activation|.
    neuron_output [B, N, F], output (logits) [B, C] -> float64 numpy [F]."""
    x = neuron_output.detach().float()
    hs = hsic_relevance(x.permute(2, 0, 1), F.softmax(output.detach().float(), dim=-1))
    act = x.abs().sum((0, 1))
    hs, act = hs.double().cpu().numpy(), act.cpu().numpy()  # the layer's one host read
    hs = (hs - np.min(hs)) / (np.max(hs) - np.min(hs))
    act = (act - np.min(act)) / (np.max(act) - np.min(act))
    return np.array((0.1 * hs + 0.9 * act).tolist())


def head_scores(head_output, output):
    """Importance of every head of one Attention (core/imp_rank.py:107-121): relevance of the
    head's channel mean minus 0.1 * its mean redundancy with the other heads.
    head_output [B, N, H, hd] -> float64 numpy [H]."""
    xm = head_output.detach().float().mean(-1).permute(2, 0, 1)  # [H, B, N]
    H = xm.shape[0]
    rel = hsic_relevance(xm, F.softmax(output.detach().float(), dim=-1))
    red = hsic_redundancy(xm)
    red = (red.sum(1) - torch.diagonal(red)) / (H - 1)
    return (rel.double() - 0.1 * red.double()).cpu().numpy()


def _first_batch(model, train_loader, mode):
    """model(data) on the loader's first batch (core/imp_rank.py:21-28, :98-105 stop after it)."""
    for data, _ in train_loader:
        if mode == 'cuda':
            p = next(model.parameters(), None)
            data = data.to(p.device if p is not None and p.is_cuda else 'cuda')
        with torch.no_grad():
            return model(data)
    raise ValueError("empty train_loader")


def mlp_neuron_rank(model, train_loader, mode='cuda'):
    """core/imp_rank.py:16-47: per Mlp, the ascending argsort of the neuron scores of the
    loader's first batch.  Reads the `neuron_output` observers after a plain model(data)."""
    output = _first_batch(model, train_loader, mode)
    return [np.argsort(neuron_scores(m.neuron_output, output))
            for m in model.modules() if _is_mlp(m)]


def attn_head_rank(model, train_loader, mode='cuda'):
    """core/imp_rank.py:93-129: per Attention, the ascending argsort of the head scores."""
    output = _first_batch(model, train_loader, mode)
    return [np.argsort(head_scores(m.head_output, output))
            for m in model.modules() if _is_attn(m)]

