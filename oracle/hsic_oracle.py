"""CPU restatement of the reference's HSIC importance ranking (TEST INFRASTRUCTURE -- see
oracle/__init__.py): core/imp_rank.py `mlp_neuron_rank` (:16-47), `attn_head_rank` (:93-129) and
the HSIC estimator they call (`center` :176-180, `GaussianKernel` :183-193, `LinearKernel`
:196-201, `HSICLoss` :204-239).

Written the way the reference computes it -- ONE unit (neuron / head) at a time, Gram matrices
multiplied out and traced -- so that the batched product implementation (devit_b200/shrink.py)
is checked against an independent formulation.  Pinned against the reference's own functions by
tests/test_hsic_cpu.py (fixtures from tests/golden/make_hsic_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

SIGMAS = (1, 2, 4, 8, 16)  # core/imp_rank.py:208-212


def center(X):
    """core/imp_rank.py:176-180."""
    return X - X.mean(0, keepdim=True) - X.mean(1, keepdim=True) + X.mean()


def gaussian_kernel(x, sigma):
    """core/imp_rank.py:189-193 (squared distances from the Gram matrix)."""
    inner = x @ x.t()
    norm = torch.diag(inner)
    dist_sq = norm + norm.reshape(-1, 1) - 2 * inner
    return torch.exp(-dist_sq / (2 * sigma ** 2))


def multi_gaussian(x):
    """Mean of the five kernels, core/imp_rank.py:229-230."""
    return sum(gaussian_kernel(x, s) for s in SIGMAS) / 5


def hsic(x, y, y_kernel='linear', mean_sub=False):
    """core/imp_rank.py:220-239.  NOTE the operator precedence of the reference's mean_sub line
    (:226): x - (mean / (std + 1e-12)), not (x - mean) / std -- restated as written."""
    if mean_sub:
        x = x - x.mean(0) / (x.std(0) + 1e-12)
        y = y - y.mean(0)
    G_X = center(multi_gaussian(x))
    G_Y = center(y @ y.t()) if y_kernel == 'linear' else center(multi_gaussian(y))
    return torch.trace(G_X @ G_Y)


def _minmax(a):
    return (a - np.min(a)) / (np.max(a) - np.min(a))


def neuron_scores(neuron_output, output):
    """Per-neuron importance of one Mlp (core/imp_rank.py:30-40): 0.1 * min-max-normalised HSIC
    relevance to the softmax prediction + 0.9 * min-max-normalised sum of |activation|.
    neuron_output [B, N, F], output (logits) [B, C] -> float64 [F]."""
    y = F.softmax(output, dim=-1)
    h = np.array([hsic(neuron_output[:, :, f], y, 'linear', True).item()
                  for f in range(neuron_output.shape[-1])])
    act = np.sum(neuron_output.abs().detach().cpu().numpy(), axis=(0, 1))
    return np.array((0.1 * _minmax(h) + 0.9 * _minmax(act)).tolist())


def head_scores(head_output, output):
    """Per-head importance of one Attention (core/imp_rank.py:107-121): relevance of the head's
    channel-mean to the prediction minus 0.1 * its mean redundancy with the other heads.
    head_output [B, N, H, hd] -> float64 [H]."""
    y = F.softmax(output, dim=-1)
    xm = [head_output[:, :, h, :].mean(-1) for h in range(head_output.shape[2])]
    H = len(xm)
    out = []
    for h1 in range(H):
        rel = hsic(xm[h1], y, 'linear', True).item()
        red = sum(hsic(xm[h1], xm[h2], 'rbf', False).item() for h2 in range(H) if h2 != h1)
        out.append(rel - 0.1 * red / (H - 1))
    return np.array(out)


def rank_from_scores(scores):
    """core/imp_rank.py:46,128: ascending argsort per layer."""
    return [np.argsort(s) for s in scores]
