"""CPU restatement of the reference's hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Plain functional PyTorch on CPU, dtype-generic (fp32 = the oracle, fp64 = tie-breaker), driven
by reference-layout state dicts.  Compute is MASKED-DENSE exactly like the reference: every head
and neuron is computed, then multiplied by its 0/1 gate.  Each function cites the reference
lines it follows.  Pinned against the reference's own modules by tests/test_oracle_golden.py
(fixtures from tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def patch_embed(sd, x):
    """timm 0.5.4 PatchEmbed as constructed at models/de_vit.py:166-168 and called at :258:
    Conv2d(k=16, s=16) then flatten(2).transpose(1, 2)."""
    w, b = sd['patch_embed.proj.weight'], sd['patch_embed.proj.bias']
    y = F.conv2d(x, w, b, stride=w.shape[-1])
    return y.flatten(2).transpose(1, 2)


def attention(sd, pre, x, num_heads, gate=None, return_qkv=False):
    """models/de_vit.py:65-87."""
    B, N, C = x.shape
    qkv = F.linear(x, sd[pre + 'qkv.weight'], sd[pre + 'qkv.bias'])
    qkv = qkv.reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    attn = (q @ k.transpose(-2, -1)) * ((C // num_heads) ** -0.5)
    attn = attn.softmax(dim=-1)
    o = (attn @ v).transpose(1, 2)  # B, N, H, hd
    if gate is not None:
        o = o * gate.to(o.dtype).view(1, 1, num_heads, 1)
    head_output = o
    y = F.linear(o.reshape(B, N, C), sd[pre + 'proj.weight'], sd[pre + 'proj.bias'])
    return (y, head_output, (q, k, v)) if return_qkv else (y, head_output)


def mlp(sd, pre, x, gate=None):
    """models/de_vit.py:35-47 (exact-erf GELU, gate after the activation)."""
    h = F.gelu(F.linear(x, sd[pre + 'fc1.weight'], sd[pre + 'fc1.bias']))
    if gate is not None:
        h = h * gate.to(h.dtype).view(1, 1, -1)
    return F.linear(h, sd[pre + 'fc2.weight'], sd[pre + 'fc2.bias']), h


def block(sd, i, x, num_heads, head_gate=None, neuron_gate=None, eps=1e-6):
    """models/de_vit.py:103-121 (DropPath is identity in eval)."""
    p = f'blocks.{i}.'
    C = x.shape[-1]
    a, _ = attention(sd, p + 'attn.', F.layer_norm(x, (C,), sd[p + 'norm1.weight'],
                                                   sd[p + 'norm1.bias'], eps),
                     num_heads, head_gate)
    x = x + a
    m, _ = mlp(sd, p + 'mlp.', F.layer_norm(x, (C,), sd[p + 'norm2.weight'],
                                            sd[p + 'norm2.bias'], eps), neuron_gate)
    return x + m


def embed_tokens(sd, x):
    """models/de_vit.py:258-264: patches, cls (and dist) tokens, + pos_embed."""
    t = patch_embed(sd, x)
    B = t.shape[0]
    toks = [sd['cls_token'].expand(B, -1, -1)]
    if 'dist_token' in sd:
        toks.append(sd['dist_token'].expand(B, -1, -1))
    return torch.cat(toks + [t], dim=1) + sd['pos_embed']


def forward_features(sd, x, num_heads=6, head_gates=None, neuron_gates=None, eps=1e-6,
                     return_blocks=False):
    """models/de_vit.py:242-292 -> (cls, dist) [B, D] each (cls only when not distilled)."""
    depth = 1 + max(int(k.split('.')[1]) for k in sd if k.startswith('blocks.'))
    x = embed_tokens(sd, x)
    per_block = [x]
    for i in range(depth):
        x = block(sd, i, x, num_heads, None if head_gates is None else head_gates[i],
                  None if neuron_gates is None else neuron_gates[i], eps)
        if return_blocks:
            per_block.append(x)
    x = F.layer_norm(x, (x.shape[-1],), sd['norm.weight'], sd['norm.bias'], eps)
    out = (x[:, 0], x[:, 1]) if 'dist_token' in sd else x[:, 0]
    return (out, per_block) if return_blocks else out


def forward_logits(sd, x, num_heads=6, head_gates=None, neuron_gates=None):
    """models/de_vit.py:294-334 in eval mode: (head(cls) + head_dist(dist)) / 2."""
    out = forward_features(sd, x, num_heads, head_gates, neuron_gates)
    if 'dist_token' in sd:
        a = F.linear(out[0], sd['head.weight'], sd['head.bias'])
        b = F.linear(out[1], sd['head_dist.weight'], sd['head_dist.bias'])
        return (a + b) / 2
    return F.linear(out, sd['head.weight'], sd['head.bias'])


def multivit(sds, x, num_heads=6, gates=None):
    """models/ensemble_models.py:32-40 ('deit' branch): the same x through every backbone."""
    cls, dist = [], []
    for s, sd in enumerate(sds):
        ng, hg = (None, None) if gates is None else gates[s]
        c, d = forward_features(sd, x, num_heads, hg, ng)
        cls.append(c)
        dist.append(d)
    return cls, dist


def ensmlp(esd, cls_list, dist_list):
    """models/ensemble_models.py:65-90 ('deit' branch, eval) -> (logits, (cls_tok, dist_tok))."""
    B = cls_list[0].shape[0]
    c = torch.stack(cls_list, 1).view(B, -1)
    d = torch.stack(dist_list, 1).view(B, -1)
    if 'cls_mlp.weight' in esd:
        c = F.linear(c, esd['cls_mlp.weight'], esd['cls_mlp.bias'])
        d = F.linear(d, esd['dist_mlp.weight'], esd['dist_mlp.bias'])
    cl = F.linear(c, esd['cls_classifier.weight'], esd['cls_classifier.bias'])
    dl = F.linear(d, esd['dist_classifier.weight'], esd['dist_classifier.bias'])
    return (cl + dl) / 2, (c, d)


def ensemble_logits(sds, esd, x, gates=None, num_heads=6):
    cls, dist = multivit(sds, x, num_heads, gates)
    logits, _ = ensmlp(esd, cls, dist)
    return logits, cls, dist


# ----------------------------------------------------------------------------- integer work
def keep_mask(width, ratio, rank):
    """core/imp_rank.py:55-58 and :137-140: num_keep = int(width * (1 - ratio)); the kept units
    are the LAST num_keep entries of the ascending-argsort rank, i.e. rank[::-1][:num_keep]."""
    num_keep = int(width * (1 - ratio))
    kept = np.asarray(rank)[::-1][:num_keep]
    mask = np.zeros(width, dtype=np.float32)
    mask[kept] = 1
    return mask


def kept_indices(mask):
    """Ascending indices of the non-zero gate entries (what compaction must select)."""
    return np.nonzero(np.asarray(mask) != 0)[0]


def shrink_macs(neuron_sparsity, head_sparsity, emb=384, seq_length=197, mlp_ratio=4, head=6,
                layer=12, num_class=1000):
    """core/compute_metric.py:31-69 (cal_shrink_flops / 2), in GMACs."""
    head_dim = emb / head
    flops = 2 * 3 * emb * 224 ** 2
    for n_s, h_s in zip(neuron_sparsity, head_sparsity):
        sa = 3 * 2 * seq_length * emb * head_dim + 4 * head_dim * seq_length ** 2
        kept = int((1 - h_s) * head)
        hidden = int(mlp_ratio * (1 - n_s) * emb)
        flops += sa * kept + seq_length * 2 * head_dim * kept * emb + 4 * seq_length * hidden * emb
    flops += 2 * emb * num_class
    return flops / 1e9 / 2


def to_dtype(sd, dtype):
    return {k: v.to(dtype) for k, v in sd.items()}


# ----------------------------------------------------------------------------- either side of
# the forward path: the eval transform's tensor part and the eval tail
def to_tensor_normalize(images_u8, mean, std, layout='nchw'):
    """torchvision ToTensor + Normalize as composed at data/get_dataset.py:107-108, on an
    already decoded / resized uint8 batch: ToTensor = HWC->CHW, `.to(float32).div(255)`;
    Normalize = `.sub_(mean[:,None,None]).div_(std[:,None,None])` with fp32 mean / std.
    `layout`: 'nchw' [B,C,H,W] or 'nhwc' [B,H,W,C].  Returns fp32 [B,C,H,W]."""
    x = images_u8 if layout == 'nchw' else images_u8.permute(0, 3, 1, 2)
    x = x.to(torch.float32).div(255)
    m = torch.as_tensor(mean, dtype=torch.float32).view(1, -1, 1, 1)
    s = torch.as_tensor(std, dtype=torch.float32).view(1, -1, 1, 1)
    return x.sub(m).div(s).contiguous()


def eval_tail(logits, target, topk=5):
    """One batch of engine.py:33-36 / 229-234: CrossEntropyLoss (mean) and timm 0.5.4
    accuracy(output, target, topk=(1, topk)) as COUNTS.  A sample is correct@k when fewer than
    k classes sort before its target (larger logit; equal logit and smaller index).
    Returns (mean loss as float, #correct@1, #correct@k)."""
    logits = logits.to(torch.float32)
    loss = F.cross_entropy(logits, target).item()
    lt = logits.gather(1, target.view(-1, 1))
    idx = torch.arange(logits.shape[1]).view(1, -1)
    ahead = ((logits > lt) | ((logits == lt) & (idx < target.view(-1, 1)))).sum(1)
    k = min(topk, logits.shape[1])
    return loss, int((ahead < 1).sum()), int((ahead < k).sum())


def eval_epoch(batches, topk=5):
    """The dict engine.evaluate returns (engine.py:39-45) for `batches` = [(logits, target), ...]:
    MetricLogger semantics (utils/dist_utils.py:30-33,59-60) -- `loss` is updated with n = 1 per
    batch, `acc1` / `acc5` (percent of the batch) with n = batch size."""
    loss_total, n_batches, c1, ck, n = 0.0, 0, 0, 0, 0
    for logits, target in batches:
        l, a, b = eval_tail(logits, target, topk)
        loss_total += l
        n_batches += 1
        c1 += a
        ck += b
        n += logits.shape[0]
    return {'loss': loss_total / n_batches, 'acc1': 100.0 * c1 / n, 'acc5': 100.0 * ck / n}
