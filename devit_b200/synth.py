"""Deterministic synthetic weights, inputs and gates.

There are no datasets or checkpoints in the build environment, so every test / bench / golden
fixture regenerates the same tensors from seeds with the CPU generator (bit-stable for a given
torch version).  State dicts use the reference's key names AND order (models/de_vit.py state_dict:
155 entries for `dedeit`), because ensemble.py:228-238 loads sub-checkpoints by position.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from . import shrink


def vit_keys(depth=12, distilled=True, with_heads=True):
    keys = ['cls_token'] + (['dist_token'] if distilled else []) + \
        ['pos_embed', 'patch_embed.proj.weight', 'patch_embed.proj.bias']
    for i in range(depth):
        for sub in ('norm1', 'attn.qkv', 'attn.proj', 'norm2', 'mlp.fc1', 'mlp.fc2'):
            keys += [f'blocks.{i}.{sub}.weight', f'blocks.{i}.{sub}.bias']
    keys += ['norm.weight', 'norm.bias']
    if with_heads:
        keys += ['head.weight', 'head.bias']
        if distilled:
            keys += ['head_dist.weight', 'head_dist.bias']
    return keys


def vit_shapes(dim=384, depth=12, mlp_ratio=4, num_classes=100, img=224, patch=16, chans=3,
               distilled=True, with_heads=True):
    tokens = (img // patch) ** 2 + (2 if distilled else 1)
    hidden = int(dim * mlp_ratio)
    shp = OrderedDict()
    for k in vit_keys(depth, distilled, with_heads):
        if k in ('cls_token', 'dist_token'):
            shp[k] = (1, 1, dim)
        elif k == 'pos_embed':
            shp[k] = (1, tokens, dim)
        elif k == 'patch_embed.proj.weight':
            shp[k] = (dim, chans, patch, patch)
        elif k.endswith('attn.qkv.weight'):
            shp[k] = (3 * dim, dim)
        elif k.endswith('attn.qkv.bias'):
            shp[k] = (3 * dim,)
        elif k.endswith('mlp.fc1.weight'):
            shp[k] = (hidden, dim)
        elif k.endswith('mlp.fc1.bias'):
            shp[k] = (hidden,)
        elif k.endswith('mlp.fc2.weight'):
            shp[k] = (dim, hidden)
        elif k.startswith('head') and k.endswith('weight'):
            shp[k] = (num_classes, dim)
        elif k.startswith('head') and k.endswith('bias'):
            shp[k] = (num_classes,)
        elif k.endswith('attn.proj.weight'):
            shp[k] = (dim, dim)
        else:  # every remaining bias / LayerNorm vector
            shp[k] = (dim,)
    return shp


def _fill(key, shape, gen, qkv_gain=1.0):
    t = torch.randn(*shape, generator=gen)
    is_norm = '.norm' in key or key.startswith('norm.')
    if is_norm and key.endswith('weight'):
        return 1.0 + 0.1 * t
    if is_norm:
        return 0.1 * t
    if key.endswith('attn.qkv.weight'):
        return 0.02 * qkv_gain * t
    return 0.02 * t


def vit_state_dict(seed, qkv_gain=1.0, **shape_kw):
    """Every parameter non-trivial (biases, LayerNorm affine, cls/dist/pos all random) so that
    every code path is live.  qkv_gain > 1 sharpens the softmax (stress variant for tests)."""
    gen = torch.Generator().manual_seed(seed)
    return OrderedDict((k, _fill(k, s, gen, qkv_gain)) for k, s in vit_shapes(**shape_kw).items())


def dedeit_state_dict(sub_idx, num_classes=25, qkv_gain=1.0, with_heads=True):
    return vit_state_dict(1000 + sub_idx, qkv_gain, dim=384, depth=12, num_classes=num_classes,
                          with_heads=with_heads)


def teacher_state_dict(num_classes=100):
    return vit_state_dict(3000, dim=768, depth=12, num_classes=num_classes)


def ensmlp_state_dict(n_sub=4, sub_size=384, teacher_size=768, num_class=100, seed=2000,
                      classifier_gain=8.0, distilled=True):
    """Key order of models/ensemble_models.py:44-63.  Classifier weights are scaled so that the
    logits are separated enough for the argmax comparison to be meaningful."""
    gen = torch.Generator().manual_seed(seed)
    k_in = n_sub * sub_size
    sd = OrderedDict()
    kinds = ['cls'] + (['dist'] if distilled else [])
    for kind in kinds:
        if teacher_size is not None:
            sd[f'{kind}_mlp.weight'] = 0.02 * torch.randn(teacher_size, k_in, generator=gen)
            sd[f'{kind}_mlp.bias'] = 0.02 * torch.randn(teacher_size, generator=gen)
            c_in = teacher_size
        else:
            c_in = k_in
        sd[f'{kind}_classifier.weight'] = 0.02 * classifier_gain * torch.randn(num_class, c_in,
                                                                              generator=gen)
        sd[f'{kind}_classifier.bias'] = 0.02 * torch.randn(num_class, generator=gen)
    return sd


def images(batch, seed=1234, img=224, chans=3):
    gen = torch.Generator().manual_seed(seed)
    return torch.randn(batch, chans, img, img, generator=gen)


def images_u8(batch, seed=1234, img=224, chans=3, nhwc=False):
    """Decoded-image stand-in: uniform uint8 pixels, [B,C,H,W] (or [B,H,W,C])."""
    gen = torch.Generator().manual_seed(seed + 77)
    x = torch.randint(0, 256, (batch, img, img, chans), generator=gen, dtype=torch.uint8)
    return x.contiguous() if nhwc else x.permute(0, 3, 1, 2).contiguous()


def eval_batches(sizes=(8, 8, 5), classes=100, seed=4242, scale=3.0):
    """[(logits fp32 [b, classes], target int64 [b]), ...] with a ragged last batch; about half
    of the targets are the arg-max so that both accuracy meters are exercised."""
    gen = torch.Generator().manual_seed(seed)
    out = []
    for b in sizes:
        logits = torch.randn(b, classes, generator=gen) * scale
        target = torch.randint(0, classes, (b,), generator=gen)
        top = logits.argmax(-1)
        second = logits.topk(min(3, classes), -1).indices[:, -1]
        pick = torch.rand(b, generator=gen)
        target = torch.where(pick < 0.4, top, torch.where(pick < 0.7, second, target))
        out.append((logits, target))
    return out


def shrink_gates(sub_idx, hidden=1536, heads=6, layer=12, shrink_ratio=0.3):
    """Sampled shrink_ratio-0.3 policy + random importance ranks -> (neuron_masks, head_masks),
    each a list of `layer` fp32 0/1 tensors, via the reference's index-selection rule."""
    rng = np.random.RandomState(4321 + sub_idx)
    n_ratio, h_ratio = shrink.sample_policy(rng, shrink_ratio=shrink_ratio, layer=layer)
    n_rank = [rng.permutation(hidden) for _ in range(layer)]
    h_rank = [rng.permutation(heads) for _ in range(layer)]
    n_masks = [shrink.keep_mask(hidden, n_ratio[i], n_rank[i]) for i in range(layer)]
    h_masks = [shrink.keep_mask(heads, h_ratio[i], h_rank[i]) for i in range(layer)]
    return n_masks, h_masks


# ------------------------------------------------------------------------------- CCT
def cct_shapes(dim=256, layers=7, mlp_ratio=2, n_conv=1, tokens=256, num_classes=100,
               in_planes=64, backbone=False):
    """Ordered {key: shape} of the reference CCT state_dict (models/cct.py:38-136): tokenizer,
    then classifier (or `encoders` for backbone=True, which has no fc)."""
    pre = 'encoders.' if backbone else 'classifier.'
    chans = [3] + [in_planes] * (n_conv - 1) + [dim]
    shapes = {}
    for i in range(n_conv):
        shapes[f'tokenizer.conv_layers.{i}.0.weight'] = (chans[i + 1], chans[i], 3, 3)
    shapes[pre + 'positional_emb'] = (1, tokens, dim)
    shapes[pre + 'attention_pool.weight'] = (1, dim)
    shapes[pre + 'attention_pool.bias'] = (1,)
    f = int(dim * mlp_ratio)
    for i in range(layers):
        b = f'{pre}blocks.{i}.'
        shapes[b + 'pre_norm.weight'] = (dim,)
        shapes[b + 'pre_norm.bias'] = (dim,)
        shapes[b + 'self_attn.qkv.weight'] = (3 * dim, dim)
        shapes[b + 'self_attn.proj.weight'] = (dim, dim)
        shapes[b + 'self_attn.proj.bias'] = (dim,)
        shapes[b + 'linear1.weight'] = (f, dim)
        shapes[b + 'linear1.bias'] = (f,)
        shapes[b + 'norm1.weight'] = (dim,)
        shapes[b + 'norm1.bias'] = (dim,)
        shapes[b + 'linear2.weight'] = (dim, f)
        shapes[b + 'linear2.bias'] = (dim,)
    shapes[pre + 'norm.weight'] = (dim,)
    shapes[pre + 'norm.bias'] = (dim,)
    if not backbone:
        shapes[pre + 'fc.weight'] = (num_classes, dim)
        shapes[pre + 'fc.bias'] = (num_classes,)
    return shapes


def cct_state_dict(sub_idx, **shape_kw):
    """Seeded synthetic CCT weights with every code path live (non-zero biases, LN affine != 1/0,
    attention_pool that is not uniform)."""
    gen = torch.Generator().manual_seed(3000 + sub_idx)
    sd = {}
    for key, shape in cct_shapes(**shape_kw).items():
        if 'conv_layers' in key:
            fan_in = shape[1] * 9
            t = torch.randn(shape, generator=gen) * (2.0 / fan_in) ** 0.5
        elif key.endswith('positional_emb'):
            t = torch.randn(shape, generator=gen) * 0.2
        elif 'norm' in key and key.endswith('weight'):
            t = 1.0 + 0.1 * torch.randn(shape, generator=gen)
        elif 'norm' in key and key.endswith('bias'):
            t = 0.1 * torch.randn(shape, generator=gen)
        elif 'attention_pool.weight' in key:
            t = torch.randn(shape, generator=gen) * 0.1
        elif key.endswith('bias'):
            t = torch.randn(shape, generator=gen) * 0.02
        elif key.endswith('fc.weight'):
            t = torch.randn(shape, generator=gen) * 0.16
        else:
            t = torch.randn(shape, generator=gen) * 0.05
        sd[key] = t
    return sd


def ensemble_cct_state_dict(n_sub=4, sub_size=256, teacher_size=None, num_classes=100, seed=4000):
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    width = n_sub * sub_size
    if teacher_size is not None:
        sd['cls_mlp.weight'] = torch.randn(teacher_size, width, generator=gen) * 0.03
        sd['cls_mlp.bias'] = torch.randn(teacher_size, generator=gen) * 0.02
        width = teacher_size
    sd['cls_classifier.weight'] = torch.randn(num_classes, width, generator=gen) * 0.16
    sd['cls_classifier.bias'] = torch.randn(num_classes, generator=gen) * 0.02
    return sd


def cifar_images(batch, seed=4321):
    return torch.randn(batch, 3, 32, 32, generator=torch.Generator().manual_seed(seed))
