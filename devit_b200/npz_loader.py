"""Loader for the Google-Brain Flax ``.npz`` ViT checkpoints, the format the reference's `devit`
entrypoint reads with ``model.load_pretrained(checkpoint_path)`` (models/de_vit.py:223-224, the
array-by-array copy at :372-449).  Load-time host work only.

The archive is a flat dict of arrays named after the Flax module tree.  Layout conventions of that
format, which the table below undoes:
  * Dense kernels are [in, out]; conv kernels [H, W, in, out]           -> torch [out, in(, H, W)]
  * attention query / key / value kernels are [D, heads, head_dim], their biases [heads, head_dim];
    torch keeps one fused qkv Linear with rows ordered [q | k | v][head][head_dim]
  * the attention output kernel is [heads, head_dim, D]                  -> [D, heads * head_dim]
  * LayerNorm parameters are called scale / bias
  * a few archives store vectors as [1, 1, 1, C]
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L


def _arr(z, key):
    a = np.asarray(z[key])
    if a.ndim == 4 and a.shape[:3] == (1, 1, 1):
        a = a.reshape(-1)
    return a


def _dense(a):   # [in, out] -> [out, in]
    return torch.from_numpy(np.ascontiguousarray(a.T))


def _conv(a):    # [H, W, in, out] -> [out, in, H, W]
    return torch.from_numpy(np.ascontiguousarray(a.transpose(3, 2, 0, 1)))


def _vec(a):
    return torch.from_numpy(np.ascontiguousarray(a.reshape(-1)))


def _fit_input_channels(w, in_chans):
    """Patch-conv weight [O, I, H, W] for a model with `in_chans` input channels (timm's
    adapt_input_conv convention: grey = sum over RGB, otherwise tile the RGB filters and rescale)."""
    o, i, h, wd = w.shape
    if i == in_chans:
        return w
    if in_chans == 1:
        return w.sum(1, keepdim=True)
    if i != 3:
        raise L.DevitError(f"cannot adapt a {i}-channel patch embedding to {in_chans} channels")
    reps = -(-in_chans // 3)
    return w.repeat(1, reps, 1, 1)[:, :in_chans] * (3.0 / in_chans)


@torch.no_grad()
def load_npz(model, checkpoint_path: str, prefix: str = '') -> None:
    """Copies a Flax ViT archive into `model` (a devit_b200 VisionTransformer) in place."""
    from .models import resize_pos_embed
    z = np.load(checkpoint_path)
    if not prefix and 'opt/target/embedding/kernel' in z:
        prefix = 'opt/target/'
    if hasattr(model.patch_embed, 'backbone'):
        raise L.DevitError("hybrid (ResNet stem) checkpoints are not supported by devit_b200")

    def get(name):
        return _arr(z, prefix + name)

    proj = model.patch_embed.proj
    proj.weight.copy_(_fit_input_channels(_conv(get('embedding/kernel')), proj.weight.shape[1]))
    proj.bias.copy_(_vec(get('embedding/bias')))
    model.cls_token.copy_(torch.from_numpy(get('cls')).reshape(model.cls_token.shape))
    pos = torch.from_numpy(get('Transformer/posembed_input/pos_embedding'))
    if pos.shape != model.pos_embed.shape:  # other resolution: bicubic re-grid of the patch part
        pos = resize_pos_embed(pos, model.pos_embed, getattr(model, 'num_tokens', 1),
                               model.patch_embed.grid_size)
    model.pos_embed.copy_(pos)
    model.norm.weight.copy_(_vec(get('Transformer/encoder_norm/scale')))
    model.norm.bias.copy_(_vec(get('Transformer/encoder_norm/bias')))
    head = model.head
    if isinstance(head, torch.nn.Linear) and head.bias.shape[0] == get('head/bias').shape[-1]:
        head.weight.copy_(_dense(get('head/kernel')))
        head.bias.copy_(_vec(get('head/bias')))
    fc = getattr(model.pre_logits, 'fc', None)
    if isinstance(fc, torch.nn.Linear) and prefix + 'pre_logits/bias' in z:
        fc.weight.copy_(_dense(get('pre_logits/kernel')))
        fc.bias.copy_(_vec(get('pre_logits/bias')))
    for i, blk in enumerate(model.blocks.children()):
        root = f'Transformer/encoderblock_{i}/'
        att = root + 'MultiHeadDotProductAttention_1/'
        blk.norm1.weight.copy_(_vec(get(root + 'LayerNorm_0/scale')))
        blk.norm1.bias.copy_(_vec(get(root + 'LayerNorm_0/bias')))
        blk.norm2.weight.copy_(_vec(get(root + 'LayerNorm_2/scale')))
        blk.norm2.bias.copy_(_vec(get(root + 'LayerNorm_2/bias')))
        d = blk.attn.qkv.weight.shape[1]
        # [D, heads, head_dim] -> [D, D] -> Linear rows; q, k, v stacked
        blk.attn.qkv.weight.copy_(torch.cat(
            [_dense(get(att + n + '/kernel').reshape(d, -1)) for n in ('query', 'key', 'value')]))
        blk.attn.qkv.bias.copy_(torch.cat(
            [_vec(get(att + n + '/bias')) for n in ('query', 'key', 'value')]))
        blk.attn.proj.weight.copy_(_dense(get(att + 'out/kernel').reshape(-1, d)))
        blk.attn.proj.bias.copy_(_vec(get(att + 'out/bias')))
        for r, lin in enumerate((blk.mlp.fc1, blk.mlp.fc2)):
            lin.weight.copy_(_dense(get(root + f'MlpBlock_3/Dense_{r}/kernel')))
            lin.bias.copy_(_vec(get(root + f'MlpBlock_3/Dense_{r}/bias')))


def state_dict_to_npz(sd, path: str, num_heads: int, prefix: str = '') -> None:
    """The inverse mapping (torch state_dict of a non-distilled ViT -> Flax archive): used by the
    tests to build fixtures, and handy for exporting."""
    out = {}

    def put(name, t):
        out[prefix + name] = t.detach().cpu().numpy()

    put('embedding/kernel', sd['patch_embed.proj.weight'].permute(2, 3, 1, 0))
    put('embedding/bias', sd['patch_embed.proj.bias'])
    put('cls', sd['cls_token'])
    put('Transformer/posembed_input/pos_embedding', sd['pos_embed'])
    put('Transformer/encoder_norm/scale', sd['norm.weight'])
    put('Transformer/encoder_norm/bias', sd['norm.bias'])
    if 'head.weight' in sd:
        put('head/kernel', sd['head.weight'].t())
        put('head/bias', sd['head.bias'])
    depth = 1 + max(int(k.split('.')[1]) for k in sd if k.startswith('blocks.'))
    for i in range(depth):
        b, root = f'blocks.{i}.', f'Transformer/encoderblock_{i}/'
        att = root + 'MultiHeadDotProductAttention_1/'
        d = sd[b + 'attn.qkv.weight'].shape[1]
        hd = d // num_heads
        put(root + 'LayerNorm_0/scale', sd[b + 'norm1.weight'])
        put(root + 'LayerNorm_0/bias', sd[b + 'norm1.bias'])
        put(root + 'LayerNorm_2/scale', sd[b + 'norm2.weight'])
        put(root + 'LayerNorm_2/bias', sd[b + 'norm2.bias'])
        for j, n in enumerate(('query', 'key', 'value')):
            put(att + n + '/kernel', sd[b + 'attn.qkv.weight'][j * d:(j + 1) * d].t()
                .reshape(d, num_heads, hd))
            put(att + n + '/bias', sd[b + 'attn.qkv.bias'][j * d:(j + 1) * d].reshape(num_heads, hd))
        put(att + 'out/kernel', sd[b + 'attn.proj.weight'].t().reshape(num_heads, hd, d))
        put(att + 'out/bias', sd[b + 'attn.proj.bias'])
        for r, n in enumerate(('fc1', 'fc2')):
            put(root + f'MlpBlock_3/Dense_{r}/kernel', sd[b + f'mlp.{n}.weight'].t())
            put(root + f'MlpBlock_3/Dense_{r}/bias', sd[b + f'mlp.{n}.bias'])
    np.savez(path, **out)
