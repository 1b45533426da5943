"""Drop-in nn.Modules for the reference's DeDeiT / DeViT sub-models (models/de_vit.py) and the
DeiT teacher (models/deit_vit.py), computing on the sm_100a library through the C ABI.

What is preserved from the reference:
  * registrations ``dedeit`` / ``devit`` (models/de_vit.py:495-513) and
    ``deit_base_distilled_patch16_224`` (models/deit_vit.py:477-485), constructor kwargs,
    ``forward`` / ``forward_features`` signatures and return structures;
  * the ``state_dict`` key names and ORDER (155 entries for ``dedeit``);
  * the gate protocol of core/imp_rank.py: sub-modules named ``Mlp`` / ``Attention`` expose
    ``gate``, ``hidden_features`` / ``num_heads`` and (opt-in) the ``neuron_output`` /
    ``head_output`` observers; gates are plain attributes, not buffers.
What is different: parameters stay fp32 ``nn.Parameter`` masters, but the arithmetic runs on
packed, gate-compacted operands (``packing.py``); the forward is inference-only (no autograd).
There is no CPU path: a CPU tensor input raises ``DevitError``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import weakref
from collections import OrderedDict
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import packing
from .registry import register_model

# timm.data.constants, used by the reference's transforms (data/get_dataset.py:11,108)
IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def u8_layout(x):
    """DEVIT layout id of a 4-d uint8 image batch: [B,C,H,W] (C <= 4) or [B,H,W,3]."""
    if x.dim() != 4:
        raise L.DevitError("uint8 image batches must be 4-d ([B,C,H,W] or [B,H,W,3])")
    if x.shape[2] == x.shape[3] and x.shape[1] <= 4:  # (sides are multiples of 16: no overlap)
        return L.LAYOUT_NCHW
    if x.shape[1] == x.shape[2] and x.shape[3] == 3:
        return L.LAYOUT_NHWC
    raise L.DevitError(f"uint8 batch {tuple(x.shape)} is neither [B,C,H,H] with C <= 4 nor "
                       f"[B,H,H,3]")

_PREC = {'bf16': L.DEVIT_BF16, 'fp32': L.DEVIT_FP32}


def default_precision() -> str:
    p = os.environ.get('DEVIT_PRECISION', 'bf16')
    if p not in _PREC:
        raise ValueError(f"DEVIT_PRECISION must be one of {list(_PREC)}")
    return p


def _trunc_normal_(t, std=.02):
    return nn.init.trunc_normal_(t, std=std, a=-2., b=2.)


def _opk(prec: int) -> int:
    return L.OUT_BF16 if prec == L.DEVIT_BF16 else L.OUT_F32_SPLIT


def _rows(x, prec):
    """[.., C] fp32 CUDA tensor -> ([M, C] operand, leading shape)."""
    lead = x.shape[:-1]
    return L.to_operand(x.reshape(-1, x.shape[-1]).float(), prec), lead


class PatchEmbed(nn.Module):
    """Parameter container with timm 0.5.4's PatchEmbed layout (``proj`` = Conv2d(k=s=patch)).
    The convolution itself runs as im2col + GEMM inside devit_vit_forward."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None,
                 flatten=True):
        super().__init__()
        self.img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        self.patch_size = (patch_size, patch_size) if isinstance(patch_size, int) \
            else tuple(patch_size)
        if self.patch_size != (16, 16) or self.img_size[0] != self.img_size[1]:
            raise L.DevitError("devit_b200 supports square images with 16x16 patches")
        self.grid_size = (self.img_size[0] // 16, self.img_size[1] // 16)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=16, stride=16)
        self.norm = nn.Identity()


class _Gated:
    """Gate attribute with an epoch counter so packed weights notice re-assignment
    (core/shrink_imp.py:150-171 sets and restores gates once per candidate policy), plus the
    lazily materialised observer the rank functions read after a plain ``model(data)`` call
    (core/imp_rank.py:27-31, :104-108): the fused forward never writes observers; the first
    read re-runs the last batch through the layer-wise path."""

    _observer_name = None

    def _observer(self):
        owner = getattr(self, '_owner', None)
        owner = owner() if owner is not None else None
        if owner is not None and owner._observers_stale:
            owner._materialize_observers()
        return self.__dict__.get('_observer_value')

    @property
    def gate(self):
        return self._gate

    @gate.setter
    def gate(self, value):
        if not torch.is_tensor(value):
            value = torch.as_tensor(value, dtype=torch.float32)
        object.__setattr__(self, '_gate', value)
        object.__setattr__(self, '_gate_epoch', getattr(self, '_gate_epoch', 0) + 1)


class Mlp(_Gated, nn.Module):
    """models/de_vit.py:21-47.  Stand-alone forward = dense fc1 + GELU, gate, fc2 (reference
    semantics, observer ``neuron_output`` is the post-GELU, post-mask activation)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.hidden_features = hidden_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        self.gate = torch.ones(hidden_features)
        self.precision = default_precision()

    @property
    def neuron_output(self):
        return self._observer()

    @torch.no_grad()
    def forward(self, x):
        prec = _PREC[self.precision]
        a, lead = _rows(x, prec)
        w1 = L.to_operand(self.fc1.weight.float(), prec)
        h = L.gemm(a, w1, precision=prec, bias=self.fc1.bias.float(), act=L.ACT_GELU_ERF,
                   out_kind=L.OUT_F32)
        h.mul_(self.gate.float().to(h.device).view(1, self.hidden_features))
        self.__dict__['_observer_value'] = h.view(*lead, self.hidden_features)  # B x N x hidden
        w2 = L.to_operand(self.fc2.weight.float(), prec)
        y = L.gemm(L.to_operand(h, prec), w2, precision=prec, bias=self.fc2.bias.float(),
                   out_kind=L.OUT_F32)
        return y.view(*lead, -1)


class Attention(_Gated, nn.Module):
    """models/de_vit.py:50-87.  Stand-alone forward keeps the reference semantics: all heads are
    computed, the gate multiplies the per-head outputs, ``head_output`` observes [B,N,H,hd]."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.gate = torch.ones(self.num_heads)
        self.precision = default_precision()

    @property
    def head_output(self):
        return self._observer()

    @torch.no_grad()
    def forward(self, x, output_qkv=False):
        prec = _PREC[self.precision]
        B, N, Cdim = x.shape
        H = self.num_heads
        a, _ = _rows(x, prec)
        wq = L.to_operand(self.qkv.weight.float(), prec)
        bq = None if self.qkv.bias is None else self.qkv.bias.float()
        qkv = L.gemm(a, wq, precision=prec, bias=bq, out_kind=_opk(prec))
        o = L.attention(qkv, B, N, H, self.scale, precision=prec)
        o32 = L.operand_to_f32(o, prec).view(B, N, H, Cdim // H)
        o32.mul_(self.gate.float().to(o32.device).view(1, 1, H, 1))
        self.__dict__['_observer_value'] = o32  # batch x seq x head x embed_chunk
        wp = L.to_operand(self.proj.weight.float(), prec)
        y = L.gemm(L.to_operand(o32.reshape(B * N, Cdim), prec), wp, precision=prec,
                   bias=self.proj.bias.float(), out_kind=L.OUT_F32).view(B, N, Cdim)
        outputs = {'output': y}
        if output_qkv:
            q, k, v = L.operand_to_f32(qkv, prec).view(B, N, 3, H, Cdim // H) \
                .permute(2, 0, 3, 1, 4).unbind(0)
            outputs['qkv'] = (q, k, v)
        else:
            outputs['qkv'] = None
        return outputs


class Block(nn.Module):
    """models/de_vit.py:90-121."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop,
                              proj_drop=drop)
        self.drop_path_rate = float(drop_path)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer,
                       drop=drop)

    def _ln(self, norm, x):
        B, N, Cdim = x.shape
        y = L.layernorm(x.reshape(B * N, Cdim).float().contiguous(), norm.weight.float(),
                        norm.bias.float(), norm.eps, L.OUT_F32)
        return y.view(B, N, Cdim)

    @torch.no_grad()
    def forward(self, x, output_qkv=False, output_att=False):
        att_outputs = self.attn(self._ln(self.norm1, x), output_qkv)
        x = x + att_outputs['output']
        x = x + self.mlp(self._ln(self.norm2, x))
        outputs = {'output': x}
        outputs['qkv'] = att_outputs['qkv'] if output_qkv else None
        outputs['attention'] = att_outputs['output'] if output_att else None
        return outputs


class VisionTransformer(nn.Module):
    """models/de_vit.py:124-334 (and, with ``tuple_api=True``, models/deit_vit.py:84-296)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768,
                 depth=12, num_heads=12, mlp_ratio=4., qkv_bias=True, representation_size=None,
                 distilled=False, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
                 embed_layer=PatchEmbed, norm_layer=None, act_layer=None, weight_init='',
                 resize_dim=None, tuple_api=False):
        super().__init__()
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_tokens = 2 if distilled else 1
        self.resize_dim = resize_dim
        self.tuple_api = tuple_api
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        act_layer = act_layer or nn.GELU
        if act_layer is not nn.GELU:
            raise L.DevitError("devit_b200 implements the exact-erf GELU MLP only")
        self.drop_rate, self.attn_drop_rate = float(drop_rate), float(attn_drop_rate)

        self.patch_embed = embed_layer(img_size=img_size, patch_size=patch_size,
                                       in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = nn.Parameter(torch.zeros(1, 1, embed_dim)) if distilled else None
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + self.num_tokens, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.Sequential(*[
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i],
                  norm_layer=norm_layer, act_layer=act_layer) for i in range(depth)])
        self.norm = norm_layer(embed_dim)

        if representation_size and not distilled:
            self.num_features = representation_size
            self.pre_logits = nn.Sequential(OrderedDict([
                ('fc', nn.Linear(embed_dim, representation_size)), ('act', nn.Tanh())]))
        else:
            self.pre_logits = nn.Identity()
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        self.head_dist = None
        if distilled:
            self.head_dist = nn.Linear(self.embed_dim, self.num_classes) if num_classes > 0 \
                else nn.Identity()
        if self.resize_dim is not None:
            self.resize_mlp = nn.Linear(self.embed_dim, self.resize_dim)
            self.resize_att_mlp = nn.Linear(self.embed_dim, self.resize_dim)
            self.resize_encoder_mlp = nn.Linear(self.embed_dim, self.resize_dim)

        self.precision = default_precision()
        self.input_norm = (IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD)
        self.export_qkv_layers = None  # layers whose q/k/v output_qkv=True keeps (None = all)
        self._packs = {}
        self._observers_stale = False
        self._last_input = None
        for blk in self.blocks:
            for m in (blk.attn, blk.mlp):
                object.__setattr__(m, '_owner', weakref.ref(self))
        self.init_weights(weight_init)

    # ------------------------------------------------------------------ reference API
    def init_weights(self, mode=''):
        """models/de_vit.py:205-216 (non-jax mode): trunc-normal(.02) on every Linear weight,
        zero biases, LayerNorm 1/0, conv left at the PyTorch default."""
        assert mode in ('jax', 'jax_nlhb', 'nlhb', '')
        _trunc_normal_(self.pos_embed, std=.02)
        if self.dist_token is not None:
            _trunc_normal_(self.dist_token, std=.02)
        _trunc_normal_(self.cls_token, std=.02)
        self.apply(_init_vit_weights)

    def _init_weights(self, m):
        _init_vit_weights(m)

    @torch.jit.ignore()
    def load_pretrained(self, checkpoint_path, prefix=''):
        """models/de_vit.py:223-224: Google-Brain Flax .npz checkpoint -> parameters."""
        from .npz_loader import load_npz
        load_npz(self, checkpoint_path, prefix)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token', 'dist_token'}

    def get_classifier(self):
        if self.dist_token is None:
            return self.head
        return self.head, self.head_dist

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        if self.num_tokens == 2:
            self.head_dist = nn.Linear(self.embed_dim, self.num_classes) if num_classes > 0 \
                else nn.Identity()

    def set_precision(self, precision: str):
        if precision not in _PREC:
            raise ValueError(f"precision must be one of {list(_PREC)}")
        self.precision = precision
        for m in self.modules():
            if isinstance(m, (Mlp, Attention)):
                m.precision = precision
        return self

    # ------------------------------------------------------------------ packed fast path
    def packed(self, device=None) -> packing.PackedVit:
        device = device or self.cls_token.device
        key = (self.precision, str(device))
        ver = packing.module_version(self)
        hit = self._packs.get(key)
        if hit is None or hit[0] != ver:
            fold = os.environ.get('DEVIT_FOLD_LN', '1') != '0'  # debug switch, see packing.py
            hit = (ver, packing.PackedVit(self, _PREC[self.precision], device, fold_ln=fold))
            self._packs = {key: hit}  # one live pack per model
        return hit[1]

    def set_input_norm(self, mean=IMAGENET_DEFAULT_MEAN, std=IMAGENET_DEFAULT_STD):
        """Normalisation applied on the device to uint8 inputs (ToTensor + Normalize of the
        reference's eval transform, data/get_dataset.py:107-108)."""
        self.input_norm = (tuple(float(v) for v in mean), tuple(float(v) for v in std))
        return self

    def _check_input(self, x, convert=True):
        if not x.is_cuda:
            raise L.DevitError("devit_b200 models run on CUDA (sm_100) tensors only; "
                               "there is no CPU fallback")
        if self.training and (self.drop_rate > 0 or self.attn_drop_rate > 0 or
                              any(b.drop_path_rate > 0 for b in self.blocks)):
            raise L.DevitError("devit_b200 is forward/inference only: call .eval() "
                               "(dropout / drop-path are not implemented)")
        H = self.patch_embed.img_size[0]
        if x.dtype == torch.uint8:
            nhwc = u8_layout(x) == L.LAYOUT_NHWC
            h, w = (x.shape[1], x.shape[2]) if nhwc else (x.shape[2], x.shape[3])
        else:
            h, w = x.shape[2], x.shape[3]
        assert h == H and w == H, f"Input image size ({h}*{w}) doesn't match model ({H}*{H})."
        if x.dtype == torch.uint8 or not convert:
            return x
        return x.float().contiguous()

    def patches_of(self, x):
        """Token-row patch matrix of a batch in this model's operand format: fp32 NCHW images as
        they are, uint8 images ([B,C,H,W] or [B,H,W,3]) normalised on the device."""
        prec = _PREC[self.precision]
        if x.dtype == torch.uint8:
            mean, std = self.input_norm
            return L.im2col_tokens_u8(x, mean, std, self.num_tokens, prec, u8_layout(x))
        return L.im2col_tokens(x.float().contiguous(), self.num_tokens, prec)

    @torch.no_grad()
    def features_into(self, x, feats_f32=None, feats_op=None, x_out=None, num_layers=-1,
                      patches=None, rows=None):
        """Fused forward of the compacted sub-model: images -> LayerNormed cls(/dist) rows,
        written into caller-provided slabs ([num_tokens, B, D] fp32 and/or operand format).
        `patches` (optional): the token-row patch matrix of x from ``patches_of`` -- the
        sub-models of an ensemble embed the same images, so MultiViT extracts it once.
        `rows` = (b0, b1) (optional): process only images [b0, b1) of x / patches and write
        their rows of the slabs (MultiViT runs the chunks of a batch as independent kernel chains
        on separate streams).  uint8 images are normalised on the device (``set_input_norm``)."""
        x = self._check_input(x, convert=patches is None)
        # the C side launches on the CURRENT device: make the operand's device current for the
        # call (a model on cuda:1 while cuda:0 is current must not launch on device 0's stream)
        with torch.cuda.device(x.device):
            pk = self.packed(x.device)
            B = x.shape[0]
            if patches is None and x.dtype == torch.uint8:
                patches = self.patches_of(x)
            b0, b1 = (0, B) if rows is None else rows
            if not (0 <= b0 < b1 <= B):
                raise L.DevitError(f"features_into: rows {rows} outside the batch of {B}")
            nb = b1 - b0
            ws = packing.workspace(x.device, pk.workspace_bytes(nb))
            T, D = self.num_tokens, self.embed_dim
            split = feats_op is not None and feats_op.dim() == 4
            plane = feats_op.stride(0) if split else 0
            for t in (feats_f32, feats_op[0] if split else feats_op):
                if t is not None and (t.shape != (T, B, D) or not t.is_contiguous()):
                    raise L.DevitError(f"features_into: feature slabs must be contiguous "
                                       f"[{T}, {B}, {D}], got {tuple(t.shape)}")
            ex = L.VitExports()
            ex.feats_kind_rows = B

            def at(t, elems):  # pointer `elems` elements into a tensor (or None)
                return None if t is None else t.data_ptr() + elems * t.element_size()

            if x_out is not None and rows is not None:
                x_out = x_out[b0:b1]
            img_ptr = pat_ptr = None
            pplane = 0
            if patches is None:
                img_ptr = at(x, b0 * x.stride(0))
            else:
                pplane = patches.stride(0) if patches.dim() == 3 else 0
                pat_ptr = at(patches, b0 * pk.tokens * patches.shape[-1])
            L.check(L.load().devit_vit_forward_ex(
                C.byref(pk.desc), img_ptr, pat_ptr, pplane, nb, ws.data_ptr(), ws.numel(),
                at(feats_f32, b0 * D), at(feats_op, b0 * D), plane, L.ptr(x_out), num_layers,
                C.byref(ex), L.stream_ptr(x.device)))
        # the observers (neuron_output / head_output, read by core/imp_rank.py) now belong to an
        # older batch: mark them stale so the next read re-runs THIS batch through the layer-wise
        # path.  `_last_input` aliases the caller's tensor (no copy of a 154 MB batch): a caller
        # that overwrites the buffer before reading an observer ranks the new contents.
        if rows is None:
            self._last_input, self._observers_stale = x, True

    def _feature_slabs(self, B, device, want_op=False):
        f32 = torch.empty(self.num_tokens, B, self.embed_dim, device=device)
        op = None
        if want_op:
            if _PREC[self.precision] == L.DEVIT_BF16:
                op = torch.empty(self.num_tokens, B, self.embed_dim, device=device,
                                 dtype=torch.bfloat16)
            else:
                op = torch.empty(2, self.num_tokens, B, self.embed_dim, device=device)
        return f32, op

    # ------------------------------------------------------------------ forward_features
    @torch.no_grad()
    def forward_features(self, x, output_qkv=False, output_att=False, output_emb=False,
                         output_encoders=False):
        depth = len(self.blocks)
        need_layers = output_qkv or output_att or output_emb or output_encoders or \
            self.resize_dim is not None
        if not need_layers:
            f32, _ = self._feature_slabs(x.shape[0], x.device)
            self.features_into(x, feats_f32=f32)
            self._last_input, self._observers_stale = x, True
            tokens = self._pre_logits(f32[0]) if self.dist_token is None else (f32[0], f32[1])
            if self.tuple_api:
                return (tokens, [], [], [])
            return {'output': tokens, 'qkv': [None] * depth, 'attention': [None] * depth,
                    'encoder': [None] * depth}
        if output_qkv and not (output_att or output_emb or output_encoders) and \
                self.resize_dim is None and x.is_cuda and self._all_heads_kept(x.device):
            return self._forward_features_qkv_export(x)
        return self._forward_features_layerwise(x, output_qkv, output_att, output_emb,
                                                output_encoders)

    # ------------------------------------------------------------------ fused path + q/k/v export
    def _all_heads_kept(self, device):
        pk = self.packed(device)
        return all(int(k.numel()) == blk.attn.num_heads
                   for k, blk in zip(pk.kept_heads, self.blocks))

    def _forward_features_qkv_export(self, x):
        """output_qkv=True on the FUSED path: every selected layer's QKV
        GEMM writes into an export buffer that the attention kernel reads back, so the per-layer
        (q, k, v) tuples the training-side consumers read (the teacher's q/k/v in
        feature_relation_loss, engine.py:70-95) come out of the same 10 ms forward instead of the
        layer-wise path.  `export_qkv_layers` (None = all layers) limits which layers are kept;
        the others are None.  Used when no head is gated off (q/k/v of gated heads do not exist
        in a compacted model).  q, k, v: [B, H, N, hd] views like the reference's
        (models/de_vit.py:67-68), fp32 in both modes (bf16 values widened / hi + lo), the dtype the
        layer-wise path returns."""
        x = self._check_input(x)
        B = x.shape[0]
        pk = self.packed(x.device)
        depth = len(self.blocks)
        H = self.blocks[0].attn.num_heads
        prec = _PREC[self.precision]
        want = range(depth) if self.export_qkv_layers is None else \
            sorted({int(l) % depth for l in self.export_qkv_layers})
        ex = L.VitExports()
        bufs = {}
        for l in want:
            shape = (B * pk.tokens, 3 * H * 64)
            bufs[l] = torch.empty(shape, device=x.device, dtype=torch.bfloat16) \
                if prec == L.DEVIT_BF16 else torch.empty((2,) + shape, device=x.device)
            ex.qkv[l] = bufs[l].data_ptr()
        f32, _ = self._feature_slabs(B, x.device)
        patches = self.patches_of(x) if x.dtype == torch.uint8 else None
        ws = packing.workspace(x.device, pk.workspace_bytes(B))
        with torch.cuda.device(x.device):  # the C side launches on the current device
            L.check(L.load().devit_vit_forward_ex(
                C.byref(pk.desc), None if patches is not None else x.data_ptr(),
                None if patches is None else patches.data_ptr(),
                patches.stride(0) if (patches is not None and patches.dim() == 3) else 0,
                B, ws.data_ptr(), ws.numel(), L.ptr(f32), None, 0, None, -1, C.byref(ex),
                L.stream_ptr(x.device)))
        self._last_input, self._observers_stale = x, True
        qkvs = [None] * depth
        for l, t in bufs.items():
            # fp32 in both precision modes, like the layer-wise path (Attention.forward): the
            # dtype of q/k/v must not depend on which path served the call
            t = t.float() if prec == L.DEVIT_BF16 else t[0] + t[1]
            q, k, v = t.view(B, pk.tokens, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
            qkvs[l] = (q, k, v)
        tokens = self._pre_logits(f32[0]) if self.dist_token is None else (f32[0], f32[1])
        if self.tuple_api:  # models/deit_vit.py:232-240: one entry per layer when asked for
            return (tokens, qkvs, [], [])
        return {'output': tokens, 'qkv': qkvs, 'attention': [None] * depth,
                'encoder': [None] * depth}

    def _pre_logits(self, t):
        if isinstance(self.pre_logits, nn.Identity):
            return t
        fc = self.pre_logits.fc
        prec = _PREC[self.precision]
        y = L.gemm(L.to_operand(t, prec), L.to_operand(fc.weight.float(), prec), precision=prec,
                   bias=fc.bias.float(), out_kind=L.OUT_F32)
        return torch.tanh(y)

    def _linear(self, lin, t):
        prec = _PREC[self.precision]
        a, lead = _rows(t, prec)
        y = L.gemm(a, L.to_operand(lin.weight.float(), prec), precision=prec,
                   bias=None if lin.bias is None else lin.bias.float(), out_kind=L.OUT_F32)
        return y.view(*lead, -1)

    def _forward_features_layerwise(self, x, output_qkv, output_att, output_emb, output_encoders):
        """Reference-shaped path for the training-side consumers (q/k/v, attention and encoder
        outputs, observers): embeddings from the fused kernel with zero blocks, then one
        Block.forward per layer (dense compute + gate multiply, like the reference)."""
        x = self._check_input(x)
        self._last_input, self._observers_stale = None, False
        B = x.shape[0]
        pk = self.packed(x.device)
        xt = torch.empty(B, pk.tokens, self.embed_dim, device=x.device)
        self.features_into(x, x_out=xt, num_layers=0)
        x = xt
        emb_output = x if self.resize_dim is None else self._linear(self.resize_encoder_mlp, x)
        encoder_outputs = [emb_output] if output_emb else []
        attention_outputs, qkv_outputs = [], []
        for block in self.blocks:
            out = block(x, output_qkv=output_qkv, output_att=output_att)
            tmp_enc, tmp_qkv, tmp_att = out['output'], out['qkv'], out['attention']
            x = tmp_enc
            if self.resize_dim is not None:
                tmp_att = self._linear(self.resize_att_mlp, tmp_att) if tmp_att is not None \
                    else None
                tmp_enc = self._linear(self.resize_encoder_mlp, tmp_enc)
            if self.tuple_api:
                if output_qkv:
                    qkv_outputs.append(tmp_qkv)
                if output_att:
                    attention_outputs.append(tmp_att)
                if output_encoders:
                    encoder_outputs.append(tmp_enc)
            else:
                qkv_outputs.append(tmp_qkv)
                attention_outputs.append(tmp_att)
                encoder_outputs.append(tmp_enc if output_encoders else None)
        Bn, N, D = x.shape
        rows = x[:, :self.num_tokens].reshape(-1, D).contiguous()
        y = L.layernorm(rows, self.norm.weight.float(), self.norm.bias.float(), self.norm.eps,
                        L.OUT_F32).view(Bn, self.num_tokens, D)
        tokens = self._pre_logits(y[:, 0]) if self.dist_token is None else (y[:, 0], y[:, 1])
        if self.tuple_api:
            return (tokens, qkv_outputs, attention_outputs, encoder_outputs)
        return {'output': tokens, 'qkv': qkv_outputs, 'attention': attention_outputs,
                'encoder': encoder_outputs}

    def _materialize_observers(self):
        x, self._observers_stale = self._last_input, False
        if x is not None:
            self._forward_features_layerwise(x, False, False, False, False)

    # ------------------------------------------------------------------ heads
    def _heads(self, tokens):
        """head / head_dist as GEMMs (models/de_vit.py:317)."""
        if self.head_dist is not None:
            cls, dist = tokens
            x = self._linear(self.head, cls) if isinstance(self.head, nn.Linear) else cls
            x_dist = self._linear(self.head_dist, dist) if isinstance(self.head_dist, nn.Linear) \
                else dist
            return x, x_dist
        return (self._linear(self.head, tokens) if isinstance(self.head, nn.Linear) else tokens), None

    def _heads_eval_avg(self, tokens):
        """Eval logits (x + x_dist) / 2 (models/de_vit.py:323): the head_dist GEMM's epilogue
        adds the head logits as the residual and scales by 0.5."""
        if not (isinstance(self.head, nn.Linear) and isinstance(self.head_dist, nn.Linear)):
            x, x_dist = self._heads(tokens)
            return (x + x_dist) / 2
        prec = _PREC[self.precision]
        cls, dist = tokens
        x = self._linear(self.head, cls)
        return L.gemm(L.to_operand(dist, prec), L.to_operand(self.head_dist.weight.float(), prec),
                      precision=prec, bias=self.head_dist.bias.float(), resid=x, alpha=0.5,
                      out_kind=L.OUT_F32)

    @torch.no_grad()
    def forward(self, x, distill_token=False, output_qkv=False, output_att=False,
                output_emb=False, output_encoders=False, output_tokens=False):
        if self.tuple_api:
            return self._forward_tuple(x, distill_token, output_qkv, output_att, output_emb,
                                       output_encoders, output_tokens)
        outputs = self.forward_features(x, output_qkv=output_qkv, output_att=output_att,
                                        output_emb=output_emb, output_encoders=output_encoders)
        tokens = outputs['output']
        last_tokens = tokens
        if self.resize_dim is not None:
            last_tokens = self._linear(self.resize_mlp, torch.stack(tokens, 0)
                                       if isinstance(tokens, tuple) else tokens)
        any_flag = distill_token or output_qkv or output_att or output_emb or output_encoders
        if self.head_dist is not None:
            if not self.training and not any_flag:
                return self._heads_eval_avg(tokens)
            x, x_dist = self._heads(tokens)
            outputs['output'] = (x, x_dist) if self.training else (x + x_dist) / 2
            outputs['last_tokens'] = last_tokens if distill_token else None
            if any_flag:
                return outputs
            if not self.training:
                return outputs['output']
            return x, x_dist
        x, _ = self._heads(tokens)
        outputs['output'] = x
        outputs['last_tokens'] = last_tokens if distill_token else None
        return outputs if any_flag else x

    def _forward_tuple(self, x, distill_last_cls_token, output_qkv, output_att, output_emb,
                       output_encoders, output_tokens):
        """models/deit_vit.py:251-296."""
        backbone_outputs = self.forward_features(x, output_qkv=output_qkv, output_att=output_att,
                                                 output_emb=output_emb,
                                                 output_encoders=output_encoders)
        if output_tokens and output_encoders:
            cls_tokens = [e[:, 0] for e in backbone_outputs[-1]]
            backbone_outputs += (cls_tokens,)
        tokens = backbone_outputs[0]
        last_tokens = tokens
        if self.head_dist is not None:
            x, x_dist = self._heads(tokens)
            if distill_last_cls_token:
                return last_tokens, x, x_dist
            if self.training:
                return x, x_dist
            return (x + x_dist) / 2
        x, _ = self._heads(tokens)
        outputs = (x,) + backbone_outputs[1:]
        if distill_last_cls_token:
            if self.resize_dim is not None:
                last_tokens = self._linear(self.resize_mlp, last_tokens)
            return outputs + (last_tokens,)
        if output_qkv or output_att or output_emb or output_encoders or output_tokens:
            return outputs
        return x


def _init_vit_weights(module: nn.Module, name: str = '', head_bias: float = 0.,
                      jax_impl: bool = False):
    """models/de_vit.py:337-369, non-jax branch as reached through ``self.apply``."""
    if isinstance(module, nn.Linear):
        if name.startswith('head'):
            nn.init.zeros_(module.weight)
            nn.init.constant_(module.bias, head_bias)
        else:
            _trunc_normal_(module.weight, std=.02)
            if module.bias is not None:
                nn.init.zeros_(module.bias)
    elif isinstance(module, (nn.LayerNorm, nn.GroupNorm, nn.BatchNorm2d)):
        nn.init.zeros_(module.bias)
        nn.init.ones_(module.weight)


def resize_pos_embed(posemb, posemb_new, num_tokens=1, gs_new=()):
    """models/de_vit.py:452-473: bicubic re-grid of the position table when loading a
    checkpoint trained at another resolution (load-time host work)."""
    ntok_new = posemb_new.shape[1]
    if num_tokens:
        posemb_tok, posemb_grid = posemb[:, :num_tokens], posemb[0, num_tokens:]
        ntok_new -= num_tokens
    else:
        posemb_tok, posemb_grid = posemb[:, :0], posemb[0]
    gs_old = int(math.sqrt(len(posemb_grid)))
    if not len(gs_new):
        gs_new = [int(math.sqrt(ntok_new))] * 2
    posemb_grid = posemb_grid.reshape(1, gs_old, gs_old, -1).permute(0, 3, 1, 2)
    posemb_grid = F.interpolate(posemb_grid, size=gs_new, mode='bicubic', align_corners=False)
    posemb_grid = posemb_grid.permute(0, 2, 3, 1).reshape(1, gs_new[0] * gs_new[1], -1)
    return torch.cat([posemb_tok, posemb_grid], dim=1)


def checkpoint_filter_fn(state_dict, model):
    """models/de_vit.py:476-492."""
    out_dict = {}
    if 'model' in state_dict:
        state_dict = state_dict['model']
    for k, v in state_dict.items():
        if 'patch_embed.proj.weight' in k and len(v.shape) < 4:
            O, I, H, W = model.patch_embed.proj.weight.shape
            v = v.reshape(O, -1, H, W)
        elif k == 'pos_embed' and v.shape != model.pos_embed.shape:
            v = resize_pos_embed(v, model.pos_embed, getattr(model, 'num_tokens', 1),
                                 model.patch_embed.grid_size)
        out_dict[k] = v
    return out_dict


def _cfg(url='', **kwargs):
    return {'url': url, 'num_classes': 1000, 'input_size': (3, 224, 224), 'pool_size': None,
            'crop_pct': .9, 'interpolation': 'bicubic', 'fixed_input_size': True,
            'mean': (0.485, 0.456, 0.406), 'std': (0.229, 0.224, 0.225),
            'first_conv': 'patch_embed.proj', 'classifier': 'head', **kwargs}


@register_model
def dedeit(pretrained=False, pretrained_path=None, **kwargs):
    """models/de_vit.py:495-503."""
    model = VisionTransformer(patch_size=16, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4,
                              qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                              distilled=True, **kwargs)
    model.default_cfg = _cfg()
    if pretrained_path is not None and pretrained:
        model.load_state_dict(torch.load(pretrained_path)['model'])
    return model


@register_model
def devit(pretrained=False, pretrained_path=None, **kwargs):
    """models/de_vit.py:506-513: `pretrained_path` is a Flax .npz archive, read by
    load_pretrained like the reference does (devit_b200/npz_loader.py); a torch checkpoint
    (.pth / .pt, optionally wrapped in {'model': ...}) is accepted as well."""
    model = VisionTransformer(patch_size=16, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4,
                              qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    model.default_cfg = _cfg()
    if pretrained_path is not None and pretrained:
        if str(pretrained_path).endswith(('.pth', '.pt')):
            sd = torch.load(pretrained_path)
            model.load_state_dict(checkpoint_filter_fn(sd, model))
        else:
            model.load_pretrained(checkpoint_path=pretrained_path)
    return model


@register_model
def deit_base_distilled_patch16_224(pretrained=False, pretrained_path=None, **kwargs):
    """models/deit_vit.py:477-485 (tuple-returning API of the teacher)."""
    model = VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4,
                              qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                              distilled=True, tuple_api=True, **kwargs)
    model.default_cfg = _cfg()
    if pretrained_path is not None and pretrained:
        sd = torch.load(pretrained_path)
        model.load_state_dict(sd['model'] if 'model' in sd else sd)
    return model
