"""CCT (Compact Convolutional Transformer) sub-models and their ensemble on the B200 kernels.

Mirrors the reference's plugin boundary for this family: class / attribute names and the
``state_dict`` layout of models/cct.py:38-136 (``tokenizer.conv_layers.{i}.0.weight``,
``classifier.*`` or -- ``backbone=True`` -- ``encoders.*``), the registered entry points
(models/cct.py:252-455), ``get_decct`` (:461-470), ``MultiCCT`` / ``EnsembleCCT``
(models/ensemble_models.py:93-151).  Inference only; the whole forward is one C call
(``devit_cct_forward``, include/devit_b200.h) plus the fc / fusion GEMMs.

Deviation, on purpose: the reference's ``MultiCCT.forward`` indexes ``model.forward(x)[0]``
(models/ensemble_models.py:111), which on a [B, 256] backbone output keeps ONE sample's row; the
documented intent (EnsembleCCT docstring :135) is the feature of every sample, which is what
this implementation returns.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L
from . import packing
from .models import _PREC, default_precision
from .registry import register_model


class Tokenizer(nn.Module):
    """models/utils/tokenizer.py:6-44 (parameters only; the math runs inside devit_cct_forward)."""

    def __init__(self, kernel_size, stride, padding, pooling_kernel_size=3, pooling_stride=2,
                 pooling_padding=1, n_conv_layers=1, n_input_channels=3, n_output_channels=64,
                 in_planes=64, activation=None, max_pool=True, conv_bias=False):
        super().__init__()
        if (kernel_size, stride, padding) != (3, 1, 1) or not max_pool or conv_bias or \
                (pooling_kernel_size, pooling_stride, pooling_padding) != (3, 2, 1) or \
                activation is not nn.ReLU:
            raise L.DevitError("devit_b200 CCT tokenizer is built for conv 3x3 s1 p1 (no bias) + "
                               "ReLU + max-pool 3/2/1 (the decct_*_3xN family)")
        chans = [n_input_channels] + [in_planes] * (n_conv_layers - 1) + [n_output_channels]
        self.conv_layers = nn.Sequential(*[
            nn.Sequential(nn.Conv2d(chans[i], chans[i + 1], 3, 1, 1, bias=False), nn.ReLU(),
                          nn.MaxPool2d(3, 2, 1)) for i in range(n_conv_layers)])
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)

    def sequence_length(self, n_channels=3, height=224, width=224):
        n = len(self.conv_layers)
        return (height >> n) * (width >> n)


class Attention(nn.Module):
    """models/utils/transformers.py:7-35 (QKV without bias)."""

    def __init__(self, dim, num_heads=8, attention_dropout=0.1, projection_dropout=0.1):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.proj = nn.Linear(dim, dim)


class TransformerEncoderLayer(nn.Module):
    """models/utils/transformers.py:73-113."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, attention_dropout=0.1,
                 drop_path_rate=0.1):
        super().__init__()
        self.pre_norm = nn.LayerNorm(d_model)
        self.self_attn = Attention(d_model, nhead, attention_dropout, dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear2 = nn.Linear(dim_feedforward, d_model)


class TransformerClassifier(nn.Module):
    """models/utils/transformers.py:262-360 (seq_pool, learnable / sine / no positions)."""

    has_fc = True

    def __init__(self, seq_pool=True, embedding_dim=768, num_layers=12, num_heads=12,
                 mlp_ratio=4.0, num_classes=1000, dropout=0.1, attention_dropout=0.1,
                 stochastic_depth=0.1, positional_embedding='learnable', sequence_length=None):
        super().__init__()
        if not seq_pool:
            raise L.DevitError("devit_b200 CCT supports seq_pool=True only")
        positional_embedding = positional_embedding if positional_embedding in \
            ('sine', 'learnable', 'none') else 'sine'
        self.embedding_dim, self.sequence_length = embedding_dim, sequence_length
        self.num_heads = num_heads
        self.attention_pool = nn.Linear(embedding_dim, 1)
        if positional_embedding == 'learnable':
            self.positional_emb = nn.Parameter(torch.zeros(1, sequence_length, embedding_dim))
            nn.init.trunc_normal_(self.positional_emb, std=0.2)
        elif positional_embedding == 'sine':
            self.positional_emb = nn.Parameter(self.sinusoidal_embedding(sequence_length,
                                                                         embedding_dim),
                                               requires_grad=False)
        else:
            self.positional_emb = None
        ff = int(embedding_dim * mlp_ratio)
        self.blocks = nn.ModuleList([TransformerEncoderLayer(embedding_dim, num_heads, ff)
                                     for _ in range(num_layers)])
        self.norm = nn.LayerNorm(embedding_dim)
        if self.has_fc:
            self.fc = nn.Linear(embedding_dim, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    @staticmethod
    def sinusoidal_embedding(n_channels, dim):
        pe = torch.FloatTensor([[p / (10000 ** (2 * (i // 2) / dim)) for i in range(dim)]
                                for p in range(n_channels)])
        pe[:, 0::2] = torch.sin(pe[:, 0::2])
        pe[:, 1::2] = torch.cos(pe[:, 1::2])
        return pe.unsqueeze(0)


class CCTTransformer(TransformerClassifier):
    """models/utils/transformers.py:386-477: the backbone flavour (no fc, returns the pooled
    feature)."""

    has_fc = False


class PackedCCT:
    """Device-resident packed weights + the ctypes descriptor of devit_cct_forward."""

    def __init__(self, model, precision: int, device: torch.device):
        self._keep = []
        enc = model.encoders if model.backbone else model.classifier
        dim, depth = enc.embedding_dim, len(enc.blocks)

        def op(t):
            t = L.to_operand(t.detach().to(device=device, dtype=torch.float32), precision)
            self._keep.append(t)
            return t.data_ptr()

        def f32(t):
            t = t.detach().to(device=device, dtype=torch.float32).contiguous()
            self._keep.append(t)
            return t.data_ptr()

        fold = precision == L.DEVIT_BF16 and dim % 128 == 0 and dim <= 768
        self.layers = (L.LayerDesc * depth)()
        for i, blk in enumerate(enc.blocks):
            d = self.layers[i]
            ff = blk.linear1.out_features
            if ff % 16:
                raise L.DevitError(f"CCT feed-forward width {ff} must be a multiple of 16")
            d.heads, d.hidden, d.hidden_ld = blk.self_attn.num_heads, ff, ff
            d.ln1_g, d.ln1_b = f32(blk.pre_norm.weight), f32(blk.pre_norm.bias)
            d.ln2_g, d.ln2_b = f32(blk.norm1.weight), f32(blk.norm1.bias)
            wq = blk.self_attn.qkv.weight.detach().float().cpu()
            bq = torch.zeros(wq.shape[0])
            w1 = blk.linear1.weight.detach().float().cpu()
            b1 = blk.linear1.bias.detach().float().cpu()
            if fold:
                wq, d.cs_qkv, bq = packing.PackedVit._fold(wq, bq, blk.pre_norm, f32)
                w1, d.cs_fc1, b1 = packing.PackedVit._fold(w1, b1, blk.norm1, f32)
            d.w_qkv, d.b_qkv = op(wq), f32(bq)
            d.w_proj, d.b_proj = op(blk.self_attn.proj.weight), f32(blk.self_attn.proj.bias)
            d.w_fc1, d.b_fc1 = op(w1), f32(b1)
            d.w_fc2, d.b_fc2 = op(blk.linear2.weight), f32(blk.linear2.bias)
        desc = L.CctDesc()
        desc.precision, desc.dim, desc.depth = precision, dim, depth
        convs = [seq[0] for seq in model.tokenizer.conv_layers]
        desc.img, desc.chans, desc.n_conv = model.img_size, convs[0].in_channels, len(convs)
        for i, cv in enumerate(convs):
            cout, cin = cv.out_channels, cv.in_channels
            kpad = (9 * cin + 7) // 8 * 8
            w = torch.zeros(cout, kpad)
            # K order (ky, kx, c_in): matches devit_im2col3x3
            w[:, :9 * cin] = cv.weight.detach().float().cpu().permute(0, 2, 3, 1).reshape(cout, -1)
            desc.conv_chans[i], desc.conv_kpad[i], desc.w_conv[i] = cout, kpad, op(w)
        desc.pos = None if enc.positional_emb is None else f32(enc.positional_emb.reshape(-1, dim))
        desc.ln_eps = float(enc.norm.eps)
        desc.norm_g, desc.norm_b = f32(enc.norm.weight), f32(enc.norm.bias)
        desc.pool_w = f32(enc.attention_pool.weight.reshape(-1))
        desc.pool_b = float(enc.attention_pool.bias.detach().float().item())
        desc.layers = C.cast(self.layers, C.POINTER(L.LayerDesc))
        self.desc, self.dim = desc, dim
        self.tokens = enc.sequence_length

    def workspace_bytes(self, batch: int) -> int:
        n = L.load().devit_cct_workspace_bytes(C.byref(self.desc), batch)
        if n == 0:
            L.check(1)
        return n


class CCT(nn.Module):
    """models/cct.py:38-178."""

    def __init__(self, img_size=224, embedding_dim=768, n_input_channels=3, n_conv_layers=1,
                 kernel_size=7, stride=2, padding=3, pooling_kernel_size=3, pooling_stride=2,
                 pooling_padding=1, dropout=0., attention_dropout=0.1, stochastic_depth=0.1,
                 num_layers=14, num_heads=6, mlp_ratio=4.0, num_classes=1000,
                 positional_embedding='learnable', resize_dim=None, backbone=False, *args,
                 **kwargs):
        super().__init__()
        self.backbone, self.img_size = backbone, img_size
        self.tokenizer = Tokenizer(n_input_channels=n_input_channels,
                                   n_output_channels=embedding_dim, kernel_size=kernel_size,
                                   stride=stride, padding=padding,
                                   pooling_kernel_size=pooling_kernel_size,
                                   pooling_stride=pooling_stride, pooling_padding=pooling_padding,
                                   max_pool=True, activation=nn.ReLU, n_conv_layers=n_conv_layers,
                                   conv_bias=False)
        kw = dict(sequence_length=self.tokenizer.sequence_length(n_input_channels, img_size,
                                                                 img_size),
                  embedding_dim=embedding_dim, seq_pool=True, dropout=float(dropout),
                  attention_dropout=attention_dropout, stochastic_depth=stochastic_depth,
                  num_layers=num_layers, num_heads=num_heads, mlp_ratio=mlp_ratio,
                  num_classes=num_classes, positional_embedding=positional_embedding)
        if backbone:
            self.encoders = CCTTransformer(**kw)
        else:
            self.classifier = TransformerClassifier(**kw)
        self.resize_dim = resize_dim
        if resize_dim is not None:
            self.resize = nn.Linear(embedding_dim, resize_dim)
        self.precision = default_precision()
        self._pack = None

    def set_precision(self, precision: str):
        if precision not in _PREC:
            raise ValueError(f"precision must be one of {list(_PREC)}")
        self.precision = precision
        return self

    def packed(self, device) -> PackedCCT:
        ver = (self.precision, str(device), packing.module_version(self))
        if self._pack is None or self._pack[0] != ver:
            self._pack = (ver, PackedCCT(self, _PREC[self.precision], device))
        return self._pack[1]

    @torch.no_grad()
    def pooled_features(self, x, x_out=None, num_layers=-1, out=None, rows=None):
        """images [B, C, H, W] -> sequence-pooled feature [B, dim] (fp32).  `out` (optional): a
        contiguous [B, dim] fp32 tensor to write into; `rows` = (b0, b1): process only that chunk
        of the batch (MultiCCT runs chunks as independent kernel chains)."""
        if not x.is_cuda:
            raise L.DevitError("devit_b200 models run on CUDA (sm_100) tensors only; "
                               "there is no CPU fallback")
        if self.training:
            raise L.DevitError("devit_b200 CCT is forward/inference only: call .eval()")
        assert x.shape[2] == self.img_size and x.shape[3] == self.img_size, \
            f"Input image size ({x.shape[2]}*{x.shape[3]}) doesn't match model ({self.img_size})"
        x = x.float().contiguous()
        pk = self.packed(x.device)
        B = x.shape[0]
        b0, b1 = (0, B) if rows is None else rows
        if not (0 <= b0 < b1 <= B):
            raise L.DevitError(f"pooled_features: rows {rows} outside the batch of {B}")
        pooled = out if out is not None else torch.empty(B, pk.dim, device=x.device)
        if pooled.shape != (B, pk.dim) or not pooled.is_contiguous() or \
                pooled.dtype != torch.float32:
            raise L.DevitError("pooled_features: `out` must be a contiguous fp32 [B, dim] tensor")
        if x_out is not None and rows is not None:
            x_out = x_out[b0:b1]
        with torch.cuda.device(x.device):  # the C side launches on the current device
            ws = packing.workspace(x.device, pk.workspace_bytes(b1 - b0))
            L.check(L.load().devit_cct_forward(
                C.byref(pk.desc), x.data_ptr() + b0 * x.stride(0) * 4, b1 - b0, ws.data_ptr(),
                ws.numel(), pooled.data_ptr() + b0 * pk.dim * 4, L.ptr(x_out), num_layers,
                L.stream_ptr(x.device)))
        return pooled

    @torch.no_grad()
    def forward(self, x, output_attention=False, output_hidden_states=False, output_pool=False,
                distill=False):
        if output_attention or output_hidden_states or distill:
            raise L.DevitError("devit_b200 CCT: per-layer attention / hidden-state outputs are "
                               "training-side consumers and are not produced by the fused path")
        pooled = self.pooled_features(x)
        if self.backbone:
            return pooled
        prec = _PREC[self.precision]
        fc = self.classifier.fc
        logits = L.gemm(L.to_operand(pooled, prec), L.to_operand(fc.weight.float(), prec),
                        precision=prec, bias=fc.bias.float(), out_kind=L.OUT_F32)
        return (logits, pooled) if output_pool else logits

    def reset_classifier(self, num_classes):
        self.classifier.fc = nn.Linear(self.classifier.embedding_dim, num_classes) \
            if num_classes > 0 else nn.Identity()


def _cct(arch, pretrained, progress, num_layers, num_heads, mlp_ratio, embedding_dim,
         kernel_size=3, stride=None, padding=None, resize_dim=None, backbone=False, *args,
         **kwargs):
    """models/cct.py:181-206 (pretrained URLs are not reachable from here: pass weights via
    load_state_dict)."""
    if pretrained:
        raise L.DevitError("devit_b200 CCT: load pretrained weights with load_state_dict")
    stride = stride if stride is not None else max(1, (kernel_size // 2) - 1)
    padding = padding if padding is not None else max(1, (kernel_size // 2))
    return CCT(num_layers=num_layers, num_heads=num_heads, mlp_ratio=mlp_ratio,
               embedding_dim=embedding_dim, kernel_size=kernel_size, stride=stride,
               padding=padding, resize_dim=resize_dim, backbone=backbone, *args, **kwargs)


def cct_7(arch=None, pretrained=False, progress=True, *args, **kwargs):
    return _cct(arch, pretrained, progress, num_layers=7, num_heads=4, mlp_ratio=2,
                embedding_dim=256, *args, **kwargs)


def cct_6(arch=None, pretrained=False, progress=True, *args, **kwargs):
    return _cct(arch, pretrained, progress, num_layers=6, num_heads=4, mlp_ratio=2,
                embedding_dim=256, *args, **kwargs)


def _entry(name, fn, n_conv, default_classes):
    def entry(pretrained=False, progress=False, img_size=32, positional_embedding='learnable',
              num_classes=default_classes, *args, **kwargs):
        return fn(name, pretrained, progress, kernel_size=3, n_conv_layers=n_conv,
                  img_size=img_size, positional_embedding=positional_embedding,
                  num_classes=num_classes, *args, **kwargs)
    entry.__name__ = name
    return register_model(entry)


cct_6_3x1_32 = _entry('cct_6_3x1_32', cct_6, 1, 10)
cct_6_3x2_32 = _entry('cct_6_3x2_32', cct_6, 2, 10)
cct_7_3x1_32 = _entry('cct_7_3x1_32', cct_7, 1, 10)
cct_7_3x1_32_c100 = _entry('cct_7_3x1_32_c100', cct_7, 1, 100)
cct_7_3x2_32 = _entry('cct_7_3x2_32', cct_7, 2, 10)


def get_decct(pretrained_path=None, num_classes=1000, progress=False, kernel_size=3,
              n_conv_layers=2, img_size=32, positional_embedding='learnable', backbone=False,
              *args, **kwargs):
    """models/cct.py:461-470."""
    model = cct_7(pretrained=False, progress=progress, kernel_size=kernel_size,
                  n_conv_layers=n_conv_layers, img_size=img_size,
                  positional_embedding=positional_embedding, num_classes=num_classes,
                  backbone=backbone, *args, **kwargs)
    if pretrained_path is not None:
        model.load_state_dict(torch.load(pretrained_path))
    return model


class MultiCCT(nn.Module):
    """models/ensemble_models.py:93-113 (see the module docstring for the `[0]` deviation)."""

    def __init__(self, model_type, num_classes_list=[25, 25, 25, 25], num_sub_models=4,
                 input_size=224):
        super().__init__()
        self.model_type = model_type
        assert len(num_classes_list) == num_sub_models, \
            'num of classes is not match num of sub-models'
        if self.model_type.split('_')[0] != 'decct':
            raise L.DevitError(f"MultiCCT: unknown model type {model_type!r}")
        kernel_size, conv_layers = [int(i) for i in self.model_type.split('_')[-1].split('x')]
        self.models = nn.ModuleList(get_decct(img_size=input_size, kernel_size=kernel_size,
                                              n_conv_layers=conv_layers, num_classes=n,
                                              backbone=True) for n in num_classes_list)
        self.precision = default_precision()

    def set_precision(self, precision: str):
        self.precision = precision
        for m in self.models:
            m.set_precision(precision)
        return self

    @torch.no_grad()
    def forward_slab(self, x, subs=None):
        """Runs the backbones in `subs` (default: all) as independent kernel chains and returns
        (slab_f32, slab_op), both [len(subs), 1, B, dim] (slab_op in the GEMM operand format: bf16,
        or [2, ...] split planes in the fp32 mode) -- the layout ShardedEnsemble all-gathers."""
        from .ensemble import chain_tasks, run_chains
        subs = list(range(len(self.models))) if subs is None else list(subs)
        B, D = x.shape[0], self.models[subs[0]].tokenizer.conv_layers[-1][0].out_channels
        f32 = torch.empty(len(subs), 1, B, D, device=x.device)

        def launch(task):
            i, s, rows = task
            m = self.models[s]
            if m.precision != self.precision:
                m.set_precision(self.precision)
            m.pooled_features(x, out=f32[i, 0], rows=rows)

        run_chains(x.device, chain_tasks(subs, B, x.is_cuda), launch)
        return f32, L.to_operand(f32, _PREC[self.precision])

    @torch.no_grad()
    def forward(self, x):
        from .ensemble import FeatureList
        f32, op = self.forward_slab(x)
        out = FeatureList(f32[s, 0] for s in range(f32.shape[0]))
        out.slab_f32, out.slab_op, out.kind, out.slab_version = f32, op, 0, f32._version
        return out


class EnsembleCCT(nn.Module):
    """models/ensemble_models.py:116-151: fusion Linear(n * sub_size -> [teacher_size ->] C) as a
    K-segmented GEMM over the per-sub-model features (no stack / view copy)."""

    def __init__(self, sub_size=256, teacher_size=None, num_sub_models=4, num_classes=100):
        super().__init__()
        self.sub_size, self.teacher_size = sub_size, teacher_size
        self.num_sub_models, self.num_classes = num_sub_models, num_classes
        self.sum_feature_dim = sub_size * num_sub_models
        if teacher_size is None:
            self.cls_classifier = nn.Linear(self.sum_feature_dim, num_classes)
        else:
            self.cls_mlp = nn.Linear(self.sum_feature_dim, teacher_size)
            self.cls_classifier = nn.Linear(teacher_size, num_classes)
        self.precision = default_precision()

    def set_precision(self, precision: str):
        self.precision = precision
        return self

    def _first(self):
        return self.cls_classifier if self.teacher_size is None else self.cls_mlp

    def _head(self, slab, order, distill=False):
        """slab: operand-format [n, 1, B, D] (or [2, n, 1, B, D]); entry j = sub-model order[j]."""
        prec = _PREC[self.precision]
        s4 = slab if prec == L.DEVIT_BF16 else slab[0]
        n, kinds, B, D = s4.shape
        if n != self.num_sub_models or n > 8 or D != self.sub_size or kinds != 1:
            raise L.DevitError(f"EnsembleCCT: expected {self.num_sub_models} (<= 8) feature "
                               f"tensors of width {self.sub_size}")
        order = list(range(n)) if order is None else list(order)
        if sorted(order) != list(range(n)):
            raise L.DevitError(f"EnsembleCCT: slab order {order} is not a permutation")
        a = slab.view(n * B, D) if prec == L.DEVIT_BF16 else slab.view(2, n * B, D)
        # K-segments in SUB-MODEL order whatever the slab layout: same summation order, hence the
        # same logits bit for bit, for every world size
        segs = sorted(((j * B, 0, order[j] * D, D) for j in range(n)), key=lambda sg: sg[2])
        first = self._first()
        w = L.to_operand(first.weight.float(), prec)
        h = L.gemm(a, w, precision=prec, m=B, segs=segs, bias=first.bias.float(),
                   out_kind=L.OUT_F32, tag=6)
        if self.teacher_size is None:
            return h
        logits = L.gemm(L.to_operand(h, prec), L.to_operand(self.cls_classifier.weight.float(), prec),
                        precision=prec, bias=self.cls_classifier.bias.float(), out_kind=L.OUT_F32,
                        tag=6)
        if distill and self.training:
            return h, logits
        return logits

    @torch.no_grad()
    def forward_gathered(self, slab, order=None):
        """Fusion head straight from a gathered operand slab (ShardedEnsemble)."""
        return self._head(slab, order)

    @torch.no_grad()
    def forward(self, sub_model_features, distill=False):
        prec = _PREC[self.precision]
        feats = sub_model_features
        slab = getattr(feats, 'slab_op', None)
        n = len(feats)
        if slab is None or feats.slab_f32._version != feats.slab_version or \
                any(feats[j].data_ptr() != feats.slab_f32[j, 0].data_ptr() for j in range(n)):
            # plain list of tensors (or a list whose entries were replaced): stack, as the
            # reference does (models/ensemble_models.py:139)
            B = feats[0].shape[0]
            slab = L.to_operand(torch.stack([f.float() for f in feats], 0)
                                .reshape(n, 1, B, -1), prec)
        return self._head(slab, None, distill)
