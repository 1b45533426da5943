"""Import shim that lets the UNMODIFIED reference run outside its own (torch 1.8 / timm 0.5.4)
environment: from /root/reference in the build container, or from the byte copy of its hot-path
modules that oracle/make_ref.py places under the git-ignored oracle/_ref/ (which travels to the
GPU box).  Used by tests/golden/make_*.py, the reference-vs-oracle cross-checks and the
`--impl reference` / `cpu_baseline` legs of bench.py.  Three non-invasive shims
 -- none edits the reference:
  1. a stand-in for the parts of timm==0.5.4 the hot path imports (timm is not installed);
  2. ``builtins.partial`` / ``builtins.nn`` so models/utils/config.py:4 imports;
  3. CPU only: ``Tensor.get_device`` returns the tensor's device so the gate lands on the
     activation's device (models/de_vit.py:42,78 fail on CPU under torch >= 2).
"""
from __future__ import annotations

import builtins
import functools
import math
import sys
import types

import torch
import torch.nn as nn

def reference_root():
    """/root/reference when present (build container), else oracle/_ref (GPU box), else None."""
    import os
    from pathlib import Path
    forced = os.environ.get("DEVIT_REF_ROOT")  # tests: force the oracle/_ref copy
    if forced:
        return forced if os.path.isdir(os.path.join(forced, "models")) else None
    if os.path.isdir("/root/reference/models"):
        return "/root/reference"
    ref = Path(__file__).resolve().parent / "_ref"
    if (ref / "models" / "de_vit.py").exists():
        return str(ref)
    return None


REFERENCE_ROOT = reference_root()


class _PatchEmbed(nn.Module):
    """2D image -> patch tokens: Conv2d(k=s=patch) then flatten(2).transpose(1, 2)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None,
                 flatten=True):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        _, _, h, w = x.shape
        assert h == self.img_size[0] and w == self.img_size[1], "input size mismatch"
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class _DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if not self.training or not self.drop_prob:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x.div(keep) * mask


class _Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


def _trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def _lecun_normal_(tensor):
    fan_in = nn.init._calculate_fan_in_and_fan_out(tensor)[0]
    return nn.init.trunc_normal_(tensor, std=math.sqrt(1.0 / fan_in) / .87962566103423978)


def _named_apply(fn, module, name='', depth_first=True, include_root=False):
    if not depth_first and include_root:
        fn(module=module, name=name)
    for child_name, child in module.named_children():
        child_name = '.'.join((name, child_name)) if name else child_name
        _named_apply(fn=fn, module=child, name=child_name, depth_first=depth_first,
                     include_root=True)
    if depth_first and include_root:
        fn(module=module, name=name)
    return module


def _adapt_input_conv(in_chans, conv_weight):
    return conv_weight


_REGISTRY = {}


def _register_model(fn):
    _REGISTRY[fn.__name__] = fn
    return fn


def _create_model(model_name, pretrained=False, **kwargs):
    kwargs = {k: v for k, v in kwargs.items() if v is not None}  # timm 0.5.4 drops None kwargs
    return _REGISTRY[model_name](pretrained=pretrained, **kwargs)


def _cfg(url='', **kwargs):
    return {'url': url, 'num_classes': 1000, 'input_size': (3, 224, 224), **kwargs}


def install():
    """Install the shims and put /root/reference on sys.path. Idempotent."""
    if 'timm' not in sys.modules:
        timm = types.ModuleType('timm')
        models = types.ModuleType('timm.models')
        layers = types.ModuleType('timm.models.layers')
        helpers = types.ModuleType('timm.models.helpers')
        registry = types.ModuleType('timm.models.registry')
        vit = types.ModuleType('timm.models.vision_transformer')
        layers.PatchEmbed, layers.DropPath, layers.Mlp = _PatchEmbed, _DropPath, _Mlp
        layers.trunc_normal_, layers.lecun_normal_ = _trunc_normal_, _lecun_normal_
        helpers.named_apply, helpers.adapt_input_conv = _named_apply, _adapt_input_conv
        registry.register_model = _register_model
        vit._cfg = _cfg
        models.create_model = _create_model
        models.layers, models.helpers, models.registry = layers, helpers, registry
        models.vision_transformer = vit
        timm.models = models
        for m in (timm, models, layers, helpers, registry, vit):
            sys.modules[m.__name__] = m
    builtins.partial = functools.partial
    builtins.nn = nn
    if not getattr(torch.Tensor.get_device, '_devit_shim', False):
        def _get_device(t):
            return t.device
        _get_device._devit_shim = True
        torch.Tensor.get_device = _get_device
    if REFERENCE_ROOT is None:
        raise ImportError("the reference is neither at /root/reference nor under oracle/_ref "
                          "(run oracle/make_ref.py where /root/reference exists)")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def _accuracy(output, target, topk=(1,)):
    """timm 0.5.4 timm/utils/metrics.py accuracy(), restated: the top-max(topk) predictions of
    every sample are compared with the target; precision@k in percent of the batch."""
    maxk = min(max(topk), output.size()[1])
    batch_size = target.size(0)
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.reshape(1, -1).expand_as(pred))
    return [correct[:min(k, maxk)].reshape(-1).float().sum(0) * 100. / batch_size for k in topk]


def _install_engine_shims():
    """timm names engine.py / utils/losses.py import at module level (engine.py:11-13,
    utils/losses.py:7).  Only `accuracy` is used by the evaluation loops; the others are
    training-only and get inert placeholders."""
    if 'timm.utils' in sys.modules:
        return
    timm = sys.modules['timm']
    data = types.ModuleType('timm.data')
    utils = types.ModuleType('timm.utils')
    clip = types.ModuleType('timm.utils.clip_grad')
    loss = types.ModuleType('timm.loss')
    data.Mixup = type('Mixup', (), {})
    utils.accuracy = _accuracy
    utils.ModelEma = type('ModelEma', (), {})
    clip.dispatch_clip_grad = lambda *a, **k: None
    loss.SoftTargetCrossEntropy = type('SoftTargetCrossEntropy', (nn.Module,), {})
    utils.clip_grad = clip
    timm.data, timm.utils, timm.loss = data, utils, loss
    for m in (data, utils, clip, loss):
        sys.modules[m.__name__] = m


def load_reference_engine():
    """The reference's engine module (evaluate / evaluate_ens_disjoint with its own MetricLogger
    and CrossEntropyLoss; timm's accuracy() restated above)."""
    install()
    _install_engine_shims()
    import engine  # noqa: E402  (reference)
    return engine


def load_reference():
    """Returns (de_vit module, deit_vit module, ensemble_models module, create_model)."""
    install()
    import models.de_vit as de_vit  # noqa: E402  (reference)
    import models.ensemble_models as ens  # noqa: E402
    import models.deit_vit as deit_vit  # noqa: E402  (registers the teacher last)
    return de_vit, deit_vit, ens, _create_model
