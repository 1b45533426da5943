"""Model registry with timm's calling convention.

The reference registers its models with ``timm.models.registry.register_model`` and builds them
with ``timm.models.create_model`` (models/de_vit.py:495-513, models/ensemble_models.py:23-27).
When timm is installed the devit_b200 entrypoints are registered THERE too (a later registration
of the same name overrides timm's stock model, exactly like importing the reference's modules
does), so the reference's scripts pick them up unchanged.  Without timm this module provides the
two functions itself.
"""
from __future__ import annotations

import logging

_ENTRYPOINTS = {}
TIMM_FAILURES = {}  # entrypoint name -> why timm's registry refused it
_log = logging.getLogger("devit_b200.registry")


def _find_timm_register():
    try:
        from timm.models.registry import register_model as reg
        return reg
    except ImportError:  # timm is not installed (the build image): use the local registry only
        return None


_timm_register = _find_timm_register()


def register_model(fn):
    _ENTRYPOINTS[fn.__name__] = fn
    if _timm_register is not None:
        try:
            _timm_register(fn)
        except Exception as e:  # noqa: BLE001  (timm versions differ in what they require of fn)
            # not fatal -- devit_b200.registry.create_model still works -- but a reference script
            # calling timm.create_model would silently get timm's stock model: say so
            TIMM_FAILURES[fn.__name__] = f"{type(e).__name__}: {e}"
            _log.warning("timm refused to register %r (%s: %s); timm.create_model(%r) will NOT "
                         "return the devit_b200 model, use devit_b200.registry.create_model",
                         fn.__name__, type(e).__name__, e, fn.__name__)
    return fn


def is_model(name: str) -> bool:
    return name in _ENTRYPOINTS


def list_models():
    return sorted(_ENTRYPOINTS)


def create_model(model_name, pretrained=False, **kwargs):
    """timm 0.5.4 semantics: kwargs whose value is None are dropped before the entrypoint is
    called (the reference relies on it: ``drop_block_rate=None`` at ensemble_models.py:27)."""
    if model_name not in _ENTRYPOINTS:
        raise RuntimeError(f"Unknown model ({model_name}); registered: {list_models()}")
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    return _ENTRYPOINTS[model_name](pretrained=pretrained, **kwargs)
