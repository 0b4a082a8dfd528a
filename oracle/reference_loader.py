"""Import the UNMODIFIED reference modules in the build container (TEST INFRASTRUCTURE).

``/root/reference`` exists only in the build container, never on the GPU box, so this
module is used by ``oracle/make_golden.py`` (fixture generation) and by CPU tests that
skip when the reference is absent.  The reference imports three symbols from ``timm``
(network_swinir.py:11, hat_arch.py:6, dat_arch.py:7) which is not installed; a stub with
the same semantics is injected first.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
import warnings

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("SRK_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "network_swinir.py"))


class _DropPath(nn.Module):
    """timm.layers.DropPath: identity when drop_prob == 0 or not training."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def install_timm_shim() -> None:
    if "timm" in sys.modules and not getattr(sys.modules["timm"], "_srk_shim", False):
        return  # a real timm is importable; use it
    for name in ("timm", "timm.layers", "timm.models", "timm.models.layers"):
        m = types.ModuleType(name)
        m._srk_shim = True
        m.DropPath = _DropPath
        m.to_2tuple = _to_2tuple
        m.trunc_normal_ = torch.nn.init.trunc_normal_
        sys.modules[name] = m
    sys.modules["timm"].layers = sys.modules["timm.layers"]
    sys.modules["timm"].models = sys.modules["timm.models"]
    sys.modules["timm.models"].layers = sys.modules["timm.models.layers"]


def load_reference_module(name: str):
    """name in {'network_swinir', 'hat_arch', 'dat_arch'} -> the imported reference module."""
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    install_timm_shim()
    mod_dir = os.path.join(REFERENCE_ROOT, "modules")
    if mod_dir not in sys.path:
        sys.path.insert(0, mod_dir)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module(name)
