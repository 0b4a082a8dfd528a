"""Deterministic synthetic inputs and weights for parity tests and the bench (TEST INFRASTRUCTURE).

Weights come from ``numpy.random.default_rng`` (stable across numpy versions and
platforms), NOT from the reference's torch-RNG init, so the very same tensors can be
rebuilt on the GPU box where ``/root/reference`` does not exist.  The key names and
shapes follow the reference's ``state_dict()``; the committed manifest
``tests/golden/swinir_x2_manifest.json`` (made from the real reference by
``oracle/make_golden.py``) pins them.

Two weight sets:
  * "init":   reference-like statistics (Linear ~ N(0, .02) clipped at 2 sigma, bias 0, LN (1, 0),
              convs kaiming-uniform-like); network_swinir.py:766-773.
  * "stress": SURVEY.md §4.3 -- logits with std 2-4, RPB table std 1, non-zero biases,
              LN affine != (1, 0), so a wrong RPB index / mask region / transposed K cannot hide
              behind a near-uniform softmax.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

from .swinir_oracle import SwinIRConfig, relative_position_index, shift_attention_mask


def swinir_manifest(cfg: SwinIRConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) in the reference's state_dict order (network_swinir.py:646-764)."""
    C, ws, nf = cfg.embed_dim, cfg.window_size, cfg.num_feat
    hid = int(C * cfg.mlp_ratio)
    N = ws * ws
    nwin = (cfg.img_size // ws) ** 2
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def conv(name, co, ci):
        out.append((f"{name}.weight", (co, ci, 3, 3), "conv_w"))
        out.append((f"{name}.bias", (co,), "conv_b"))

    def ln(name):
        out.append((f"{name}.weight", (C,), "ln_w"))
        out.append((f"{name}.bias", (C,), "ln_b"))

    conv("conv_first", C, cfg.in_chans)
    ln("patch_embed.norm")
    for g, (depth, nh) in enumerate(zip(cfg.depths, cfg.num_heads)):
        for b in range(depth):
            pre = f"layers.{g}.residual_group.blocks.{b}."
            if b % 2 == 1:
                out.append((pre + "attn_mask", (nwin, N, N), "attn_mask"))   # buffer built for img_size
            ln(pre + "norm1")
            out.append((pre + "attn.relative_position_bias_table", ((2 * ws - 1) ** 2, nh), "rpb"))
            out.append((pre + "attn.relative_position_index", (N, N), "rpi"))
            out.append((pre + "attn.qkv.weight", (3 * C, C), "qkv_w"))
            out.append((pre + "attn.qkv.bias", (3 * C,), "lin_b"))
            out.append((pre + "attn.proj.weight", (C, C), "lin_w"))
            out.append((pre + "attn.proj.bias", (C,), "lin_b"))
            ln(pre + "norm2")
            out.append((pre + "mlp.fc1.weight", (hid, C), "lin_w"))
            out.append((pre + "mlp.fc1.bias", (hid,), "lin_b"))
            out.append((pre + "mlp.fc2.weight", (C, hid), "lin_w"))
            out.append((pre + "mlp.fc2.bias", (C,), "lin_b"))
        conv(f"layers.{g}.conv", C, C)
    ln("norm")
    conv("conv_after_body", C, C)
    conv("conv_before_upsample.0", nf, C)
    if cfg.upscale & (cfg.upscale - 1) == 0:
        for i in range(int(math.log2(cfg.upscale))):
            conv(f"upsample.{2 * i}", 4 * nf, nf)
    elif cfg.upscale == 3:
        conv("upsample.0", 9 * nf, nf)
    conv("conv_last", cfg.in_chans, nf)
    return out


def _draw(rng, k: str, shape, kind: str, C: int) -> np.ndarray:
    """One tensor of the synthetic state_dict; `k` is the manifest kind, `kind` the weight set."""
    if k == "conv_w":
        fan_in = shape[1] * shape[2] * shape[3]
        bound = 1.0 / math.sqrt(fan_in)                       # kaiming_uniform(a=sqrt(5)) bound
        return rng.uniform(-bound, bound, size=shape)
    if k == "conv_b":
        return rng.uniform(-0.05, 0.05, size=shape)
    if k == "ln_w":
        return np.ones(shape) if kind == "init" else rng.uniform(0.6, 1.4, size=shape)
    if k == "ln_b":
        return np.zeros(shape) if kind == "init" else rng.normal(0, 0.1, size=shape)
    if k == "rpb":
        return np.clip(rng.normal(0, 0.02, size=shape), -0.04, 0.04) if kind == "init" else rng.normal(0, 1.0, size=shape)
    if k == "qkv_w":
        if kind == "init":
            return np.clip(rng.normal(0, 0.02, size=shape), -0.04, 0.04)
        # logits std ~= sigma_q * sigma_k * C: 0.13^2 * 180 ~= 3
        return np.concatenate([rng.normal(0, 0.13, size=(2 * C, C)), rng.normal(0, 0.05, size=(C, C))], 0)
    if k == "lin_w":
        return np.clip(rng.normal(0, 0.02, size=shape), -0.04, 0.04) if kind == "init" else rng.normal(0, 0.03, size=shape)
    if k == "lin_b":
        return np.zeros(shape) if kind == "init" else rng.normal(0, 0.05, size=shape)
    raise KeyError(k)


def make_swinir_state_dict(cfg: SwinIRConfig, seed: int = 1234, kind: str = "init") -> Dict[str, torch.Tensor]:
    """Synthetic state_dict with the reference's keys/shapes (fp32; index buffers int64)."""
    assert kind in ("init", "stress")
    rng = np.random.default_rng(seed)
    C = cfg.embed_dim
    ws = cfg.window_size
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, k in swinir_manifest(cfg):
        if k == "rpi":
            sd[key] = relative_position_index(ws)
            continue
        if k == "attn_mask":
            sd[key] = shift_attention_mask(cfg.img_size, cfg.img_size, ws, ws // 2)
            continue
        sd[key] = torch.from_numpy(np.ascontiguousarray(_draw(rng, k, shape, kind, C), dtype=np.float32))
    return sd


def hat_manifest(cfg) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) in the reference HAT's state_dict order (hat_arch.py:766-880)."""
    C, ws, nf = cfg.embed_dim, cfg.window_size, cfg.num_feat
    hid = int(C * cfg.mlp_ratio)
    wse = ws + int(cfg.overlap_ratio * ws)
    out: List[Tuple[str, Tuple[int, ...], str]] = [("relative_position_index_SA", (ws * ws, ws * ws), "rpi_sa"),
                                                   ("relative_position_index_OCA", (ws * ws, wse * wse), "rpi_oca")]

    def conv(name, co, ci, k=3):
        out.append((f"{name}.weight", (co, ci, k, k), "conv_w"))
        out.append((f"{name}.bias", (co,), "conv_b"))

    def ln(name):
        out.append((f"{name}.weight", (C,), "ln_w"))
        out.append((f"{name}.bias", (C,), "ln_b"))

    def lin(name, co, ci, wk="lin_w"):
        out.append((f"{name}.weight", (co, ci), wk))
        out.append((f"{name}.bias", (co,), "lin_b"))

    def mlp_(pre):
        lin(pre + "mlp.fc1", hid, C)
        lin(pre + "mlp.fc2", C, hid)

    conv("conv_first", C, cfg.in_chans)
    ln("patch_embed.norm")
    for g, (depth, nh) in enumerate(zip(cfg.depths, cfg.num_heads)):
        for b in range(depth):
            pre = f"layers.{g}.residual_group.blocks.{b}."
            ln(pre + "norm1")
            out.append((pre + "attn.relative_position_bias_table", ((2 * ws - 1) ** 2, nh), "rpb"))
            lin(pre + "attn.qkv", 3 * C, C, "qkv_w")
            lin(pre + "attn.proj", C, C)
            conv(pre + "conv_block.cab.0", C // cfg.compress_ratio, C)
            conv(pre + "conv_block.cab.2", C, C // cfg.compress_ratio)
            conv(pre + "conv_block.cab.3.attention.1", C // cfg.squeeze_factor, C, 1)
            conv(pre + "conv_block.cab.3.attention.3", C, C // cfg.squeeze_factor, 1)
            ln(pre + "norm2")
            mlp_(pre)
        pre = f"layers.{g}.residual_group.overlap_attn."
        out.append((pre + "relative_position_bias_table", ((ws + wse - 1) ** 2, nh), "rpb"))
        ln(pre + "norm1")
        lin(pre + "qkv", 3 * C, C, "qkv_w")
        lin(pre + "proj", C, C)
        ln(pre + "norm2")
        mlp_(pre)
        conv(f"layers.{g}.conv", C, C)
    ln("norm")
    conv("conv_after_body", C, C)
    conv("conv_before_upsample.0", nf, C)
    for i in range(int(math.log2(cfg.upscale))):
        conv(f"upsample.{2 * i}", 4 * nf, nf)
    conv("conv_last", cfg.in_chans, nf)
    return out


def make_hat_state_dict(cfg, seed: int = 1234, kind: str = "init") -> Dict[str, torch.Tensor]:
    """Synthetic HAT state_dict with the reference's keys/shapes (hat_arch.py); see make_swinir_state_dict."""
    from .hat_oracle import rpi_oca, rpi_sa
    assert kind in ("init", "stress")
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, k in hat_manifest(cfg):
        if k == "rpi_sa":
            sd[key] = rpi_sa(cfg.window_size)
        elif k == "rpi_oca":
            sd[key] = rpi_oca(cfg.window_size, cfg.overlap_ratio)
        else:
            sd[key] = torch.from_numpy(np.ascontiguousarray(_draw(rng, k, shape, kind, cfg.embed_dim), dtype=np.float32))
    return sd


def make_lr_batch(batch: int, h: int = 64, w: int = 64, seed: int = 0, chans: int = 3) -> torch.Tensor:
    """LR tiles in [0,1) from a numpy RNG (float32, (B, chans, h, w))."""
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.random((batch, chans, h, w), dtype=np.float32))


def make_tokens(batch: int, h: int, w: int, c: int, seed: int = 0, scale: float = 1.0) -> torch.Tensor:
    """Feature-map tokens (B, h*w, c) with a per-channel offset so LayerNorm has work to do."""
    rng = np.random.default_rng(seed)
    a = rng.normal(0, scale, size=(batch, h * w, c)) + rng.normal(0, 0.5 * scale, size=(1, 1, c))
    return torch.from_numpy(a.astype(np.float32))


CONFIGS = {
    # BASELINE.json configs[0]: SwinIR classical x2, fp32, 1x3x64x64 (finetune_swinir.py:269-281)
    "swinir_x2": SwinIRConfig(upscale=2),
    # BASELINE.json configs[1] and [4]: SwinIR classical x4
    "swinir_x4": SwinIRConfig(upscale=4),
    # reduced-depth variants for fast KATs (same widths, so the same kernels run)
    "swinir_x2_d2": SwinIRConfig(upscale=2, depths=[2, 2], num_heads=[6, 6]),
    "swinir_x4_d2": SwinIRConfig(upscale=4, depths=[2, 2], num_heads=[6, 6]),
}


def _hat_configs():
    from .hat_oracle import HATConfig
    return {
        # BASELINE.json configs[2]: HAT x4, window 16, overlap 0.5
        "hat_x4": HATConfig(upscale=4),
        # reduced depth, full widths (same kernels): 2 RHAGs of 2 HABs + OCAB
        "hat_x4_d2": HATConfig(upscale=4, depths=[2, 2], num_heads=[6, 6]),
        "hat_x2_d2": HATConfig(upscale=2, depths=[2, 2], num_heads=[6, 6]),
    }


HAT_CONFIGS = _hat_configs()
