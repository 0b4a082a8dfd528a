"""Deterministic synthetic inputs, weights and the constructor-argument records of the BASELINE.json configs.

Data generation only (used by bench.py, __graft_entry__.smoke(), the tools and the tests); no arithmetic of the path lives here
and nothing here imports oracle/.  The index / mask buffers of the state_dicts come from the drop-in modules' own closed
forms (swinir.calculate_mask, hat.calculate_rpi_*, dat._rect_*), which the CPU tests pin to the reference's buffers.


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

from dataclasses import dataclass, field
from typing import Sequence


@dataclass
class SwinIRConfig:
    """Constructor arguments of the reference SwinIR (network_swinir.py:646-652)."""
    upscale: int = 2
    in_chans: int = 3
    img_size: int = 64
    window_size: int = 8
    img_range: float = 1.0
    depths: Sequence[int] = field(default_factory=lambda: [6] * 6)
    embed_dim: int = 180
    num_heads: Sequence[int] = field(default_factory=lambda: [6] * 6)
    mlp_ratio: float = 2.0
    upsampler: str = "pixelshuffle"
    resi_connection: str = "1conv"
    num_feat: int = 64

    def as_kwargs(self) -> dict:
        return dict(upscale=self.upscale, in_chans=self.in_chans, img_size=self.img_size,
                    window_size=self.window_size, img_range=self.img_range,
                    depths=list(self.depths), embed_dim=self.embed_dim,
                    num_heads=list(self.num_heads), mlp_ratio=self.mlp_ratio,
                    upsampler=self.upsampler, resi_connection=self.resi_connection)


@dataclass
class HATConfig:
    """Constructor arguments of the reference HAT (hat_arch.py:738-764); SURVEY.md 8(d) cfg3 values."""
    upscale: int = 4
    in_chans: int = 3
    img_size: int = 64
    window_size: int = 16
    compress_ratio: int = 3
    squeeze_factor: int = 30
    conv_scale: float = 0.01
    overlap_ratio: float = 0.5
    img_range: float = 1.0
    depths: Sequence[int] = field(default_factory=lambda: [6] * 6)
    embed_dim: int = 180
    num_heads: Sequence[int] = field(default_factory=lambda: [6] * 6)
    mlp_ratio: float = 2.0
    upsampler: str = "pixelshuffle"
    resi_connection: str = "1conv"
    num_feat: int = 64

    def as_kwargs(self) -> dict:
        return dict(upscale=self.upscale, in_chans=self.in_chans, img_size=self.img_size, window_size=self.window_size,
                    compress_ratio=self.compress_ratio, squeeze_factor=self.squeeze_factor, conv_scale=self.conv_scale,
                    overlap_ratio=self.overlap_ratio, img_range=self.img_range, depths=list(self.depths),
                    embed_dim=self.embed_dim, num_heads=list(self.num_heads), mlp_ratio=self.mlp_ratio,
                    upsampler=self.upsampler, resi_connection=self.resi_connection)


@dataclass
class DATConfig:
    """Constructor arguments of the reference DAT (dat_arch.py:717-737); SURVEY.md 8(d) cfg4 values."""
    upscale: int = 2
    in_chans: int = 3
    img_size: int = 64
    img_range: float = 1.0
    depth: Sequence[int] = field(default_factory=lambda: [6] * 6)
    embed_dim: int = 180
    num_heads: Sequence[int] = field(default_factory=lambda: [6] * 6)
    expansion_factor: float = 4.0
    resi_connection: str = "1conv"
    split_size: Sequence[int] = field(default_factory=lambda: [8, 32])
    num_feat: int = 64

    def as_kwargs(self) -> dict:
        return dict(upscale=self.upscale, in_chans=self.in_chans, img_size=self.img_size, img_range=self.img_range,
                    depth=list(self.depth), embed_dim=self.embed_dim, num_heads=list(self.num_heads),
                    expansion_factor=self.expansion_factor, resi_connection=self.resi_connection,
                    split_size=list(self.split_size))


def relative_position_index(ws: int) -> torch.Tensor:
    """(N, N) int64 table index, closed form of network_swinir.py:93-102 (same as WindowAttention's buffer)."""
    t = torch.arange(ws * ws)
    y, x = t // ws, t % ws
    return (y[:, None] - y[None, :] + ws - 1) * (2 * ws - 1) + (x[:, None] - x[None, :] + ws - 1)


def shift_attention_mask(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    from .swinir import calculate_mask
    return calculate_mask((H, W), ws, shift)



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
    from .hat import calculate_rpi_oca as rpi_oca, calculate_rpi_sa as rpi_sa
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


def dat_manifest(cfg) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) in the reference DAT's state_dict order (dat_arch.py:262-316, 452-479, 57-90, 531-555, 738-812)."""
    from .dat import is_shifted
    C, nf = cfg.embed_dim, cfg.num_feat
    hid = int(C * cfg.expansion_factor)
    s0, s1 = cfg.split_size
    N = s0 * s1
    nwin = (cfg.img_size // s0) * (cfg.img_size // s1)
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def conv(name, co, ci, k=3, wk="conv_w"):
        out.append((f"{name}.weight", (co, ci, k, k), wk))
        out.append((f"{name}.bias", (co,), "conv_b"))

    def ln(name, n=None):
        out.append((f"{name}.weight", (n or C,), "ln_w"))
        out.append((f"{name}.bias", (n or C,), "ln_b"))

    def lin(name, co, ci, wk="lin_w"):
        out.append((f"{name}.weight", (co, ci), wk))
        out.append((f"{name}.bias", (co,), "lin_b"))

    def bn(name, n):
        ln(name, n)
        out.append((f"{name}.running_mean", (n,), "bn_mean"))
        out.append((f"{name}.running_var", (n,), "bn_var"))
        out.append((f"{name}.num_batches_tracked", (), "bn_count"))

    def aim(pre):
        conv(pre + "dwconv.0", C, 1, 3, "dw_w")
        bn(pre + "dwconv.1", C)
        conv(pre + "channel_interaction.1", C // 8, C, 1)
        bn(pre + "channel_interaction.2", C // 8)
        conv(pre + "channel_interaction.4", C, C // 8, 1)
        conv(pre + "spatial_interaction.0", C // 16, C, 1)
        bn(pre + "spatial_interaction.1", C // 16)
        conv(pre + "spatial_interaction.3", 1, C // 16, 1)

    conv("conv_first", C, cfg.in_chans)
    ln("before_RG.1")
    for g, (depth, nh) in enumerate(zip(cfg.depth, cfg.num_heads)):
        for b in range(depth):
            pre = f"layers.{g}.blocks.{b}."
            ln(pre + "norm1")
            a = pre + "attn."
            if b % 2 == 0:
                if is_shifted(g, b):
                    out.append((a + "attn_mask_0", (nwin, N, N), "dat_mask0"))
                    out.append((a + "attn_mask_1", (nwin, N, N), "dat_mask1"))
                lin(a + "qkv", 3 * C, C, "qkv_w")
                lin(a + "proj", C, C)
                pd = (C // 2) // 4 // 4                       # DynamicPosBias(dim // 4) -> pos_dim = dim // 16 (dat_arch.py:170, :105)
                for i in range(2):
                    s = a + f"attns.{i}."
                    out.append((s + "rpe_biases", ((2 * s0 - 1) * (2 * s1 - 1), 2), f"rpe{i}"))
                    out.append((s + "relative_position_index", (N, N), f"rpi{i}"))
                    lin(s + "pos.pos_proj", pd, 2, "pos_w")
                    for j, o in ((1, pd), (2, pd), (3, nh // 2)):
                        ln(s + f"pos.pos{j}.0", pd)
                        lin(s + f"pos.pos{j}.2", o, pd, "pos_w")
            else:
                out.append((a + "temperature", (nh, 1, 1), "temperature"))
                lin(a + "qkv", 3 * C, C, "qkv_w")
                lin(a + "proj", C, C)
            aim(a)
            f = pre + "ffn."
            lin(f + "fc1", hid, C)
            ln(f + "sg.norm", hid // 2)
            conv(f + "sg.conv", hid // 2, 1, 3, "dw_w")
            lin(f + "fc2", C, hid // 2)
            ln(pre + "norm2")
        conv(f"layers.{g}.conv", C, C)
    ln("norm")
    conv("conv_after_body", C, C)
    conv("conv_before_upsample.0", nf, C)
    for i in range(int(math.log2(cfg.upscale))):
        conv(f"upsample.{2 * i}", 4 * nf, nf)
    conv("conv_last", cfg.in_chans, nf)
    return out


def make_dat_state_dict(cfg, seed: int = 1234, kind: str = "init") -> Dict[str, torch.Tensor]:
    """Synthetic DAT state_dict with the reference's keys/shapes.  "stress" also gives the BatchNorms non-trivial running
    statistics and the dynamic-position-bias MLP weights of order one, so the bias table is far from constant."""
    from .dat import _rect_rpi as rect_relative_position_index, _rect_mask as rect_shift_mask
    assert kind in ("init", "stress")
    rng = np.random.default_rng(seed)
    s0, s1 = cfg.split_size
    R = cfg.img_size
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, k in dat_manifest(cfg):
        if k in ("rpe0", "rpe1"):
            hs, ws = (s0, s1) if k == "rpe0" else (s1, s0)
            dy, dx = torch.arange(1 - hs, hs), torch.arange(1 - ws, ws)
            sd[key] = torch.stack(torch.meshgrid(dy, dx, indexing="ij")).flatten(1).transpose(0, 1).contiguous().float()
        elif k in ("rpi0", "rpi1"):
            sd[key] = rect_relative_position_index(*((s0, s1) if k == "rpi0" else (s1, s0)))
        elif k == "dat_mask0":
            sd[key] = rect_shift_mask(R, R, s0, s1, s0 // 2, s1 // 2)
        elif k == "dat_mask1":
            sd[key] = rect_shift_mask(R, R, s1, s0, s1 // 2, s0 // 2)
        elif k == "bn_count":
            sd[key] = torch.tensor(0, dtype=torch.int64)
        else:
            if k == "bn_mean":
                a = np.zeros(shape) if kind == "init" else rng.normal(0, 0.2, size=shape)
            elif k == "bn_var":
                a = np.ones(shape) if kind == "init" else rng.uniform(0.5, 1.5, size=shape)
            elif k == "temperature":
                a = np.ones(shape) if kind == "init" else rng.uniform(0.5, 3.0, size=shape)
            elif k == "dw_w":
                a = rng.uniform(-1.0 / 3.0, 1.0 / 3.0, size=shape)
            elif k == "pos_w":
                a = np.clip(rng.normal(0, 0.02, size=shape), -0.04, 0.04) if kind == "init" else rng.normal(0, 0.7, size=shape)
            else:
                a = _draw(rng, k, shape, kind, cfg.embed_dim)
            sd[key] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return sd


def _dat_configs():
    return {
        # BASELINE.json configs[3]: DAT x2, split 8x32, expansion 4
        "dat_x2": DATConfig(upscale=2),
        # reduced depth, full widths: RG 0 holds an un-shifted and a shifted spatial block (b = 0, 2), RG 1 a shifted one at b = 0
        "dat_x2_d3": DATConfig(upscale=2, depth=[3, 2], num_heads=[6, 6]),
    }


def _hat_configs():
    return {
        # BASELINE.json configs[2]: HAT x4, window 16, overlap 0.5
        "hat_x4": HATConfig(upscale=4),
        # reduced depth, full widths (same kernels): 2 RHAGs of 2 HABs + OCAB
        "hat_x4_d2": HATConfig(upscale=4, depths=[2, 2], num_heads=[6, 6]),
        "hat_x2_d2": HATConfig(upscale=2, depths=[2, 2], num_heads=[6, 6]),
    }


HAT_CONFIGS = _hat_configs()
DAT_CONFIGS = _dat_configs()


# ---------------------------------------------------------------------------------------------------------------
# generic synthetic weights for the constructor variants (upsamplers, '3conv', ape): keyed by parameter NAME, so the
# reference model and the drop-in model get the same tensors whatever their registration order
# ---------------------------------------------------------------------------------------------------------------
VARIANT_KW = dict(img_size=16, window_size=8, img_range=1.0, depths=[2], embed_dim=180, num_heads=[6], mlp_ratio=2.0, in_chans=3)
SWINIR_VARIANTS = {
    "pixelshuffledirect_x2": dict(VARIANT_KW, upscale=2, upsampler="pixelshuffledirect", resi_connection="1conv"),
    "nearestconv_x4": dict(VARIANT_KW, upscale=4, upsampler="nearest+conv", resi_connection="1conv"),
    "denoise_x1": dict(VARIANT_KW, upscale=1, upsampler="", resi_connection="1conv"),
    "pixelshuffle_3conv_x2": dict(VARIANT_KW, upscale=2, upsampler="pixelshuffle", resi_connection="3conv"),
    "pixelshuffle_ape_x2": dict(VARIANT_KW, upscale=2, upsampler="pixelshuffle", resi_connection="1conv", ape=True),
    "pixelshuffle_x3": dict(VARIANT_KW, upscale=3, upsampler="pixelshuffle", resi_connection="1conv"),
}


def generic_state_dict(template: Dict[str, torch.Tensor], seed: int = 7) -> Dict[str, torch.Tensor]:
    """Fill every floating-point PARAMETER of `template` (a model's state_dict) from a numpy RNG seeded by (seed, key name):
    weights ~ N(0, 0.06) (attention logits of order one), biases ~ N(0, 0.05), norm weights 1 + N(0, 0.1), bias tables ~ N(0, 0.5).
    Integer buffers and masks (relative_position_index, attn_mask) keep the template's values."""
    import zlib
    out = {}
    for k, v in template.items():
        if not v.is_floating_point() or k.endswith("attn_mask"):
            out[k] = v.clone()
            continue
        rng = np.random.default_rng([seed, zlib.crc32(k.encode())])
        if "relative_position_bias_table" in k:
            a = rng.normal(0, 0.5, size=tuple(v.shape))
        elif "absolute_pos_embed" in k:
            a = rng.normal(0, 0.1, size=tuple(v.shape))
        elif "norm" in k and k.endswith("weight"):
            a = 1.0 + rng.normal(0, 0.1, size=tuple(v.shape))
        elif k.endswith("bias"):
            a = rng.normal(0, 0.05, size=tuple(v.shape))
        else:
            fan_in = int(np.prod(v.shape[1:])) if v.dim() > 1 else int(v.shape[0])
            a = rng.normal(0, min(0.06, 1.5 / np.sqrt(fan_in)), size=tuple(v.shape))
        out[k] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return out
