"""Functional CPU restatement of the HAT hot path (TEST INFRASTRUCTURE -- never imported by product code).

Restates ``modules/hat_arch.py`` of the reference from closed-form index math: 16x16 window
attention with the relative-position table (no index buffer on the module: the index arrives as a forward argument,
hat_arch.py:166), the per-forward shift mask (:921-940), HAB with its CAB conv branch on the un-shifted LN1 output
(:267-310), OCAB with zero-padded overlapping key/value windows and the negative-index wrap of ``rpi_oca``
(:393-439, :897-918), RHAG (:619-620) and the HAT trunk (:950-992).

Parity pin: tests/golden/hat_*.npz and kat_hat_*.npz, outputs of the unmodified reference (oracle/make_golden_hat.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .swinir_oracle import (RGB_MEAN, conv3x3, gelu, image_to_tokens, layer_norm, mlp, pixel_shuffle,
                            relative_position_index, shift_attention_mask, tokens_to_image, window_token_pixels)
from tpu_superresolution_b200.synth import HATConfig  # noqa: F401  (constructor-argument records live with the synthetic-data generators)

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------
# index math
# ----------------------------------------------------------------------------------------
def rpi_sa(ws: int) -> Tensor:
    """calculate_rpi_sa, hat_arch.py:882-895 (same closed form as SwinIR's buffer)."""
    return relative_position_index(ws)


def rpi_oca(ws: int, overlap_ratio: float) -> Tensor:
    """calculate_rpi_oca, hat_arch.py:897-918: (ws*ws, wse*wse) with NEGATIVE entries.

    idx(i, j) = ((yj - yi) + ws - wse + 1) * (ws + wse - 1) + ((xj - xi) + ws - wse + 1); query i = (yi, xi) on the
    ws grid, key j = (yj, xj) on the wse grid.  Indexing the (ws+wse-1)^2 table with it wraps negatives (python
    indexing), i.e. the effective row is idx mod (ws+wse-1)^2.
    """
    wse = ws + int(overlap_ratio * ws)
    ti = torch.arange(ws * ws)
    tj = torch.arange(wse * wse)
    yi, xi = ti // ws, ti % ws
    yj, xj = tj // wse, tj % wse
    dy = yj[None, :] - yi[:, None] + ws - wse + 1
    dx = xj[None, :] - xi[:, None] + ws - wse + 1
    return dy * (ws + wse - 1) + dx


def ocab_key_pixels(H: int, W: int, ws: int, wse: int) -> Tuple[Tensor, Tensor]:
    """nn.Unfold(kernel=wse, stride=ws, padding=(wse-ws)//2), hat_arch.py:378: for window w = (wy, wx) and key
    (oy, ox) the source pixel (wy*ws - pad + oy, wx*ws - pad + ox); returns (flat index clamped to 0, validity)."""
    pad = (wse - ws) // 2
    nwy, nwx = H // ws, W // ws
    y = torch.arange(nwy).view(nwy, 1, 1, 1) * ws - pad + torch.arange(wse).view(1, 1, wse, 1)
    x = torch.arange(nwx).view(1, nwx, 1, 1) * ws - pad + torch.arange(wse).view(1, 1, 1, wse)
    valid = ((y >= 0) & (y < H) & (x >= 0) & (x < W)).reshape(nwy * nwx, wse * wse)
    flat = (y.clamp(0, H - 1) * W + x.clamp(0, W - 1)).reshape(nwy * nwx, wse * wse)
    return flat, valid


# ----------------------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------------------
def hat_window_attention(xw: Tensor, p: Dict[str, Tensor], pre: str, num_heads: int, ws: int,
                         mask: Optional[Tensor]) -> Tensor:
    """WindowAttention.forward(x, rpi, mask), hat_arch.py:166-197 with rpi = calculate_rpi_sa()."""
    B_, N, C = xw.shape
    d = C // num_heads
    qkv = (xw @ p[pre + "qkv.weight"].T + p[pre + "qkv.bias"]).reshape(B_, N, 3, num_heads, d)
    q = qkv[:, :, 0].permute(0, 2, 1, 3) * d ** -0.5
    k = qkv[:, :, 1].permute(0, 2, 1, 3)
    v = qkv[:, :, 2].permute(0, 2, 1, 3)
    attn = q @ k.transpose(-2, -1)
    bias = p[pre + "relative_position_bias_table"][rpi_sa(ws).reshape(-1)].reshape(N, N, -1).permute(2, 0, 1)
    attn = attn + bias[None]
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.reshape(B_ // nW, nW, num_heads, N, N) + mask[None, :, None]).reshape(B_, num_heads, N, N)
    attn = torch.softmax(attn, dim=-1)
    out = (attn @ v).permute(0, 2, 1, 3).reshape(B_, N, C)
    return out @ p[pre + "proj.weight"].T + p[pre + "proj.bias"]


def cab(x_img: Tensor, p: Dict[str, Tensor], pre: str) -> Tensor:
    """CAB.forward, hat_arch.py:62-75 + ChannelAttention :41-59.  x_img: (B, C, H, W)."""
    y = conv3x3(x_img, p[pre + "cab.0.weight"], p[pre + "cab.0.bias"])
    y = gelu(y)
    y = conv3x3(y, p[pre + "cab.2.weight"], p[pre + "cab.2.bias"])
    s = y.mean(dim=(2, 3), keepdim=True)                                         # AdaptiveAvgPool2d(1)
    s = F.conv2d(s, p[pre + "cab.3.attention.1.weight"], p[pre + "cab.3.attention.1.bias"]).clamp_min(0)
    s = torch.sigmoid(F.conv2d(s, p[pre + "cab.3.attention.3.weight"], p[pre + "cab.3.attention.3.bias"]))
    return y * s


def hab(x: Tensor, x_size: Tuple[int, int], p: Dict[str, Tensor], pre: str, num_heads: int, ws: int, shift: int,
        conv_scale: float, input_resolution: Optional[Tuple[int, int]] = None) -> Tensor:
    """HAB.forward, hat_arch.py:267-310.  The window/shift clamp of :247-250 looks at the CONSTRUCTOR's input_resolution
    (img_size), not at x_size: a 16x16 input to a model built for 64x64 still runs shifted 16x16 windows."""
    H, W = x_size
    B, L, C = x.shape
    if input_resolution is not None and min(input_resolution) <= ws:             # :247-250
        shift, ws = 0, min(input_resolution)
    xn = layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"])
    conv_x = image_to_tokens(cab(tokens_to_image(xn, x_size), p, pre + "conv_block."))   # on the UN-shifted LN1 output
    pix = window_token_pixels(H, W, ws, shift)
    nW, N = pix.shape
    xw = xn[:, pix.reshape(-1)].reshape(B * nW, N, C)
    mask = shift_attention_mask(H, W, ws, shift, x.dtype) if shift > 0 else None  # HAB ignores the mask when shift == 0 (:281-286)
    aw = hat_window_attention(xw, p, pre + "attn.", num_heads, ws, mask)
    merged = torch.empty_like(x)
    merged[:, pix.reshape(-1)] = aw.reshape(B, nW * N, C)
    x = x + merged + conv_x * conv_scale                                         # :307
    return x + mlp(layer_norm(x, p[pre + "norm2.weight"], p[pre + "norm2.bias"]), p, pre + "mlp.")


def ocab(x: Tensor, x_size: Tuple[int, int], p: Dict[str, Tensor], pre: str, num_heads: int, ws: int,
         overlap_ratio: float) -> Tensor:
    """OCAB.forward, hat_arch.py:393-439."""
    H, W = x_size
    B, L, C = x.shape
    d = C // num_heads
    wse = int(ws * overlap_ratio) + ws
    xn = layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"])
    qkv = xn @ p[pre + "qkv.weight"].T + p[pre + "qkv.bias"]                     # :401, on all tokens
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    pix = window_token_pixels(H, W, ws, 0)
    nW, N = pix.shape
    kpix, kvalid = ocab_key_pixels(H, W, ws, wse)                                # zero padding of the PROJECTED k, v
    NK = wse * wse
    qw = q[:, pix.reshape(-1)].reshape(B, nW, N, num_heads, d).permute(0, 1, 3, 2, 4) * d ** -0.5
    kw = (k[:, kpix.reshape(-1)] * kvalid.reshape(1, -1, 1)).reshape(B, nW, NK, num_heads, d).permute(0, 1, 3, 2, 4)
    vw = (v[:, kpix.reshape(-1)] * kvalid.reshape(1, -1, 1)).reshape(B, nW, NK, num_heads, d).permute(0, 1, 3, 2, 4)
    attn = qw @ kw.transpose(-2, -1)                                             # (B, nW, nH, N, NK)
    table = p[pre + "relative_position_bias_table"]
    idx = rpi_oca(ws, overlap_ratio) % table.shape[0]                            # negative indices wrap (A.3)
    bias = table[idx.reshape(-1)].reshape(N, NK, -1).permute(2, 0, 1)
    attn = torch.softmax(attn + bias[None, None], dim=-1)
    out = (attn @ vw).permute(0, 1, 3, 2, 4).reshape(B, nW * N, C)
    merged = torch.empty_like(x)
    merged[:, pix.reshape(-1)] = out
    x = merged @ p[pre + "proj.weight"].T + p[pre + "proj.bias"] + x             # :436
    return x + mlp(layer_norm(x, p[pre + "norm2.weight"], p[pre + "norm2.bias"]), p, pre + "mlp.")   # :438


def rhag(x: Tensor, x_size: Tuple[int, int], p: Dict[str, Tensor], pre: str, depth: int, num_heads: int,
         cfg: HATConfig) -> Tensor:
    """RHAG.forward, hat_arch.py:619-620: depth HABs (shift 0 / ws//2 alternating, :499), OCAB, 3x3 conv, + x."""
    ws = cfg.window_size
    y = x
    for b in range(depth):
        y = hab(y, x_size, p, f"{pre}residual_group.blocks.{b}.", num_heads, ws, 0 if b % 2 == 0 else ws // 2,
                cfg.conv_scale)
    y = ocab(y, x_size, p, pre + "residual_group.overlap_attn.", num_heads, ws, cfg.overlap_ratio)
    y = conv3x3(tokens_to_image(y, x_size), p[pre + "conv.weight"], p[pre + "conv.bias"])
    return image_to_tokens(y) + x


def hat_forward(lr: Tensor, p: Dict[str, Tensor], cfg: HATConfig) -> Tensor:
    """HAT.forward for upsampler='pixelshuffle', hat_arch.py:978-994."""
    H, W = lr.shape[2:]
    ws = cfg.window_size
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    x = F.pad(lr, (0, pw, 0, ph), "reflect") if (ph or pw) else lr
    mean = torch.tensor(RGB_MEAN if cfg.in_chans == 3 else (0.0,), dtype=lr.dtype).view(1, -1, 1, 1)
    x = (x - mean) * cfg.img_range
    f0 = conv3x3(x, p["conv_first.weight"], p["conv_first.bias"])
    x_size = (f0.shape[2], f0.shape[3])
    t = layer_norm(image_to_tokens(f0), p["patch_embed.norm.weight"], p["patch_embed.norm.bias"])
    for g, (depth, nh) in enumerate(zip(cfg.depths, cfg.num_heads)):
        t = rhag(t, x_size, p, f"layers.{g}.", depth, nh, cfg)
    t = layer_norm(t, p["norm.weight"], p["norm.bias"])
    body = conv3x3(tokens_to_image(t, x_size), p["conv_after_body.weight"], p["conv_after_body.bias"]) + f0
    y = F.leaky_relu(conv3x3(body, p["conv_before_upsample.0.weight"], p["conv_before_upsample.0.bias"]), 0.01)
    for i in range(int(math.log2(cfg.upscale))):
        y = pixel_shuffle(conv3x3(y, p[f"upsample.{2 * i}.weight"], p[f"upsample.{2 * i}.bias"]), 2)
    y = conv3x3(y, p["conv_last.weight"], p["conv_last.bias"])
    y = y / cfg.img_range + mean
    return y[:, :, :H * cfg.upscale, :W * cfg.upscale]
