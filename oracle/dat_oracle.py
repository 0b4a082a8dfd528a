"""Functional CPU restatement of the DAT hot path (TEST INFRASTRUCTURE -- never imported by product code).

Restates ``modules/dat_arch.py`` of the reference: rectangular split-window attention on the two channel halves with the
dynamic position bias (a 4-layer MLP on relative offsets, input independent; dat_arch.py:93-130, :133-244), the per-axis
shifted-window masks (:318-361), the adaptive interaction module (:418-431), channel attention with token-wise L2
normalisation (:481-528), the spatial-gate feed-forward (:38-90), DATB / ResidualGroup / DAT (:531-652, :828-858).
Window gathers are written as closed-form index math instead of the reference's view/permute/roll chain.

Parity pin: tests/golden/dat_*.npz and kat_dat_*.npz, outputs of the unmodified reference (oracle/make_golden_dat.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Sequence, Tuple

import torch
import torch.nn.functional as F

from .swinir_oracle import RGB_MEAN, conv3x3, gelu, image_to_tokens, layer_norm, pixel_shuffle, tokens_to_image
from tpu_superresolution_b200.synth import DATConfig  # noqa: F401  (constructor-argument records live with the synthetic-data generators)

Tensor = torch.Tensor


def is_shifted(rg_idx: int, b_idx: int) -> bool:
    """dat_arch.py:290 / :389: which spatial blocks run on the shifted window grid."""
    return (rg_idx % 2 == 0 and b_idx > 0 and (b_idx - 2) % 4 == 0) or (rg_idx % 2 != 0 and b_idx % 4 == 0)


def rect_window_pixels(H: int, W: int, hs: int, ws: int, sy: int, sx: int) -> Tensor:
    """(nW, hs*ws) flat pixel index read by token (ty, tx) of window (wy, wx) after torch.roll(-sy, -sx) and img2windows
    (dat_arch.py:391-394, 15-24): ((wy*hs + ty + sy) mod H, (wx*ws + tx + sx) mod W)."""
    nwy, nwx = H // hs, W // ws
    y = (torch.arange(nwy).view(nwy, 1, 1, 1) * hs + torch.arange(hs).view(1, 1, hs, 1) + sy) % H
    x = (torch.arange(nwx).view(1, nwx, 1, 1) * ws + torch.arange(ws).view(1, 1, 1, ws) + sx) % W
    return (y * W + x).reshape(nwy * nwx, hs * ws)


def rect_shift_mask(H: int, W: int, hs: int, ws: int, sy: int, sx: int, dtype=torch.float32) -> Tensor:
    """(nW, N, N) 0 / -100 mask of one branch (dat_arch.py:318-361): region(p, L, win, shift) = [p >= L - win] + [p >= L - shift]
    per axis on the shifted grid; tokens in different regions do not attend."""
    def region(p, L, win, shift):
        return (p >= L - win).long() + (p >= L - shift).long()
    nwy, nwx = H // hs, W // ws
    hp = (torch.arange(nwy).view(nwy, 1, 1, 1) * hs + torch.arange(hs).view(1, 1, hs, 1)).expand(nwy, nwx, hs, ws)
    wp = (torch.arange(nwx).view(1, nwx, 1, 1) * ws + torch.arange(ws).view(1, 1, 1, ws)).expand(nwy, nwx, hs, ws)
    ids = (3 * region(hp, H, hs, sy) + region(wp, W, ws, sx)).reshape(nwy * nwx, hs * ws)
    diff = ids[:, None, :] != ids[:, :, None]
    return torch.where(diff, torch.tensor(-100.0, dtype=dtype), torch.tensor(0.0, dtype=dtype))


def dynamic_pos_bias_table(p: Dict[str, Tensor], pre: str, hs: int, ws: int) -> Tensor:
    """DynamicPosBias (residual=False) on the (2hs-1)(2ws-1) offsets, dat_arch.py:93-130, :171-177 -> (offsets, heads)."""
    dy = torch.arange(1 - hs, hs)
    dx = torch.arange(1 - ws, ws)
    b = torch.stack(torch.meshgrid(dy, dx, indexing="ij")).flatten(1).transpose(0, 1).to(p[pre + "pos_proj.weight"].dtype)
    t = b @ p[pre + "pos_proj.weight"].T + p[pre + "pos_proj.bias"]
    for name in ("pos1", "pos2", "pos3"):
        t = layer_norm(t, p[pre + name + ".0.weight"], p[pre + name + ".0.bias"]).clamp_min(0)
        t = t @ p[pre + name + ".2.weight"].T + p[pre + name + ".2.bias"]
    return t


def rect_relative_position_index(hs: int, ws: int) -> Tensor:
    """dat_arch.py:180-190: (yi - yj + hs - 1) * (2 ws - 1) + (xi - xj + ws - 1)."""
    t = torch.arange(hs * ws)
    y, x = t // ws, t % ws
    return (y[:, None] - y[None, :] + hs - 1) * (2 * ws - 1) + (x[:, None] - x[None, :] + ws - 1)


def spatial_branch(q: Tensor, k: Tensor, v: Tensor, H: int, W: int, hs: int, ws: int, sy: int, sx: int, heads: int,
                   p: Dict[str, Tensor], pre: str, shifted: bool) -> Tensor:
    """Spatial_Attention.forward on one channel half, dat_arch.py:203-244.  q, k, v: (B, H*W, Ch) -> (B, H*W, Ch)."""
    B, L, Ch = q.shape
    d = Ch // heads
    pix = rect_window_pixels(H, W, hs, ws, sy if shifted else 0, sx if shifted else 0)
    nW, N = pix.shape

    def win(t):
        return t[:, pix.reshape(-1)].reshape(B, nW, N, heads, d).permute(0, 1, 3, 2, 4)

    attn = (win(q) * d ** -0.5) @ win(k).transpose(-2, -1)
    table = dynamic_pos_bias_table(p, pre + "pos.", hs, ws)
    bias = table[rect_relative_position_index(hs, ws).reshape(-1)].reshape(N, N, heads).permute(2, 0, 1)
    attn = attn + bias[None, None]
    if shifted:
        attn = attn + rect_shift_mask(H, W, hs, ws, sy, sx, q.dtype)[None, :, None]
    out = (torch.softmax(attn, dim=-1) @ win(v)).permute(0, 1, 3, 2, 4).reshape(B, nW * N, Ch)
    merged = torch.empty_like(q)
    merged[:, pix.reshape(-1)] = out
    return merged


def _bn(x: Tensor, p: Dict[str, Tensor], pre: str, eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm2d in eval mode on (B, C, H, W)."""
    sh = (1, -1, 1, 1)
    return (x - p[pre + "running_mean"].view(sh)) / torch.sqrt(p[pre + "running_var"].view(sh) + eps) * p[pre + "weight"].view(sh) \
        + p[pre + "bias"].view(sh)


def _dwconv_bn_gelu(v_img: Tensor, p: Dict[str, Tensor], pre: str) -> Tensor:
    """self.dwconv: depthwise 3x3 + BatchNorm + GELU (dat_arch.py:300-304)."""
    y = F.conv2d(v_img, p[pre + "dwconv.0.weight"], p[pre + "dwconv.0.bias"], padding=1, groups=v_img.shape[1])
    return gelu(_bn(y, p, pre + "dwconv.1."))


def _channel_interaction(x_img: Tensor, p: Dict[str, Tensor], pre: str) -> Tensor:
    s = x_img.mean(dim=(2, 3), keepdim=True)
    s = gelu(_bn(F.conv2d(s, p[pre + "channel_interaction.1.weight"], p[pre + "channel_interaction.1.bias"]), p,
                 pre + "channel_interaction.2."))
    return F.conv2d(s, p[pre + "channel_interaction.4.weight"], p[pre + "channel_interaction.4.bias"])     # (B, C, 1, 1)


def _spatial_interaction(x_img: Tensor, p: Dict[str, Tensor], pre: str) -> Tensor:
    s = gelu(_bn(F.conv2d(x_img, p[pre + "spatial_interaction.0.weight"], p[pre + "spatial_interaction.0.bias"]), p,
                 pre + "spatial_interaction.1."))
    return F.conv2d(s, p[pre + "spatial_interaction.3.weight"], p[pre + "spatial_interaction.3.bias"])     # (B, 1, H, W)


def adaptive_spatial_attention(x: Tensor, H: int, W: int, p: Dict[str, Tensor], pre: str, heads: int,
                               split: Sequence[int], shifted: bool) -> Tensor:
    """Adaptive_Spatial_Attention.forward, dat_arch.py:363-438, including the zero padding of the PROJECTED q, k, v to a multiple
    of max(split) (:376-385: the qkv bias is not re-added on the pad), masks for the padded size (:396-399) and the crop (:406-407)."""
    B, L, C = x.shape
    qkv = x @ p[pre + "qkv.weight"].T + p[pre + "qkv.bias"]
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    h2, c2 = heads // 2, C // 2
    s0, s1 = split[0] // 2, split[1] // 2
    m = max(split)
    Hp, Wp = H + (m - H % m) % m, W + (m - W % m) % m

    def padded(t):          # (B, H*W, c) -> (B, Hp*Wp, c), zeros at the bottom / right
        if (Hp, Wp) == (H, W):
            return t
        return F.pad(t.reshape(B, H, W, -1), (0, 0, 0, Wp - W, 0, Hp - H)).reshape(B, Hp * Wp, -1)

    def cropped(t):         # (B, Hp, Wp, c) -> (B, H*W, c)
        return t[:, :H, :W, :].reshape(B, H * W, -1)

    qp, kp, vp = padded(q), padded(k), padded(v)
    x1 = cropped(spatial_branch(qp[..., :c2], kp[..., :c2], vp[..., :c2], Hp, Wp, split[0], split[1], s0, s1, h2, p, pre + "attns.0.", shifted).reshape(B, Hp, Wp, c2))
    x2 = cropped(spatial_branch(qp[..., c2:], kp[..., c2:], vp[..., c2:], Hp, Wp, split[1], split[0], s1, s0, h2, p, pre + "attns.1.", shifted).reshape(B, Hp, Wp, c2))
    att = torch.cat([x1, x2], dim=2)
    conv_x = _dwconv_bn_gelu(tokens_to_image(v, (H, W)), p, pre)
    channel_map = _channel_interaction(conv_x, p, pre).permute(0, 2, 3, 1).reshape(B, 1, C)
    spatial_map = _spatial_interaction(tokens_to_image(att, (H, W)), p, pre)
    att = att * torch.sigmoid(channel_map)
    conv_x = image_to_tokens(torch.sigmoid(spatial_map) * conv_x)
    return (att + conv_x) @ p[pre + "proj.weight"].T + p[pre + "proj.bias"]


def adaptive_channel_attention(x: Tensor, H: int, W: int, p: Dict[str, Tensor], pre: str, heads: int) -> Tensor:
    """Adaptive_Channel_Attention.forward, dat_arch.py:481-528."""
    B, N, C = x.shape
    d = C // heads
    qkv = (x @ p[pre + "qkv.weight"].T + p[pre + "qkv.bias"]).reshape(B, N, 3, heads, d).permute(2, 0, 3, 4, 1)   # (3, B, h, d, N)
    q, k, v = qkv[0], qkv[1], qkv[2]
    v_img = v.reshape(B, C, H, W)
    qn = q / q.norm(dim=-1, keepdim=True).clamp_min(1e-12)                    # F.normalize over the tokens (:497-498)
    kn = k / k.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    attn = torch.softmax((qn @ kn.transpose(-2, -1)) * p[pre + "temperature"], dim=-1)       # (B, h, d, d)
    att = (attn @ v).permute(0, 3, 1, 2).reshape(B, N, C)
    conv_x = _dwconv_bn_gelu(v_img, p, pre)
    channel_map = _channel_interaction(tokens_to_image(att, (H, W)), p, pre)                  # (B, C, 1, 1)
    spatial_map = _spatial_interaction(conv_x, p, pre).permute(0, 2, 3, 1).reshape(B, N, 1)
    att = att * torch.sigmoid(spatial_map)
    conv_x = image_to_tokens(conv_x * torch.sigmoid(channel_map))
    return (att + conv_x) @ p[pre + "proj.weight"].T + p[pre + "proj.bias"]


def sgfn(x: Tensor, H: int, W: int, p: Dict[str, Tensor], pre: str) -> Tensor:
    """SGFN.forward + SpatialGate, dat_arch.py:38-90."""
    h = gelu(x @ p[pre + "fc1.weight"].T + p[pre + "fc1.bias"])
    x1, x2 = h.chunk(2, dim=-1)
    x2 = layer_norm(x2, p[pre + "sg.norm.weight"], p[pre + "sg.norm.bias"])
    x2 = image_to_tokens(F.conv2d(tokens_to_image(x2, (H, W)), p[pre + "sg.conv.weight"], p[pre + "sg.conv.bias"], padding=1,
                                  groups=x2.shape[-1]))
    return (x1 * x2) @ p[pre + "fc2.weight"].T + p[pre + "fc2.bias"]


def datb(x: Tensor, x_size: Tuple[int, int], p: Dict[str, Tensor], pre: str, heads: int, split: Sequence[int], rg_idx: int,
         b_idx: int) -> Tensor:
    """DATB.forward, dat_arch.py:556-565."""
    H, W = x_size
    xn = layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"])
    if b_idx % 2 == 0:
        x = x + adaptive_spatial_attention(xn, H, W, p, pre + "attn.", heads, split, is_shifted(rg_idx, b_idx))
    else:
        x = x + adaptive_channel_attention(xn, H, W, p, pre + "attn.", heads)
    return x + sgfn(layer_norm(x, p[pre + "norm2.weight"], p[pre + "norm2.bias"]), H, W, p, pre + "ffn.")


def residual_group(x: Tensor, x_size: Tuple[int, int], p: Dict[str, Tensor], pre: str, depth: int, heads: int,
                   split: Sequence[int], rg_idx: int) -> Tensor:
    """ResidualGroup.forward with '1conv', dat_arch.py:635-652."""
    y = x
    for b in range(depth):
        y = datb(y, x_size, p, f"{pre}blocks.{b}.", heads, split, rg_idx, b)
    y = conv3x3(tokens_to_image(y, x_size), p[pre + "conv.weight"], p[pre + "conv.bias"])
    return x + image_to_tokens(y)


def dat_forward(lr: Tensor, p: Dict[str, Tensor], cfg: DATConfig) -> Tensor:
    """DAT.forward for upsampler='pixelshuffle', dat_arch.py:828-858 (no padding / cropping at model level)."""
    mean = torch.tensor(RGB_MEAN if cfg.in_chans == 3 else (0.0,), dtype=lr.dtype).view(1, -1, 1, 1)
    x = (lr - mean) * cfg.img_range
    f0 = conv3x3(x, p["conv_first.weight"], p["conv_first.bias"])
    x_size = (f0.shape[2], f0.shape[3])
    t = layer_norm(image_to_tokens(f0), p["before_RG.1.weight"], p["before_RG.1.bias"])
    for g, (depth, nh) in enumerate(zip(cfg.depth, cfg.num_heads)):
        t = residual_group(t, x_size, p, f"layers.{g}.", depth, nh, cfg.split_size, g)
    t = layer_norm(t, p["norm.weight"], p["norm.bias"])
    body = conv3x3(tokens_to_image(t, x_size), p["conv_after_body.weight"], p["conv_after_body.bias"]) + f0
    y = F.leaky_relu(conv3x3(body, p["conv_before_upsample.0.weight"], p["conv_before_upsample.0.bias"]), 0.01)
    if cfg.upscale & (cfg.upscale - 1) == 0:
        for i in range(int(math.log2(cfg.upscale))):
            y = pixel_shuffle(conv3x3(y, p[f"upsample.{2 * i}.weight"], p[f"upsample.{2 * i}.bias"]), 2)
    else:
        y = pixel_shuffle(conv3x3(y, p["upsample.0.weight"], p["upsample.0.bias"]), 3)
    y = conv3x3(y, p["conv_last.weight"], p["conv_last.bias"])
    return y / cfg.img_range + mean
