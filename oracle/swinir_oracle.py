"""Functional CPU restatement of the SwinIR window-attention path (TEST INFRASTRUCTURE).

A plain-PyTorch (CPU, fp32 or fp64) restatement of what the reference computes in
``modules/network_swinir.py``; every function cites the reference lines it follows.
It is written from the closed-form index math rather than the
reference's view/permute/roll chain, so agreement with the golden fixtures made from
the real reference (``oracle/make_golden.py``) is a genuine cross-check.

Weights are taken as a flat ``dict[str, Tensor]`` with the reference's state_dict key
names, so the same dict can be loaded ``strict=True`` into the reference and into the
drop-in modules.

Parity pin: tests/golden/swinir_*.npz (outputs of the unmodified reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from tpu_superresolution_b200.synth import SwinIRConfig  # noqa: F401  (constructor-argument records live with the synthetic-data generators)

Tensor = torch.Tensor


RGB_MEAN = (0.4488, 0.4371, 0.4040)  # network_swinir.py:659


# ----------------------------------------------------------------------------------------
# index math (closed forms)
# ----------------------------------------------------------------------------------------
def relative_position_index(ws_h: int, ws_w: Optional[int] = None) -> Tensor:
    """(N, N) int64 table index; network_swinir.py:93-102.

    idx(i, j) = (yi - yj + ws_h - 1) * (2*ws_w - 1) + (xi - xj + ws_w - 1), token t = (t // ws_w, t % ws_w).
    """
    ws_w = ws_h if ws_w is None else ws_w
    t = torch.arange(ws_h * ws_w)
    y, x = t // ws_w, t % ws_w
    dy = y[:, None] - y[None, :] + ws_h - 1
    dx = x[:, None] - x[None, :] + ws_w - 1
    return dy * (2 * ws_w - 1) + dx


def relative_position_bias(table: Tensor, ws: int) -> Tensor:
    """(nH, N, N) bias gathered from the (2ws-1)^2 x nH table; network_swinir.py:127-129."""
    idx = relative_position_index(ws)
    n = ws * ws
    return table[idx.reshape(-1)].reshape(n, n, -1).permute(2, 0, 1).contiguous()


def window_token_pixels(H: int, W: int, ws: int, shift: int) -> Tensor:
    """(nW, N) int64: flat pixel index y*W+x read by token t of window w after the cyclic
    shift by (-shift, -shift) and the partition (network_swinir.py:249-256, 33-45).

    Window w = (wy, wx) row-major, token t = (ty, tx): source pixel
    ((wy*ws + ty + shift) mod H, (wx*ws + tx + shift) mod W).  window_reverse + the
    inverse roll write back to the same pixel (network_swinir.py:264-272, 48-62).
    """
    nwy, nwx = H // ws, W // ws
    wy = torch.arange(nwy).view(nwy, 1, 1, 1)
    wx = torch.arange(nwx).view(1, nwx, 1, 1)
    ty = torch.arange(ws).view(1, 1, ws, 1)
    tx = torch.arange(ws).view(1, 1, 1, ws)
    y = (wy * ws + ty + shift) % H
    x = (wx * ws + tx + shift) % W
    return (y * W + x).reshape(nwy * nwx, ws * ws)


def shift_attention_mask(H: int, W: int, ws: int, shift: int, dtype=torch.float32) -> Tensor:
    """(nW, N, N) mask of 0 / -100.0; network_swinir.py:216-237.

    In shifted coordinates region(p, L) = [p >= L - ws] + [p >= L - shift]; tokens with
    different 3*region(h) + region(w) ids must not attend to each other.
    """
    def region(p: Tensor, L: int) -> Tensor:
        return (p >= L - ws).long() + (p >= L - shift).long()

    nwy, nwx = H // ws, W // ws
    hp = (torch.arange(nwy).view(nwy, 1, 1, 1) * ws + torch.arange(ws).view(1, 1, ws, 1)).expand(nwy, nwx, ws, ws)
    wp = (torch.arange(nwx).view(1, nwx, 1, 1) * ws + torch.arange(ws).view(1, 1, 1, ws)).expand(nwy, nwx, ws, ws)
    ids = (3 * region(hp, H) + region(wp, W)).reshape(nwy * nwx, ws * ws)
    diff = ids[:, None, :] != ids[:, :, None]
    return torch.where(diff, torch.tensor(-100.0, dtype=dtype), torch.tensor(0.0, dtype=dtype))


def window_partition(x: Tensor, ws: int) -> Tensor:
    """(B,H,W,C) -> (B*nW, ws, ws, C); network_swinir.py:33-45 (restated as a gather)."""
    B, H, W, C = x.shape
    pix = window_token_pixels(H, W, ws, 0)
    return x.reshape(B, H * W, C)[:, pix.reshape(-1)].reshape(-1, ws, ws, C)


def window_reverse(windows: Tensor, ws: int, H: int, W: int) -> Tensor:
    """(B*nW, ws, ws, C) -> (B,H,W,C); network_swinir.py:48-62 (restated as a scatter)."""
    C = windows.shape[-1]
    nW = (H // ws) * (W // ws)
    B = windows.shape[0] // nW
    pix = window_token_pixels(H, W, ws, 0).reshape(-1)
    out = torch.empty(B, H * W, C, dtype=windows.dtype)
    out[:, pix] = windows.reshape(B, nW * ws * ws, C)
    return out.reshape(B, H, W, C)


def pixel_shuffle(x: Tensor, r: int) -> Tensor:
    """out[b, c, h*r+i, w*r+j] = in[b, c*r*r + i*r + j, h, w]  (nn.PixelShuffle; network_swinir.py:585)."""
    B, C, H, W = x.shape
    c = C // (r * r)
    return x.reshape(B, c, r, r, H, W).permute(0, 1, 4, 2, 5, 3).reshape(B, c, H * r, W * r)


# ----------------------------------------------------------------------------------------
# arithmetic blocks
# ----------------------------------------------------------------------------------------
def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm over the last dim, biased variance, eps 1e-5 (network_swinir.py:199,205)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu(x: Tensor) -> Tensor:
    """Exact erf GELU, the nn.GELU() default (network_swinir.py:15,20)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def window_attention(xw: Tensor, p: Dict[str, Tensor], pre: str, num_heads: int, ws: int,
                     mask: Optional[Tensor]) -> Tensor:
    """WindowAttention.forward, network_swinir.py:114-145.  xw: (B_, N, C); mask (nW, N, N) or None."""
    B_, N, C = xw.shape
    d = C // num_heads
    scale = d ** -0.5                                                       # :86
    qkv = xw @ p[pre + "qkv.weight"].T + p[pre + "qkv.bias"]                # :121
    qkv = qkv.reshape(B_, N, 3, num_heads, d)
    q = qkv[:, :, 0].permute(0, 2, 1, 3) * scale                            # (B_, nH, N, d)  :124
    k = qkv[:, :, 1].permute(0, 2, 1, 3)
    v = qkv[:, :, 2].permute(0, 2, 1, 3)
    attn = q @ k.transpose(-2, -1)                                          # :125
    attn = attn + relative_position_bias(p[pre + "relative_position_bias_table"], ws)[None]  # :127-130
    if mask is not None:                                                    # :132-135
        nW = mask.shape[0]
        attn = (attn.reshape(B_ // nW, nW, num_heads, N, N) + mask[None, :, None]).reshape(B_, num_heads, N, N)
    attn = torch.softmax(attn, dim=-1)                                      # :136
    out = (attn @ v).permute(0, 2, 1, 3).reshape(B_, N, C)                  # :142
    return out @ p[pre + "proj.weight"].T + p[pre + "proj.bias"]            # :143


def mlp(x: Tensor, p: Dict[str, Tensor], pre: str) -> Tensor:
    """Mlp.forward, network_swinir.py:24-30."""
    h = gelu(x @ p[pre + "fc1.weight"].T + p[pre + "fc1.bias"])
    return h @ p[pre + "fc2.weight"].T + p[pre + "fc2.bias"]


def swin_block(x: Tensor, x_size: Tuple[int, int], p: Dict[str, Tensor], pre: str, num_heads: int,
               ws: int, shift: int) -> Tensor:
    """SwinTransformerBlock.forward, network_swinir.py:239-279.  x: (B, H*W, C)."""
    H, W = x_size
    B, L, C = x.shape
    if min(H, W) <= ws:                                                     # :193-196 (ctor-time clamp)
        shift, ws = 0, min(H, W)
    xn = layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"])      # :245
    pix = window_token_pixels(H, W, ws, shift)                              # roll + partition  :249-256
    nW, N = pix.shape
    xw = xn[:, pix.reshape(-1)].reshape(B * nW, N, C)
    mask = shift_attention_mask(H, W, ws, shift, x.dtype) if shift > 0 else None   # :259-262
    aw = window_attention(xw, p, pre + "attn.", num_heads, ws, mask)
    merged = torch.empty_like(x)
    merged[:, pix.reshape(-1)] = aw.reshape(B, nW * N, C)                   # reverse + un-shift :264-272
    x = x + merged                                                          # :276
    return x + mlp(layer_norm(x, p[pre + "norm2.weight"], p[pre + "norm2.bias"]), p, pre + "mlp.")  # :277


def conv3x3(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """nn.Conv2d(k=3, s=1, p=1) on (B,C,H,W); network_swinir.py:465, 669, 729, 742-745."""
    return F.conv2d(x, w, b, stride=1, padding=1)


def tokens_to_image(x: Tensor, x_size: Tuple[int, int]) -> Tensor:
    """PatchUnEmbed.forward, network_swinir.py:562-565."""
    B, L, C = x.shape
    return x.transpose(1, 2).reshape(B, C, x_size[0], x_size[1])


def image_to_tokens(x: Tensor) -> Tensor:
    """PatchEmbed.forward without norm, network_swinir.py:524-525."""
    return x.flatten(2).transpose(1, 2)


def rstb(x: Tensor, x_size: Tuple[int, int], p: Dict[str, Tensor], pre: str, depth: int, num_heads: int,
         ws: int) -> Tensor:
    """RSTB.forward with '1conv', network_swinir.py:481-482; block shifts per :383."""
    y = x
    for b in range(depth):
        y = swin_block(y, x_size, p, f"{pre}residual_group.blocks.{b}.", num_heads, ws,
                       0 if b % 2 == 0 else ws // 2)
    y = conv3x3(tokens_to_image(y, x_size), p[pre + "conv.weight"], p[pre + "conv.bias"])
    return image_to_tokens(y) + x


def forward_features(x: Tensor, p: Dict[str, Tensor], cfg: SwinIRConfig) -> Tensor:
    """SwinIR.forward_features, network_swinir.py:790-803 (ape=False, patch_norm=True)."""
    x_size = (x.shape[2], x.shape[3])
    t = layer_norm(image_to_tokens(x), p["patch_embed.norm.weight"], p["patch_embed.norm.bias"])  # :792, :526-527
    for g, (depth, nh) in enumerate(zip(cfg.depths, cfg.num_heads)):
        t = rstb(t, x_size, p, f"layers.{g}.", depth, nh, cfg.window_size)      # :797-798
    t = layer_norm(t, p["norm.weight"], p["norm.bias"])                          # :800
    return tokens_to_image(t, x_size)                                            # :801


def upsample_tail(x: Tensor, p: Dict[str, Tensor], cfg: SwinIRConfig) -> Tensor:
    """conv_before_upsample + Upsample + conv_last, network_swinir.py:742-745, 580-585, 816-817."""
    x = F.leaky_relu(conv3x3(x, p["conv_before_upsample.0.weight"], p["conv_before_upsample.0.bias"]), 0.01)
    if cfg.upscale & (cfg.upscale - 1) == 0:
        for i in range(int(math.log2(cfg.upscale))):
            x = pixel_shuffle(conv3x3(x, p[f"upsample.{2 * i}.weight"], p[f"upsample.{2 * i}.bias"]), 2)
    elif cfg.upscale == 3:
        x = pixel_shuffle(conv3x3(x, p["upsample.0.weight"], p["upsample.0.bias"]), 3)
    else:
        raise ValueError(f"scale {cfg.upscale} is not supported")
    return conv3x3(x, p["conv_last.weight"], p["conv_last.bias"])


def swinir_forward(lr: Tensor, p: Dict[str, Tensor], cfg: SwinIRConfig) -> Tensor:
    """SwinIR.forward for upsampler='pixelshuffle', network_swinir.py:805-840."""
    assert cfg.upsampler == "pixelshuffle" and cfg.resi_connection == "1conv"
    H, W = lr.shape[2:]
    ws = cfg.window_size
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws                              # :783-788
    x = F.pad(lr, (0, pw, 0, ph), "reflect") if (ph or pw) else lr
    mean = torch.tensor(RGB_MEAN if cfg.in_chans == 3 else (0.0,), dtype=lr.dtype).view(1, -1, 1, 1)
    x = (x - mean) * cfg.img_range                                               # :809-810
    f0 = conv3x3(x, p["conv_first.weight"], p["conv_first.bias"])                # :814
    body = conv3x3(forward_features(f0, p, cfg), p["conv_after_body.weight"], p["conv_after_body.bias"]) + f0  # :815
    y = upsample_tail(body, p, cfg)                                              # :816-817
    y = y / cfg.img_range + mean                                                 # :838
    return y[:, :, :H * cfg.upscale, :W * cfg.upscale]                           # :840


def batch_psnr(pred: Tensor, target: Tensor) -> Tensor:
    """PSNR on clamped [0,1] images: 20*log10(1/sqrt(mse + 1e-8)); finetune_swinir.py:69-74."""
    pred, target = pred.clamp(0, 1), target.clamp(0, 1)
    mse = ((pred - target) ** 2).flatten(1).mean(1)
    return (20.0 * torch.log10(1.0 / torch.sqrt(mse + 1e-8))).mean()


def to_dtype(p: Dict[str, Tensor], dtype) -> Dict[str, Tensor]:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in p.items()}
