"""The 3x3 convolutions of SwinIR / HAT / DAT on the tcgen05 implicit-GEMM kernel (csrc/conv_kernel.cu, srk_conv3x3_fwd).

The drop-in models keep their ``nn.Conv2d`` modules (same state_dict keys as the reference); a ``FusedConv3x3`` wraps one of
them, packs its weights once per weight version (packing.pack_conv3x3) and runs it on fp16 NHWC activations with the layer's
tail fused into the epilogue:

  * body convs (RSTB / RHAG / ResidualGroup conv, conv_after_body; network_swinir.py:465, :729): fp32 token rows out,
    ``+ residual`` fused (network_swinir.py:482, :829);
  * conv_first (:720) on ``(x - mean) * img_range`` (:803-804) with the fp16 rounding of image and weights compensated;
  * the pixelshuffle tail (:742-745, :816-817): conv_before_upsample + LeakyReLU -> fp16, [conv + PixelShuffle(2)] x log2(s) in
    one kernel each, conv_last with ``/ img_range + mean`` (:838) folded into its weights -> the fp32 image;
  * HAT's CAB (hat_arch.py:67-72): conv -> GELU -> conv.

``SRK_CONV=cudnn`` selects the round-1 path (library convolutions + separate bias / activation / shuffle passes) for A/B runs.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import packing

USE_FUSED_CONV = os.environ.get("SRK_CONV", "fused") != "cudnn"


class _Cache:
    """Packed weights keyed on (data_ptr, _version, device) of the source parameters plus an explicit generation counter
    (``invalidate_packed()`` on the models bumps it: in-place updates through ``.data`` do not change ``_version``)."""
    generation = 0

    def __init__(self):
        self.key, self.value = None, None

    def get(self, params, build):
        key = (_Cache.generation,) + tuple((p.data_ptr(), p._version, str(p.device)) for p in params if p is not None)
        if key != self.key:
            self.value, self.key = build(), key
        return self.value


def invalidate_all() -> None:
    """Forget every packed weight image (convolutions, attention, MLP, DAT tables): call after editing parameters in place
    through ``.data`` (EMA updates, weight surgery) -- those edits bump neither ``_version`` nor ``data_ptr``."""
    _Cache.generation += 1


class FusedConv3x3:
    """conv = FusedConv3x3(nn.Conv2d(cin, cout, 3, 1, 1), ...); conv(x16, B, H, W, out=..., mode=..., ...)."""

    def __init__(self, conv: nn.Conv2d, *, split_first: bool = False, pixel_shuffle: bool = False, out_scale: float = 1.0, out_shift=None):
        if not isinstance(conv, nn.Conv2d) or conv.kernel_size != (3, 3) or conv.stride != (1, 1) or conv.padding != (1, 1) or \
                conv.dilation != (1, 1) or conv.groups != 1 or conv.padding_mode != "zeros":
            raise RuntimeError(f"FusedConv3x3: only Conv2d(k=3, s=1, p=1, groups=1, zero padding) is implemented, got {conv}")
        self.conv = conv
        self.kw = dict(split_first=split_first, pixel_shuffle=pixel_shuffle, out_scale=out_scale, out_shift=out_shift)
        self._cache = _Cache()

    def packed(self):
        c = self.conv
        return self._cache.get([c.weight, c.bias], lambda: packing.pack_conv3x3(c.weight, c.bias, **self.kw))

    @property
    def cin_pad(self) -> int:
        return 64 * self.packed()[2]["k_atoms"]

    def __call__(self, x16: torch.Tensor, B: int, H: int, W: int, *, out: torch.Tensor, mode: int, ld_out: int, act: int = L.ACT_NONE,
                 slope: float = 0.0, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        ws, bias, meta = self.packed()
        L.conv3x3(x16, ws, bias, out, batch=B, height=H, width=W, k_atoms=meta["k_atoms"], np_=meta["np"], cout=meta["cout"],
                  out_mode=mode, ld_out=ld_out, act=act, slope=slope, residual=residual)
        return out


def rows_to_f16(rows: torch.Tensor, channels: int) -> torch.Tensor:
    """fp32 token rows (..., C) -> fp16 NHWC (pixels, 64 * ceil(C / 64)), zero padded (srk_rows_to_f16)."""
    pixels = rows.numel() // rows.shape[-1]
    cp = (channels + 63) // 64 * 64
    out = torch.empty(pixels, cp, dtype=torch.float16, device=rows.device)
    L.rows_to_f16(rows, out, channels=channels, ld_in=rows.shape[-1], pixels=pixels)
    return out


class PixelShuffleTail:
    """conv_before_upsample (+ LeakyReLU) -> Upsample([conv, PixelShuffle(2)] x n) -> conv_last -> x / img_range + mean
    (network_swinir.py:742-745, :816-817, :838; identical in hat_arch.py:864-869, :989-992 and dat_arch.py:806-811, :848-858)."""

    def __init__(self, conv_before_upsample: nn.Sequential, upsample: nn.Sequential, conv_last: nn.Conv2d, img_range: float, mean):
        act = conv_before_upsample[1]
        if not isinstance(act, nn.LeakyReLU):
            raise RuntimeError("PixelShuffleTail: conv_before_upsample must be Sequential(Conv2d, LeakyReLU)")
        self.slope = act.negative_slope
        self.before = FusedConv3x3(conv_before_upsample[0])
        ups = [m for m in upsample if isinstance(m, nn.Conv2d)]
        if any(c.out_channels != 4 * c.in_channels or c.in_channels != 64 for c in ups):
            raise RuntimeError("PixelShuffleTail: only [conv 64 -> 256, PixelShuffle(2)] stages (scale 2^n, num_feat 64) are implemented")
        self.ups = [FusedConv3x3(c, pixel_shuffle=True) for c in ups]
        m = [float(v) for v in torch.as_tensor(mean).reshape(-1)]
        self.last = FusedConv3x3(conv_last, out_scale=1.0 / img_range, out_shift=(m + m * conv_last.out_channels)[:conv_last.out_channels])
        self.out_ch = conv_last.out_channels

    def __call__(self, x16: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
        """x16: fp16 NHWC (B*H*W, 192) features -> (B, out_ch, H*s, W*s) fp32 image (channels-last memory)."""
        dev = x16.device
        t = torch.empty(B * H * W, 64, dtype=torch.float16, device=dev)
        self.before(x16, B, H, W, out=t, mode=L.CONV_OUT_NHWC_F16, ld_out=64, act=L.ACT_LEAKY_RELU, slope=self.slope)
        for up in self.ups:
            u = torch.empty(B * 4 * H * W, 64, dtype=torch.float16, device=dev)
            up(t, B, H, W, out=u, mode=L.CONV_OUT_SHUFFLE2_F16, ld_out=64)
            t, H, W = u, 2 * H, 2 * W
        y = torch.empty(B, H, W, self.out_ch, dtype=torch.float32, device=dev)
        self.last(t, B, H, W, out=y, mode=L.CONV_OUT_IMAGE, ld_out=self.out_ch)
        return y.permute(0, 3, 1, 2)
