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

Tight mode (``SwinIR.set_precision("fp16")`` marks the model and its RSTBs with ``split_conv``): the same layers through ``SplitConv3x3`` /
``SplitPixelShuffleTail`` -- activations and weights as hi / lo fp16 pairs, three products per layer accumulated in fp32 in one
launch (fp32-class accuracy from the fp16 tensor-core kernel; intermediates stay fp32 rows).

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


def rows_split(rows: torch.Tensor, channels: int, *, act: int = L.ACT_NONE, slope: float = 0.0, shuffle=None) -> torch.Tensor:
    """fp32 token rows (..., C) -> the tight mode's fp16 pair image (P, 2 cp) = [lo | hi], cp = 64 * ceil(C / 64)
    (srk_rows_to_f16_split).  shuffle=(H, W): rows are a conv + PixelShuffle(2) stage's 256 channels at H x W; the pair comes out at
    2H x 2W (P = 4 x pixels, cp = 64)."""
    pixels = rows.numel() // rows.shape[-1]
    opix = 4 * pixels if shuffle else pixels
    cp = 64 if shuffle else (channels + 63) // 64 * 64
    out = torch.empty(opix, 2 * cp, dtype=torch.float16, device=rows.device)
    L.rows_to_f16_split(rows, out, channels=channels, ld_in=rows.shape[-1], pixels=pixels, act=act, slope=slope, shuffle=shuffle)
    return out


class SplitConv3x3:
    """FusedConv3x3 for the tight mode: fp32-class accuracy out of the fp16 tensor-core kernel.  Activations and weights are fp16
    pairs (x = hi + lo to 22 bits) and conv(x, w) = lo(x) * hi(w) + hi(x) * lo(w) + hi(x) * hi(w), accumulated in fp32 in ONE
    launch, small terms first: 3 C/64 k-steps over the [lo | hi] image against the weight stream [hi(w) | lo(w) | hi(w)]
    (srk.h: SrkConvDesc.a_atoms)."""

    def __init__(self, conv: nn.Conv2d, *, pixel_shuffle: bool = False, out_scale: float = 1.0, out_shift=None):
        FusedConv3x3(conv)                 # geometry check
        self.conv = conv
        self.kw = dict(pixel_shuffle=pixel_shuffle, out_scale=out_scale, out_shift=out_shift, split=True)
        self._cache = _Cache()

    def packed(self):
        c = self.conv
        return self._cache.get([c.weight, c.bias], lambda: packing.pack_conv3x3(c.weight, c.bias, **self.kw))

    def __call__(self, xs: torch.Tensor, B: int, H: int, W: int, *, out: torch.Tensor, mode: int, ld_out: int,
                 residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        """xs: rows_split(...) of the input.  The activation of the layer (if any) is applied by the NEXT rows_split."""
        ws, bias, meta = self.packed()
        L.conv3x3(xs, ws, bias, out, batch=B, height=H, width=W, k_atoms=meta["k_atoms"], a_atoms=meta["a_atoms"], np_=meta["np"],
                  cout=meta["cout"], out_mode=mode, ld_out=ld_out, residual=residual)
        return out


class SplitPixelShuffleTail:
    """PixelShuffleTail for the tight mode: every intermediate stays fp32 rows, re-split into fp16 pairs between the layers (the
    LeakyReLU and the PixelShuffle(2) permutation ride on that pass)."""

    def __init__(self, conv_before_upsample: nn.Sequential, upsample: nn.Sequential, conv_last: nn.Conv2d, img_range: float, mean):
        ref = PixelShuffleTail(conv_before_upsample, upsample, conv_last, img_range, mean)      # structure checks
        self.slope, self.out_ch = ref.slope, ref.out_ch
        self.before = SplitConv3x3(conv_before_upsample[0])
        self.ups = [SplitConv3x3(u.conv, pixel_shuffle=True) for u in ref.ups]
        self.last = SplitConv3x3(conv_last, **{k: ref.last.kw[k] for k in ("out_scale", "out_shift")})

    def __call__(self, xs: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
        dev = xs.device
        t = torch.empty(B * H * W, 64, dtype=torch.float32, device=dev)
        self.before(xs, B, H, W, out=t, mode=L.CONV_OUT_ROWS_F32, ld_out=64)
        ts = rows_split(t, 64, act=L.ACT_LEAKY_RELU, slope=self.slope)
        for up in self.ups:
            u = torch.empty(B * H * W, 256, dtype=torch.float32, device=dev)
            up(ts, B, H, W, out=u, mode=L.CONV_OUT_ROWS_F32, ld_out=256)
            ts = rows_split(u, 256, shuffle=(H, W))
            H, W = 2 * H, 2 * W
        y = torch.empty(B, H, W, self.out_ch, dtype=torch.float32, device=dev)
        self.last(ts, B, H, W, out=y, mode=L.CONV_OUT_IMAGE, ld_out=self.out_ch)
        return y.permute(0, 3, 1, 2)


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


def group_conv_residual(module: nn.Module, conv: nn.Conv2d, t: torch.Tensor, x: torch.Tensor, x_size) -> torch.Tensor:
    """The tail of RSTB / RHAG / ResidualGroup (network_swinir.py:482, hat_arch.py:620, dat_arch.py:649-651):
    ``x + conv(t)`` with t, x fp32 token rows (B, H*W, C); the result is written over t."""
    B, Ltok, C = t.shape
    t = t.contiguous()
    if getattr(module, "split_conv", False):      # tight mode (set_precision): hi / lo fp16 pairs on the same kernel
        if not hasattr(module, "_sconv"):
            object.__setattr__(module, "_sconv", SplitConv3x3(conv))
        return module._sconv(rows_split(t, C), B, x_size[0], x_size[1], out=t, mode=L.CONV_OUT_ROWS_F32, ld_out=C, residual=x.contiguous())
    if not hasattr(module, "_fconv"):
        object.__setattr__(module, "_fconv", FusedConv3x3(conv))
    return module._fconv(rows_to_f16(t, C), B, x_size[0], x_size[1], out=t, mode=L.CONV_OUT_ROWS_F32, ld_out=C, residual=x.contiguous())


def fused_ok(model: nn.Module) -> bool:
    """Whole-model fused-conv forward: the pixelshuffle tail with 2^n scale, '1conv' residual connections, <= 3 input channels."""
    return (USE_FUSED_CONV and getattr(model, "upsampler", None) == 'pixelshuffle' and isinstance(model.conv_after_body, nn.Conv2d)
            and model.upscale in (2, 4, 8) and model.embed_dim == L.DIM and model.conv_first.in_channels <= 3)


def fused_forward(model: nn.Module, x: torch.Tensor, first_norm: Optional[nn.LayerNorm], run_layers) -> torch.Tensor:
    """SwinIR.forward / HAT.forward / DAT.forward (network_swinir.py:800-840, hat_arch.py:971-994, dat_arch.py:839-858) with every
    3x3 convolution on the implicit-GEMM kernel.  x: (B, C_in, H, W) fp32, already padded to the model's window multiple.
    run_layers(tokens (B, H*W, C), (H, W)) -> tokens runs the residual groups.  Returns the (B, C_out, H*s, W*s) image."""
    B, Cin, H, W = x.shape
    C = model.embed_dim
    dev = x.device
    if not hasattr(model, "_f_first"):
        object.__setattr__(model, "_f_first", FusedConv3x3(model.conv_first, split_first=True))
        object.__setattr__(model, "_f_after", FusedConv3x3(model.conv_after_body))
        object.__setattr__(model, "_f_tail", PixelShuffleTail(model.conv_before_upsample, model.upsample, model.conv_last,
                                                               model.img_range, model.mean))
        object.__setattr__(model, "_mean_vals", [float(v) for v in model.mean.detach().cpu().reshape(-1)])
    x16 = torch.empty(B * H * W, 64, dtype=torch.float16, device=dev)
    L.image_to_f16_split(x, x16, model._mean_vals, model.img_range)                    # (x - mean) * img_range
    feat0 = torch.empty(B, H * W, C, dtype=torch.float32, device=dev)
    model._f_first(x16, B, H, W, out=feat0, mode=L.CONV_OUT_ROWS_F32, ld_out=C)
    t = feat0
    if first_norm is not None:
        t = torch.empty_like(feat0)
        L.layernorm(feat0, t, first_norm.weight, first_norm.bias, num_tokens=B * H * W, ld_in=C, ld_out=C)
    t = run_layers(t, (H, W))
    if getattr(model, "split_conv", False):
        if not hasattr(model, "_s_after"):
            object.__setattr__(model, "_s_after", SplitConv3x3(model.conv_after_body))
            object.__setattr__(model, "_s_tail", SplitPixelShuffleTail(model.conv_before_upsample, model.upsample, model.conv_last,
                                                                        model.img_range, model.mean))
        t = t.contiguous()
        tn = torch.empty_like(feat0)
        L.layernorm(t, tn, model.norm.weight, model.norm.bias, num_tokens=B * H * W, ld_in=C, ld_out=C)
        model._s_after(rows_split(tn, C), B, H, W, out=tn, mode=L.CONV_OUT_ROWS_F32, ld_out=C, residual=feat0)
        return model._s_tail(rows_split(tn, C), B, H, W)
    t16 = torch.empty(B * H * W, L.DIM_PAD, dtype=torch.float16, device=dev)
    L.layernorm_f16(t.contiguous(), t16, model.norm.weight, model.norm.bias, num_tokens=B * H * W, ld_in=C)   # final norm, straight to the conv's layout
    if t.data_ptr() == feat0.data_ptr():
        t = torch.empty_like(feat0)
    model._f_after(t16, B, H, W, out=t, mode=L.CONV_OUT_ROWS_F32, ld_out=C, residual=feat0)      # conv_after_body(...) + x
    return model._f_tail(rows_to_f16(t, C), B, H, W)
