"""Drop-in HAT modules (reference: ``modules/hat_arch.py``) backed by the sm_100a kernels in libsrk.so.

Same constructor arguments, forward signatures and state_dict keys as the reference: HAT has no
``relative_position_index`` buffer on the attention module -- the index arrives as a forward argument (``rpi``,
hat_arch.py:166) and the top-level model owns ``relative_position_index_SA`` / ``_OCA`` (:784-787).  The kernels evaluate
those indices in closed form, so a passed ``rpi`` is only checked against the standard one.

What runs where (per HAB, hat_arch.py:267-310)
  * LN1 + qkv Linear                      -> srk_linear_fwd          (tcgen05, q/k/v head-pair planes in bf16)
  * roll + partition + softmax(qk^T+rpb+mask)v + reverse + un-roll -> srk_window_attention_fwd (tcgen05, kind HAT_WMSA)
  * proj Linear + shortcut                -> srk_linear_fwd          (bulk reduce-add into the residual stream)
  * CAB conv branch on LN1(x)             -> srk_layernorm_fwd + cuDNN 3x3 convs (library) + srk_cab_gate_add (squeeze-excite
                                             gate fused with the `+ conv_x * conv_scale` residual update)
  * LN2 + Mlp + shortcut                  -> srk_swin_mlp_fwd
OCAB (hat_arch.py:393-439): the same three kernels with kind HAT_OCAB -- the nn.Unfold(24, stride 16, pad 4) of k, v is the
source addressing of the key-window copies.
Inference only; no CPU / eager fallback.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import convs
from . import packing
from .swinir import (Mlp, PatchEmbed, PatchUnEmbed, Upsample, _PackedCache, _conv_tail, _inference_only, _to_2tuple,
                     calculate_mask as _calculate_mask)

WS = 16          # window size served by the kernels
WSE = 24         # OCAB overlapping window: int(16 * 0.5) + 16
_KV_PHASE4 = 0b111111000   # OCAB: the 24-wide key windows start 4 pixels left of a multiple of 8 -> k, v planes use phase 4


def calculate_rpi_sa(window_size: int) -> torch.Tensor:
    """hat_arch.py:882-895 in closed form."""
    t = torch.arange(window_size * window_size)
    y, x = t // window_size, t % window_size
    return (y[:, None] - y[None, :] + window_size - 1) * (2 * window_size - 1) + (x[:, None] - x[None, :] + window_size - 1)


def calculate_rpi_oca(window_size: int, overlap_ratio: float) -> torch.Tensor:
    """hat_arch.py:897-918 in closed form (entries are negative for most pairs; the table lookup wraps)."""
    ws, wse = window_size, window_size + int(overlap_ratio * window_size)
    ti, tj = torch.arange(ws * ws), torch.arange(wse * wse)
    dy = (tj // wse)[None, :] - (ti // ws)[:, None] + ws - wse + 1
    dx = (tj % wse)[None, :] - (ti % ws)[:, None] + ws - wse + 1
    return dy * (ws + wse - 1) + dx


def _planes(n_planes: int, tokens: int, device) -> torch.Tensor:
    return torch.empty((n_planes, tokens, 64), dtype=torch.bfloat16, device=device)


def _check_geometry(dim, num_heads, window_size):
    if dim != L.DIM or num_heads != L.HEADS or window_size != WS:
        raise RuntimeError(f"HAT kernels serve dim 180 / 6 heads / window 16 only (got dim={dim}, heads={num_heads}, "
                           f"window={window_size})")


class ChannelAttention(nn.Module):
    """hat_arch.py:41-59 (RCAN squeeze-excite)."""

    def __init__(self, num_feat, squeeze_factor=16):
        super().__init__()
        self.attention = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(num_feat, num_feat // squeeze_factor, 1, padding=0),
                                       nn.ReLU(inplace=True), nn.Conv2d(num_feat // squeeze_factor, num_feat, 1, padding=0),
                                       nn.Sigmoid())

    def gate(self, x):
        return self.attention(x)

    def forward(self, x):
        return x * self.attention(x)


class CAB(nn.Module):
    """hat_arch.py:62-75: conv3x3 (C -> C/3), GELU, conv3x3 (C/3 -> C), channel attention."""

    def __init__(self, num_feat, compress_ratio=3, squeeze_factor=30):
        super().__init__()
        self.cab = nn.Sequential(nn.Conv2d(num_feat, num_feat // compress_ratio, 3, 1, 1), nn.GELU(),
                                 nn.Conv2d(num_feat // compress_ratio, num_feat, 3, 1, 1),
                                 ChannelAttention(num_feat, squeeze_factor))

    def forward(self, x):
        return self.cab(x)

    def body(self, x):
        """conv3x3 -> GELU -> conv3x3 (cuDNN); the channel attention that follows is fused into srk_cab_gate_add by HAB."""
        return self.cab[2](self.cab[1](self.cab[0](x)))

    def body_nobias(self, x):
        """The same with the library convs run bias-free: conv -> (+bias, GELU in one pass) -> conv; returns (y, bias of the last
        conv), which srk_cab_gate_add folds into its pool and gate."""
        approx = getattr(self.cab[1], "approximate", "none")
        if approx != "none":
            return self.body(x), None
        h = _conv_tail(self.cab[0], x, act=L.ACT_GELU)
        c2 = self.cab[2]
        return F.conv2d(h, c2.weight, None, c2.stride, c2.padding, c2.dilation, c2.groups), c2.bias


class WindowAttention(nn.Module):
    """hat_arch.py:130-197.  forward(x: (nW*B, 256, 180), rpi: (256, 256), mask: (nW, 256, 256) | None)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * self.window_size[0] - 1) * (2 * self.window_size[1] - 1), num_heads))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)
        self._cache = _PackedCache()
        self._rpi_ok = None

    def _packed(self, norm: Optional[nn.LayerNorm] = None):
        if self.window_size != (WS, WS):
            raise RuntimeError(f"WindowAttention(window={self.window_size}): kernels serve 16x16 windows only")
        _check_geometry(self.dim, self.num_heads, WS)
        ps = [self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias, self.relative_position_bias_table] + \
             ([norm.weight, norm.bias] if norm is not None else [])

        def build():
            qw, qb = packing.pack_qkv_planes(self.qkv.weight, self.qkv.bias, None if norm is None else norm.weight,
                                             None if norm is None else norm.bias, self.scale)
            pw, pb = packing.pack_proj_planes(self.proj.weight, self.proj.bias)
            return qw, qb, pw, pb, packing.pack_bias_table_wmsa(self.relative_position_bias_table, WS)
        return self._cache.get(ps, build)

    def _check_rpi(self, rpi):
        if rpi is None:
            return
        key = (rpi.data_ptr(), rpi._version)
        if self._rpi_ok != key:
            if not torch.equal(rpi.detach().cpu().long(), calculate_rpi_sa(WS)):
                raise RuntimeError("WindowAttention: rpi differs from calculate_rpi_sa(); the kernel evaluates the standard "
                                   "relative position index in closed form")
            self._rpi_ok = key

    def attend(self, x_rows, out_rows, *, batch, height, width, shift, norm, ld, apply_ln, add_residual, mask_shift,
               emask=None):
        """qkv -> window attention -> proj on token rows (batch, height*width, ld); out_rows (+)= result."""
        tokens = batch * height * width
        qw, qb, pw, pb, tab = self._packed(norm)
        qkv = _planes(9, tokens, x_rows.device)
        L.linear(x_rows, qw, qb, qkv, num_tokens=tokens, a_mode=L.LIN_A_ROWS, ld_in=ld, apply_ln=apply_ln, n_chunks=3,
                 out_mode=L.LIN_OUT_PLANES)
        o = _planes(3, tokens, x_rows.device)
        L.window_attention(qkv[0:3], qkv[3:6], qkv[6:9], tab, o, kind=L.WA_HAT_WMSA, batch=batch, height=height, width=width,
                           shift=(shift, shift), mask_shift=mask_shift, emask=emask)
        L.linear(o, pw, pb, out_rows, num_tokens=tokens, a_mode=L.LIN_A_PLANES, n_chunks=1, out_mode=L.LIN_OUT_ROWS, ld_out=ld,
                 add_residual=add_residual)

    def forward(self, x, rpi=None, mask=None):
        _inference_only(self)
        self._check_rpi(rpi)
        B_, N, C = x.shape
        if N != WS * WS:
            raise RuntimeError(f"WindowAttention: expected {WS * WS} tokens per window, got {N}")
        x = x.contiguous()
        y = torch.empty_like(x)
        if mask is not None:
            if B_ % mask.shape[0] != 0:
                raise RuntimeError("WindowAttention: B_ must be a multiple of mask.shape[0]")
            mask = mask.to(device=x.device, dtype=torch.float32).contiguous()
        # every window is its own 16x16 "image": window b_ uses mask[b_ % nW] exactly as hat_arch.py:189-191
        self.attend(x, y, batch=B_, height=WS, width=WS, shift=0, norm=None, ld=C, apply_ln=False, add_residual=False,
                    mask_shift=False, emask=mask)
        return y

    def extra_repr(self) -> str:
        return f'dim={self.dim}, window_size={self.window_size}, num_heads={self.num_heads}'


class HAB(nn.Module):
    """hat_arch.py:200-310.  forward(x: (B, H*W, 180), x_size, rpi_sa, attn_mask) -> (B, H*W, 180)."""

    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, compress_ratio=3, squeeze_factor=30,
                 conv_scale=0.01, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, input_resolution, num_heads
        self.window_size, self.shift_size, self.mlp_ratio = window_size, shift_size, mlp_ratio
        if min(self.input_resolution) <= self.window_size:      # :247-250
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, 'shift_size must in 0-window_size'
        if norm_layer is not nn.LayerNorm:
            raise RuntimeError("HAB: only nn.LayerNorm is implemented")
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size=_to_2tuple(self.window_size), num_heads=num_heads, qkv_bias=qkv_bias,
                                    qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.conv_scale = conv_scale
        self.conv_block = CAB(num_feat=dim, compress_ratio=compress_ratio, squeeze_factor=squeeze_factor)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward_into(self, x, x_size, out, rpi_sa=None):
        """out = HAB(x); ``out`` may be ``x`` (in place).  The shift mask is evaluated in closed form inside the kernel
        (it equals HAT.calculate_mask(x_size), hat_arch.py:921-940); HAB ignores the mask when shift_size == 0 (:281-286)."""
        _inference_only(self)
        H, W = x_size
        B, Ltok, C = x.shape
        if Ltok != H * W:
            raise RuntimeError("input feature has wrong size")
        if self.window_size != WS or self.shift_size not in (0, WS // 2):
            raise RuntimeError(f"HAB: window {self.window_size} / shift {self.shift_size} unsupported (16 and 0 or 8)")
        if H % WS or W % WS:
            raise RuntimeError(f"x_size {x_size} must be a multiple of the window size {WS}")
        self.attn._check_rpi(rpi_sa)
        tokens = B * Ltok
        # conv branch on the un-shifted LN1 output (:276-278)
        fused_cab = convs.USE_FUSED_CONV and C == L.DIM and getattr(self.conv_block.cab[1], "approximate", "none") == "none"
        if fused_cab:
            # LN1 straight into the conv's fp16 NHWC layout; conv (180 -> 60) + GELU -> fp16; conv (60 -> 180) + bias -> fp32 rows
            cb = self.conv_block
            if not hasattr(cb, "_f1"):
                object.__setattr__(cb, "_f1", convs.FusedConv3x3(cb.cab[0]))
                object.__setattr__(cb, "_f2", convs.FusedConv3x3(cb.cab[2]))
            xn16 = torch.empty(tokens, L.DIM_PAD, dtype=torch.float16, device=x.device)
            L.layernorm_f16(x, xn16, self.norm1.weight, self.norm1.bias, num_tokens=tokens, ld_in=C)
            mid_cp = cb._f2.cin_pad
            mid = torch.empty(tokens, mid_cp, dtype=torch.float16, device=x.device)
            cb._f1(xn16, B, H, W, out=mid, mode=L.CONV_OUT_NHWC_F16, ld_out=mid_cp, act=L.ACT_GELU)
            y_tok = torch.empty(B, Ltok, C, dtype=torch.float32, device=x.device)
            cb._f2(mid, B, H, W, out=y_tok, mode=L.CONV_OUT_ROWS_F32, ld_out=C)
            y_bias = None
        else:
            xn = torch.empty_like(x)
            L.layernorm(x, xn, self.norm1.weight, self.norm1.bias, num_tokens=tokens, ld_in=C, ld_out=C)
            y, y_bias = self.conv_block.body_nobias(xn.view(B, H, W, C).permute(0, 3, 1, 2))     # channels-last views
        # attention branch (reads x before `out` is touched)
        src = x
        mw, mv = self.mlp._packed(self.norm2)
        qw, qb, pw, pb, tab = self.attn._packed(self.norm1)
        qkv = _planes(9, tokens, x.device)
        L.linear(src, qw, qb, qkv, num_tokens=tokens, a_mode=L.LIN_A_ROWS, ld_in=C, apply_ln=True, n_chunks=3,
                 out_mode=L.LIN_OUT_PLANES)
        o = _planes(3, tokens, x.device)
        L.window_attention(qkv[0:3], qkv[3:6], qkv[6:9], tab, o, kind=L.WA_HAT_WMSA, batch=B, height=H, width=W,
                           shift=(self.shift_size, self.shift_size), mask_shift=self.shift_size > 0)
        if out.data_ptr() != x.data_ptr():
            out.copy_(x)
        L.linear(o, pw, pb, out, num_tokens=tokens, a_mode=L.LIN_A_PLANES, n_chunks=1, out_mode=L.LIN_OUT_ROWS, ld_out=C,
                 add_residual=True)                                                           # out = shortcut + attn
        if not fused_cab:
            y_tok = y.permute(0, 2, 3, 1).reshape(B, Ltok, C).contiguous()                    # a view when y is channels-last
        ca = self.conv_block.cab[3].attention                                                 # squeeze-excite gate + `+ conv_x * conv_scale` (:307)
        L.cab_gate_add(y_tok, out, ca[1].weight.reshape(ca[1].weight.shape[0], C), ca[1].bias, ca[3].weight.reshape(C, -1), ca[3].bias,
                       scale=self.conv_scale, batch=B, tokens_per_image=Ltok, y_bias=y_bias)
        L.swin_mlp(out, out, mw, mv, num_tokens=tokens, ld_in=C, ld_out=C, apply_ln=True, add_residual=True, operands=self.mlp.operands)
        return out

    def forward(self, x, x_size, rpi_sa=None, attn_mask=None):
        x = x.contiguous()
        return self.forward_into(x, x_size, torch.empty_like(x), rpi_sa)


class OCAB(nn.Module):
    """hat_arch.py:353-439 overlapping cross-attention block.  forward(x, x_size, rpi)."""

    def __init__(self, dim, input_resolution, window_size, overlap_ratio, num_heads, qkv_bias=True, qk_scale=None, mlp_ratio=2,
                 norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim, self.input_resolution, self.window_size, self.num_heads = dim, input_resolution, window_size, num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.overlap_win_size = int(window_size * overlap_ratio) + window_size
        self.norm1 = norm_layer(dim)
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.unfold = nn.Unfold(kernel_size=(self.overlap_win_size, self.overlap_win_size), stride=window_size,
                                padding=(self.overlap_win_size - window_size) // 2)
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((window_size + self.overlap_win_size - 1) * (window_size + self.overlap_win_size - 1), num_heads))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)
        self.proj = nn.Linear(dim, dim)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=nn.GELU)
        self._cache = _PackedCache()
        self._rpi_ok = None

    def _packed(self):
        _check_geometry(self.dim, self.num_heads, self.window_size)
        if self.overlap_win_size != WSE:
            raise RuntimeError(f"OCAB: overlap window {self.overlap_win_size} unsupported (kernels serve 24 = 16 + 16 * 0.5)")
        ps = [self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias, self.relative_position_bias_table,
              self.norm1.weight, self.norm1.bias]

        def build():
            qw, qb = packing.pack_qkv_planes(self.qkv.weight, self.qkv.bias, self.norm1.weight, self.norm1.bias, self.scale)
            pw, pb = packing.pack_proj_planes(self.proj.weight, self.proj.bias)
            return qw, qb, pw, pb, packing.pack_bias_table_ocab(self.relative_position_bias_table, WS, WSE)
        return self._cache.get(ps, build)

    def _check_rpi(self, rpi):
        if rpi is None:
            return
        key = (rpi.data_ptr(), rpi._version)
        if self._rpi_ok != key:
            if not torch.equal(rpi.detach().cpu().long(), calculate_rpi_oca(WS, 0.5)):
                raise RuntimeError("OCAB: rpi differs from calculate_rpi_oca(); the kernel evaluates it in closed form")
            self._rpi_ok = key

    def forward_into(self, x, x_size, out, rpi=None):
        _inference_only(self)
        self._check_rpi(rpi)
        H, W = x_size
        B, Ltok, C = x.shape
        if H % WS or W % WS:
            raise RuntimeError(f"x_size {x_size} must be a multiple of the window size {WS}")
        tokens = B * Ltok
        qw, qb, pw, pb, tab = self._packed()
        mw, mv = self.mlp._packed(self.norm2)
        qkv = _planes(9, tokens, x.device)
        L.linear(x, qw, qb, qkv, num_tokens=tokens, a_mode=L.LIN_A_ROWS, ld_in=C, apply_ln=True, n_chunks=3,
                 out_mode=L.LIN_OUT_PLANES, plane_phase_mask=_KV_PHASE4)
        o = _planes(3, tokens, x.device)
        L.window_attention(qkv[0:3], qkv[3:6], qkv[6:9], tab, o, kind=L.WA_HAT_OCAB, batch=B, height=H, width=W)
        if out.data_ptr() != x.data_ptr():
            out.copy_(x)
        L.linear(o, pw, pb, out, num_tokens=tokens, a_mode=L.LIN_A_PLANES, n_chunks=1, out_mode=L.LIN_OUT_ROWS, ld_out=C,
                 add_residual=True)                                                           # proj(x) + shortcut (:436)
        L.swin_mlp(out, out, mw, mv, num_tokens=tokens, ld_in=C, ld_out=C, apply_ln=True, add_residual=True, operands=self.mlp.operands)
        return out

    def forward(self, x, x_size, rpi=None):
        x = x.contiguous()
        return self.forward_into(x, x_size, torch.empty_like(x), rpi)


class AttenBlocks(nn.Module):
    """hat_arch.py:442-535: depth HABs (shift 0 / window_size // 2 alternating) then one OCAB."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, compress_ratio, squeeze_factor, conv_scale,
                 overlap_ratio, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False):
        super().__init__()
        self.dim, self.input_resolution, self.depth, self.use_checkpoint = dim, input_resolution, depth, use_checkpoint
        self.blocks = nn.ModuleList([
            HAB(dim=dim, input_resolution=input_resolution, num_heads=num_heads, window_size=window_size,
                shift_size=0 if (i % 2 == 0) else window_size // 2, compress_ratio=compress_ratio,
                squeeze_factor=squeeze_factor, conv_scale=conv_scale, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                drop=drop, attn_drop=attn_drop, drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                norm_layer=norm_layer) for i in range(depth)])
        self.overlap_attn = OCAB(dim=dim, input_resolution=input_resolution, window_size=window_size,
                                 overlap_ratio=overlap_ratio, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                 mlp_ratio=mlp_ratio, norm_layer=norm_layer)
        if downsample is not None:
            raise RuntimeError("AttenBlocks: downsample (PatchMerging) is never used by HAT and is not implemented")
        self.downsample = None

    def forward(self, x, x_size, params):
        x = x.contiguous()
        out = torch.empty_like(x)
        src = x
        for blk in self.blocks:          # first block out of place (x is the group's residual), the rest in place
            blk.forward_into(src, x_size, out, params.get('rpi_sa'))
            src = out
        self.overlap_attn.forward_into(src, x_size, out, params.get('rpi_oca'))
        return out


class RHAG(nn.Module):
    """hat_arch.py:538-620: AttenBlocks -> 3x3 conv -> + input."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, compress_ratio, squeeze_factor, conv_scale,
                 overlap_ratio, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False, img_size=224, patch_size=4,
                 resi_connection='1conv'):
        super().__init__()
        self.dim, self.input_resolution = dim, input_resolution
        self.residual_group = AttenBlocks(dim=dim, input_resolution=input_resolution, depth=depth, num_heads=num_heads,
                                          window_size=window_size, compress_ratio=compress_ratio, squeeze_factor=squeeze_factor,
                                          conv_scale=conv_scale, overlap_ratio=overlap_ratio, mlp_ratio=mlp_ratio,
                                          qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                                          drop_path=drop_path, norm_layer=norm_layer, downsample=downsample,
                                          use_checkpoint=use_checkpoint)
        if resi_connection == '1conv':
            self.conv = nn.Conv2d(dim, dim, 3, 1, 1)
        elif resi_connection == 'identity':
            self.conv = nn.Identity()
        else:
            raise ValueError(resi_connection)
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=0, embed_dim=dim, norm_layer=None)
        self.patch_unembed = PatchUnEmbed(img_size=img_size, patch_size=patch_size, in_chans=0, embed_dim=dim, norm_layer=None)

    def forward(self, x, x_size, params):
        t = self.residual_group(x, x_size, params)
        if isinstance(self.conv, nn.Conv2d) and convs.USE_FUSED_CONV and t.shape[-1] == L.DIM:
            return convs.group_conv_residual(self, self.conv, t, x, x_size)     # '1conv' on the tcgen05 implicit-GEMM kernel (:620)
        y = self.patch_unembed(t, x_size)
        if isinstance(self.conv, nn.Conv2d):
            return self.patch_embed(_conv_tail(self.conv, y, residual=x))
        return self.patch_embed(self.conv(y)) + x


class HAT(nn.Module):
    """hat_arch.py:710-994.  forward(x: (B, 3, H, W)) -> (B, 3, H*s, W*s)."""

    def __init__(self, img_size=64, patch_size=1, in_chans=3, embed_dim=96, depths=(6, 6, 6, 6), num_heads=(6, 6, 6, 6),
                 window_size=7, compress_ratio=3, squeeze_factor=30, conv_scale=0.01, overlap_ratio=0.5, mlp_ratio=4.,
                 qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1, norm_layer=nn.LayerNorm,
                 ape=False, patch_norm=True, use_checkpoint=False, upscale=2, img_range=1., upsampler='', resi_connection='1conv',
                 **kwargs):
        super().__init__()
        self.window_size, self.shift_size, self.overlap_ratio = window_size, window_size // 2, overlap_ratio
        num_in_ch = num_out_ch = in_chans
        num_feat = 64
        self.img_range = img_range
        self.mean = torch.Tensor((0.4488, 0.4371, 0.4040)).view(1, 3, 1, 1) if in_chans == 3 else torch.zeros(1, 1, 1, 1)
        self.upscale, self.upsampler = upscale, upsampler
        self.register_buffer('relative_position_index_SA', calculate_rpi_sa(window_size))
        self.register_buffer('relative_position_index_OCA', calculate_rpi_oca(window_size, overlap_ratio))
        self.conv_first = nn.Conv2d(num_in_ch, embed_dim, 3, 1, 1)
        self.num_layers, self.embed_dim, self.ape, self.patch_norm = len(depths), embed_dim, ape, patch_norm
        self.num_features, self.mlp_ratio = embed_dim, mlp_ratio
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=embed_dim, embed_dim=embed_dim,
                                      norm_layer=norm_layer if patch_norm else None)
        self.patches_resolution = self.patch_embed.patches_resolution
        self.patch_unembed = PatchUnEmbed(img_size=img_size, patch_size=patch_size, in_chans=embed_dim, embed_dim=embed_dim,
                                          norm_layer=norm_layer if patch_norm else None)
        if ape:
            self.absolute_pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches, embed_dim))
            nn.init.trunc_normal_(self.absolute_pos_embed, std=.02)
        self.pos_drop = nn.Dropout(p=drop_rate)
        res = (self.patches_resolution[0], self.patches_resolution[1])
        self.layers = nn.ModuleList([
            RHAG(dim=embed_dim, input_resolution=res, depth=depths[i], num_heads=num_heads[i], window_size=window_size,
                 compress_ratio=compress_ratio, squeeze_factor=squeeze_factor, conv_scale=conv_scale, overlap_ratio=overlap_ratio,
                 mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate,
                 drop_path=0., norm_layer=norm_layer, downsample=None, use_checkpoint=use_checkpoint, img_size=img_size,
                 patch_size=patch_size, resi_connection=resi_connection) for i in range(self.num_layers)])
        self.norm = norm_layer(self.num_features)
        if resi_connection == '1conv':
            self.conv_after_body = nn.Conv2d(embed_dim, embed_dim, 3, 1, 1)
        elif resi_connection == 'identity':
            self.conv_after_body = nn.Identity()
        if self.upsampler == 'pixelshuffle':
            self.conv_before_upsample = nn.Sequential(nn.Conv2d(embed_dim, num_feat, 3, 1, 1), nn.LeakyReLU(inplace=True))
            self.upsample = Upsample(upscale, num_feat)
            self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)
        else:
            raise RuntimeError("HAT: only upsampler='pixelshuffle' exists in the reference (hat_arch.py:864-869)")
        self.apply(self._init_weights)
        self._channels_last_done = None

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def calculate_rpi_sa(self):
        return calculate_rpi_sa(self.window_size)

    def calculate_rpi_oca(self):
        return calculate_rpi_oca(self.window_size, self.overlap_ratio)

    def calculate_mask(self, x_size):
        """hat_arch.py:921-940 (closed form).  Kept for API parity; the kernels never read this tensor."""
        return _calculate_mask(x_size, self.window_size, self.shift_size)

    def check_image_size(self, x):
        _, _, h, w = x.size()
        ph = (self.window_size - h % self.window_size) % self.window_size
        pw = (self.window_size - w % self.window_size) % self.window_size
        return F.pad(x, (0, pw, 0, ph), 'reflect') if (ph or pw) else x

    def _prepare(self, device):
        if self._channels_last_done != str(device):
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
            self._channels_last_done = str(device)

    def forward_features(self, x):
        x_size = (x.shape[2], x.shape[3])
        # the reference rebuilds the (nW, 256, 256) mask on the CPU and uploads it every forward (:955); here the kernel
        # evaluates the same region ids in registers, so nothing is built
        params = {'attn_mask': None, 'rpi_sa': self.relative_position_index_SA, 'rpi_oca': self.relative_position_index_OCA}
        x = self.patch_embed(x)
        if self.ape:
            x = x + self.absolute_pos_embed
        for layer in self.layers:
            x = layer(x, x_size, params)
        x = x.contiguous()
        B, Ltok, C = x.shape
        L.layernorm(x, x, self.norm.weight, self.norm.bias, num_tokens=B * Ltok, ld_in=C, ld_out=C)
        return self.patch_unembed(x, x_size)

    def invalidate_packed(self) -> None:
        """Forget all packed weight images: needed only after parameters were edited in place through ``.data`` (EMA, weight surgery),
        which changes neither ``_version`` nor ``data_ptr`` (the cache keys).  A ``GraphedModel`` around the model must be ``reset()``."""
        convs.invalidate_all()

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("HAT: CUDA input required (no CPU fallback)")
        _inference_only(self.conv_first)
        if convs.fused_ok(self) and not self.ape and x.dtype == torch.float32:
            H0, W0 = x.shape[2:]
            params = {'attn_mask': None, 'rpi_sa': self.relative_position_index_SA, 'rpi_oca': self.relative_position_index_OCA}

            def run_layers(t, x_size):
                for layer in self.layers:
                    t = layer(t, x_size, params)
                return t

            y = convs.fused_forward(self, self.check_image_size(x), self.patch_embed.norm, run_layers)
            return y[:, :, :H0 * self.upscale, :W0 * self.upscale]
        self._prepare(x.device)
        H, W = x.shape[2:]
        x = self.check_image_size(x)
        self.mean = self.mean.type_as(x)
        x = ((x - self.mean) * self.img_range).contiguous(memory_format=torch.channels_last)
        x = _conv_tail(self.conv_first, x)
        if isinstance(self.conv_after_body, nn.Conv2d):      # bias + long skip in one pass behind the (bias-free) conv
            x = _conv_tail(self.conv_after_body, self.forward_features(x), residual=x)
        else:
            x = self.conv_after_body(self.forward_features(x)) + x
        cbu = self.conv_before_upsample                       # Sequential(conv, LeakyReLU)
        x = _conv_tail(cbu[0], x, act=L.ACT_LEAKY_RELU, slope=cbu[1].negative_slope)
        x = self.conv_last(self.upsample(x))
        x = x / self.img_range + self.mean
        return x[:, :, :H * self.upscale, :W * self.upscale]
