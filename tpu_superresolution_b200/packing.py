"""Weight repack: reference state_dict tensors -> the bf16 slab streams and fp32 vectors the kernels read.

A "slab" is one k-atom of a UMMA B (or A) operand: R rows x 64 K-elements of bf16 = R x 128 bytes, stored
with the 128-byte swizzle (16-byte chunk c of row r at chunk position c ^ (r & 7)), so that one 1-D bulk
TMA copy drops it into shared memory ready for tcgen05.mma (csrc/umma.cuh).  Slabs are concatenated in
the exact order the kernel's producer warp streams them (csrc/swin_kernels.cu).

Folded at pack time: q scale (head_dim**-0.5, network_swinir.py:86,124) and log2(e) for the
exp2-domain softmax into Wq/bq; log2(e) into the relative-position-bias table; head_dim 30 -> 32 and
C 180 -> 192 zero padding; W_proj columns re-indexed to the padded head layout; LayerNorm's gamma into the
columns of the following weight and beta into its bias (the kernels only compute (x - mean) * rstd); the k bias
is dropped (it adds the same q.b_k to every logit of a row, which softmax cancels); the v bias moves into the
proj bias (softmax rows sum to one, so P (V + 1 b_v^T) = P V + b_v^T).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib as L

LOG2E = 1.4426950408889634


def swizzle_slab(mat: torch.Tensor) -> torch.Tensor:
    """(R, 64) -> (R, 64) with the 8-element (16-byte) chunks of each row permuted by c ^ (r & 7)."""
    R = mat.shape[0]
    assert mat.shape[1] == 64 and R % 8 == 0
    chunks = mat.reshape(R, 8, 8)
    r = torch.arange(R, device=mat.device)[:, None]
    c = torch.arange(8, device=mat.device)[None, :]
    out = torch.empty_like(chunks)
    out[r.expand(R, 8), c ^ (r & 7)] = chunks
    return out.reshape(R, 64)


def unswizzle_slab(slab: torch.Tensor) -> torch.Tensor:
    """Inverse of swizzle_slab (the permutation is an involution per row)."""
    return swizzle_slab(slab)


OPERAND_DTYPES = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp16_fast": torch.float16}      # include/srk.h: SRK_OPERANDS_*


def _slabs(mat: torch.Tensor, operands: str = "bf16") -> list:
    """(R, K) with K % 64 == 0 -> list of K/64 swizzled bf16 (or fp16) slabs in k order."""
    if operands != "bf16" and bool((mat.abs() > 65504.0).any()):
        raise RuntimeError("packing: a weight exceeds the fp16 range; use operands='bf16'")
    mat = mat.to(OPERAND_DTYPES[operands])
    return [swizzle_slab(mat[:, k:k + 64].contiguous()) for k in range(0, mat.shape[1], 64)]


def _pad_heads(w: torch.Tensor) -> torch.Tensor:
    """(180, ...) rows in head layout h*30+d -> (192, ...) rows in padded layout h*32+d."""
    out = w.new_zeros((L.HEADS * L.HEAD_PAD,) + tuple(w.shape[1:]))
    out.view(L.HEADS, L.HEAD_PAD, *w.shape[1:])[:, :L.HEAD_DIM] = w.view(L.HEADS, L.HEAD_DIM, *w.shape[1:])
    return out


def _pad_cols(w: torch.Tensor, cols: int) -> torch.Tensor:
    out = w.new_zeros(w.shape[0], cols)
    out[:, :w.shape[1]] = w
    return out


@torch.no_grad()
def pack_attention(qkv_w, qkv_b, proj_w, proj_b, rpb_table, ln_w=None, ln_b=None, scale=None, operands: str = "bf16"
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (wstream uint8[ATTN_WSTREAM_BYTES], vec float32[ATTN_VEC_FLOATS]) on qkv_w's device.

    qkv_w (540,180), qkv_b (540) or None, proj_w (180,180), proj_b (180), rpb_table (225, 6);
    ln_w/ln_b (180) or None for the identity (WindowAttention used stand-alone).
    """
    dev = qkv_w.device
    # pack on the host (dozens of tiny index ops; one H2D copy of the result instead of hundreds of launches)
    qkv_w, qkv_b, proj_w, proj_b, rpb_table, ln_w, ln_b = [None if t is None else t.detach().cpu() for t in
                                                           (qkv_w, qkv_b, proj_w, proj_b, rpb_table, ln_w, ln_b)]
    C = L.DIM
    if tuple(qkv_w.shape) != (3 * C, C) or tuple(proj_w.shape) != (C, C) or tuple(rpb_table.shape) != (225, L.HEADS):
        raise RuntimeError(f"unsupported attention geometry qkv {tuple(qkv_w.shape)} proj {tuple(proj_w.shape)} "
                           f"rpb {tuple(rpb_table.shape)}: kernels serve dim 180 / 6 heads / window 8 only")
    scale = (L.HEAD_DIM ** -0.5) if scale is None else float(scale)
    qkv_w = qkv_w.detach().double()
    qkv_b = torch.zeros(3 * C, dtype=torch.float64) if qkv_b is None else qkv_b.detach().double()
    gamma = torch.ones(C, dtype=torch.float64) if ln_w is None else ln_w.detach().double()
    beta = torch.zeros(C, dtype=torch.float64) if ln_b is None else ln_b.detach().double()
    qkv_b = qkv_b + qkv_w @ beta                     # LN beta -> bias
    qkv_w = qkv_w * gamma[None, :]                   # LN gamma -> weight columns
    wq = _pad_cols(_pad_heads((qkv_w[:C] * (scale * LOG2E)).float()), L.DIM_PAD)          # (192, 192)
    wk = _pad_cols(_pad_heads(qkv_w[C:2 * C].float()), L.DIM_PAD)
    wv = _pad_cols(_pad_heads(qkv_w[2 * C:].float()), L.DIM_PAD)
    bq = _pad_heads((qkv_b[:C] * (scale * LOG2E)).float())
    proj_b_eff = (proj_b.detach().double() + proj_w.detach().double() @ qkv_b[2 * C:]).float()   # v bias -> proj bias

    slabs = _slabs(wv, operands)                         # V = xhat Wv^T: B operand, N = 192 padded v-dims (3 k-atoms x 24 KB)
    for h in range(0, L.HEADS, 2):                       # heads h, h+1: [q_h | k_h | q_h+1 | k_h+1] rows (B operand, N = 128)
        rows = torch.cat([wq[32 * h:32 * h + 32], wk[32 * h:32 * h + 32],
                          wq[32 * h + 32:32 * h + 64], wk[32 * h + 32:32 * h + 64]], 0)
        slabs += _slabs(rows, operands)
    # proj: K index is the padded head layout of O
    wp = proj_w.detach().float().view(C, L.HEADS, L.HEAD_DIM)
    wp_pad = wp.new_zeros(L.DIM_PAD, L.HEADS, L.HEAD_PAD)
    wp_pad[:C, :, :L.HEAD_DIM] = wp
    slabs += _slabs(wp_pad.reshape(L.DIM_PAD, L.DIM_PAD), operands)
    wstream = torch.cat([s.reshape(-1) for s in slabs]).contiguous().view(torch.uint8)
    assert wstream.numel() == L.ATTN_WSTREAM_BYTES

    vec = torch.zeros(L.ATTN_VEC_FLOATS, dtype=torch.float32)
    vec[L.AV_BIAS_Q:L.AV_BIAS_Q + 192] = bq
    vec[L.AV_BIAS_PROJ:L.AV_BIAS_PROJ + C] = proj_b_eff
    rpb = vec[L.AV_RPB:L.AV_RPB + L.HEADS * L.AV_RPB_STRIDE].view(L.HEADS, L.AV_RPB_STRIDE)
    rpb[:, :225] = rpb_table.detach().float().t() * LOG2E
    return wstream.to(dev), vec.to(dev)


@torch.no_grad()
def pack_mlp(fc1_w, fc1_b, fc2_w, fc2_b, ln_w=None, ln_b=None, operands: str = "bf16") -> Tuple[torch.Tensor, torch.Tensor]:
    """fc1 (360,180), fc2 (180,360) -> (wstream uint8[MLP_WSTREAM_BYTES], vec float32[MLP_VEC_FLOATS])."""
    dev = fc1_w.device
    fc1_w, fc1_b, fc2_w, fc2_b, ln_w, ln_b = [None if t is None else t.detach().cpu() for t in
                                              (fc1_w, fc1_b, fc2_w, fc2_b, ln_w, ln_b)]
    C, Hd = L.DIM, L.HIDDEN
    if tuple(fc1_w.shape) != (Hd, C) or tuple(fc2_w.shape) != (C, Hd):
        raise RuntimeError(f"unsupported MLP geometry fc1 {tuple(fc1_w.shape)} fc2 {tuple(fc2_w.shape)}: "
                           "kernels serve dim 180 / mlp_ratio 2 only")
    f1 = fc1_w.detach().double()
    b1 = torch.zeros(Hd, dtype=torch.float64) if fc1_b is None else fc1_b.detach().double()
    if ln_b is not None:
        b1 = b1 + f1 @ ln_b.detach().double()        # LN beta -> bias
    if ln_w is not None:
        f1 = f1 * ln_w.detach().double()[None, :]    # LN gamma -> weight columns
    w1 = fc1_w.new_zeros(L.HIDDEN_PAD, L.DIM_PAD, dtype=torch.float32)
    w1[:Hd, :C] = f1.float()
    w2 = fc2_w.new_zeros(L.DIM_PAD, L.HIDDEN_PAD, dtype=torch.float32)
    w2[:C, :Hd] = fc2_w.detach().float()
    f1 = [_slabs(w1[128 * c:128 * (c + 1)], operands) for c in range(3)]     # fc1 in three 128-unit hidden chunks (3 k-atoms each)
    f2 = _slabs(w2, operands)                                        # fc2: 6 k-atoms of 64 hidden units
    slabs = f1[0] + f1[1] + f2[0:2] + f1[2] + f2[2:6]               # the order the MMA warp consumes them
    wstream = torch.cat([s.reshape(-1) for s in slabs]).contiguous().view(torch.uint8)
    assert wstream.numel() == L.MLP_WSTREAM_BYTES
    vec = torch.zeros(L.MLP_VEC_FLOATS, dtype=torch.float32)
    vec[L.MV_B1:L.MV_B1 + Hd] = b1.float()
    if fc2_b is not None:
        vec[L.MV_B2:L.MV_B2 + C] = fc2_b.detach().float()
    return wstream.to(dev), vec.to(dev)


# ------------------------------------------------------------------------------------------------
# HAT / DAT path: token-linear weight streams (csrc/linear_kernel.cu) and strided bias tables (csrc/winattn_kernel.cu)
# ------------------------------------------------------------------------------------------------
def pack_linear_stream(w_pad: torch.Tensor) -> torch.Tensor:
    """(n_chunks * 192, k_atoms * 64) padded weight -> uint8 stream of n_chunks x k_atoms swizzled slabs (192 x 64 bf16)."""
    N, K = w_pad.shape
    assert N % 192 == 0 and K % 64 == 0
    slabs = []
    for c in range(N // 192):
        slabs += _slabs(w_pad[192 * c:192 * (c + 1)])
    return torch.cat([s.reshape(-1) for s in slabs]).contiguous().view(torch.uint8)


@torch.no_grad()
def pack_qkv_planes(qkv_w, qkv_b, ln_w=None, ln_b=None, scale=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """qkv Linear (540, 180) -> (wstream, bias[576]) producing 9 planes: q heads (0,1) (2,3) (4,5), then k, then v, each
    head padded 30 -> 32.  LayerNorm's affine is folded in; q carries head_dim**-0.5 * log2(e) (hat_arch.py:182, exp2 softmax).
    Unlike the SwinIR kernel the k and v biases stay: OCAB zero-pads the PROJECTED k, v (hat_arch.py:409), so they are not
    row constants there."""
    dev = qkv_w.device
    qkv_w, qkv_b, ln_w, ln_b = [None if t is None else t.detach().cpu().double() for t in (qkv_w, qkv_b, ln_w, ln_b)]
    C = L.DIM
    if tuple(qkv_w.shape) != (3 * C, C):
        raise RuntimeError(f"unsupported qkv geometry {tuple(qkv_w.shape)}: kernels serve dim 180 / 6 heads only")
    scale = (L.HEAD_DIM ** -0.5) if scale is None else float(scale)
    b = torch.zeros(3 * C, dtype=torch.float64) if qkv_b is None else qkv_b.clone()
    w = qkv_w.clone()
    if ln_b is not None:
        b = b + w @ ln_b
    if ln_w is not None:
        w = w * ln_w[None, :]
    w[:C] *= scale * LOG2E
    b[:C] *= scale * LOG2E
    w_pad = torch.cat([_pad_cols(_pad_heads(w[i * C:(i + 1) * C].float()), L.DIM_PAD) for i in range(3)], 0)   # (576, 192)
    b_pad = torch.cat([_pad_heads(b[i * C:(i + 1) * C].float()) for i in range(3)], 0)
    # ones column: padded dim 30 of every v head is the constant 1, so the attention kernel's P v GEMM also returns the
    # softmax row sum (column 30 of O) and the row threads never add up p
    b_pad.view(3, L.HEADS, L.HEAD_PAD)[2, :, L.HEAD_DIM] = 1.0
    return pack_linear_stream(w_pad).to(dev), b_pad.contiguous().to(dev)


@torch.no_grad()
def pack_proj_planes(proj_w, proj_b) -> Tuple[torch.Tensor, torch.Tensor]:
    """proj Linear (180, 180) consuming the attention output planes (K index = head * 32 + d) -> (wstream, bias[192])."""
    dev = proj_w.device
    C = L.DIM
    if tuple(proj_w.shape) != (C, C):
        raise RuntimeError(f"unsupported proj geometry {tuple(proj_w.shape)}")
    wp = proj_w.detach().cpu().float().view(C, L.HEADS, L.HEAD_DIM)
    w_pad = wp.new_zeros(L.DIM_PAD, L.HEADS, L.HEAD_PAD)
    w_pad[:C, :, :L.HEAD_DIM] = wp
    b = torch.zeros(L.DIM_PAD)
    if proj_b is not None:
        b[:C] = proj_b.detach().cpu().float()
    return pack_linear_stream(w_pad.reshape(L.DIM_PAD, L.DIM_PAD)).to(dev), b.to(dev)


@torch.no_grad()
def pack_bias_table_wmsa(table: torch.Tensor, ws: int = 16, sy: int = 48) -> torch.Tensor:
    """(961, nH) relative-position table (hat_arch.py:153-154) -> [nH][31 * sy] floats * log2(e); entry dy * sy + dx holds the
    bias of query-key offset (dy, dx) = (yi - yj + ws - 1, xi - xj + ws - 1).  The padded row stride sy keeps the row threads'
    shared-memory reads bank-conflict free (16 query columns per warp -> stride = 16 mod 32)."""
    n = 2 * ws - 1
    t = table.detach().cpu().float().t().reshape(-1, n, n) * LOG2E
    out = torch.zeros(t.shape[0], n, sy)
    out[:, :, :n] = t
    return out.reshape(t.shape[0], n * sy).contiguous().to(table.device)


@torch.no_grad()
def pack_bias_table_ocab(table: torch.Tensor, ws: int = 16, wse: int = 24, sy: int = 48) -> torch.Tensor:
    """(1521, nH) OCAB table -> [nH][39 * sy] * log2(e) with entry dy * sy + dx for (dy, dx) = (yi - yj + wse - 1, xi - xj + wse - 1).

    hat_arch.py:911-918 indexes the table with ((yj - yi) + ws - wse + 1) * (ws + wse - 1) + ((xj - xi) + ws - wse + 1), which is
    negative for most pairs and wraps (python indexing) -- SURVEY.md A.3; the wrap is applied here, once."""
    n = ws + wse - 1
    d = torch.arange(n)
    e = ws - d                      # (yj - yi) + ws - wse + 1 with yj - yi = wse - 1 - dy
    idx = (e[:, None] * n + e[None, :]) % (n * n)
    t = table.detach().cpu().float()[idx.reshape(-1)].reshape(n, n, -1).permute(2, 0, 1) * LOG2E
    out = torch.zeros(t.shape[0], n, sy)
    out[:, :, :n] = t
    return out.reshape(t.shape[0], n * sy).contiguous().to(table.device)


def unswizzle_planes(planes: torch.Tensor, phase: int = 0) -> torch.Tensor:
    """(P, T, 64) bf16 plane buffer -> logical channel order (inverse of chunk ^ ((tok + phase) & 7)); debugging / tests."""
    P, T, _ = planes.shape
    ch = planes.reshape(P, T, 8, 8)
    tok = torch.arange(T, device=planes.device)
    key = (tok + phase) & 7
    c = torch.arange(8, device=planes.device)
    src = c[None, :] ^ key[:, None]                      # logical chunk c lives at position c ^ key
    return torch.gather(ch, 2, src[None, :, :, None].expand(P, T, 8, 8)).reshape(P, T, 64)


def make_pad_pages(device) -> torch.Tensor:
    """8 KB for srk_window_attention_fwd: 4 KB of zeros (k rows of OCAB's zero padding, hat_arch.py:378) then 32 v padding rows:
    zeros except bf16 1.0 at padded dim 30 of both heads (the ones column), 16-byte chunks permuted by chunk ^ (row & 7)."""
    rows = torch.zeros(1, 32, 64, dtype=torch.bfloat16)
    rows[0, :, L.HEAD_DIM] = 1.0
    rows[0, :, L.HEAD_PAD + L.HEAD_DIM] = 1.0
    v = unswizzle_planes(rows, 0)[0].contiguous().view(torch.uint8).reshape(-1)      # involution: this applies the swizzle
    return torch.cat([torch.zeros(4096, dtype=torch.uint8), v]).to(device)


# ------------------------------------------------------------------------------------------------
# DAT path
# ------------------------------------------------------------------------------------------------
def _fold_ln(w: torch.Tensor, b, ln_w, ln_b):
    """(W, b) of a Linear applied to LayerNorm(x) -> (W', b') applied to (x - mean) * rstd."""
    w = w.detach().cpu().double()
    b = torch.zeros(w.shape[0], dtype=torch.float64) if b is None else b.detach().cpu().double()
    if ln_b is not None:
        b = b + w @ ln_b.detach().cpu().double()
    if ln_w is not None:
        w = w * ln_w.detach().cpu().double()[None, :]
    return w, b


@torch.no_grad()
def pack_rows_linear(w, b, ln_w=None, ln_b=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Linear (N, 180) with N a multiple of 180 -> (wstream, bias) for srk_linear_fwd with fp32 row output: chunk c = output
    columns [180 c, 180 c + 180), each padded to 192 weight rows."""
    dev = w.device
    N, K = w.shape
    if K != L.DIM or N % L.DIM:
        raise RuntimeError(f"unsupported linear geometry {tuple(w.shape)}: K must be 180 and N a multiple of 180")
    wf, bf = _fold_ln(w, b, ln_w, ln_b)
    nch = N // L.DIM
    w_pad = torch.zeros(nch, 192, L.DIM_PAD)
    w_pad[:, :L.DIM, :K] = wf.float().view(nch, L.DIM, K)
    b_pad = torch.zeros(nch, 192)
    b_pad[:, :L.DIM] = bf.float().view(nch, L.DIM)
    return pack_linear_stream(w_pad.reshape(nch * 192, L.DIM_PAD)).to(dev), b_pad.reshape(-1).contiguous().to(dev)


@torch.no_grad()
def pack_planes_linear(w, b) -> Tuple[torch.Tensor, torch.Tensor]:
    """Linear (180, K), K <= 384, consuming bf16 planes in plain channel order (plane = 64 channels; K padded with zero columns to a
    whole number of planes: 3 or 6) -> (wstream, bias[192]) for srk_linear_fwd(a_mode = PLANES, fp32 row output, one chunk)."""
    dev = w.device
    N, K = w.shape
    if N != L.DIM or K > 384:
        raise RuntimeError(f"unsupported linear geometry {tuple(w.shape)}: N must be 180 and K <= 384")
    kp = 192 if K <= 192 else 384
    w_pad = torch.zeros(192, kp)
    w_pad[:N, :K] = w.detach().cpu().float()
    b_pad = torch.zeros(192)
    if b is not None:
        b_pad[:N] = b.detach().cpu().float()
    return pack_linear_stream(w_pad).to(dev), b_pad.to(dev)


@torch.no_grad()
def pack_dat_qkv_planes(qkv_w, qkv_b, ln_w=None, ln_b=None, scale=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """DAT spatial block qkv (540, 180) -> (wstream, bias[768]) producing 12 planes: for each of q, k, v the head pairs
    (h0, h1) (h2, -) of the first channel half (8x32 windows) and (h3, h4) (h5, -) of the second (32x8 windows),
    dat_arch.py:286-291, :409-415.  q carries head_dim**-0.5 * log2(e); dim 30 of every real v head is the ones column."""
    dev = qkv_w.device
    C = L.DIM
    if tuple(qkv_w.shape) != (3 * C, C):
        raise RuntimeError(f"unsupported qkv geometry {tuple(qkv_w.shape)}")
    scale = (L.HEAD_DIM ** -0.5) if scale is None else float(scale)
    w, b = _fold_ln(qkv_w, qkv_b, ln_w, ln_b)
    w[:C] *= scale * LOG2E
    b[:C] *= scale * LOG2E
    slots = [0, 1, 2, None, 3, 4, 5, None]                         # head in each 32-row slot of a 256-row part
    w_pad = torch.zeros(3, 8, L.HEAD_PAD, L.DIM_PAD)
    b_pad = torch.zeros(3, 8, L.HEAD_PAD)
    for part in range(3):
        for s, h in enumerate(slots):
            if h is None:
                continue
            rows = slice(part * C + h * L.HEAD_DIM, part * C + (h + 1) * L.HEAD_DIM)
            w_pad[part, s, :L.HEAD_DIM, :C] = w[rows].float()
            b_pad[part, s, :L.HEAD_DIM] = b[rows].float()
            if part == 2:
                b_pad[part, s, L.HEAD_DIM] = 1.0
    return pack_linear_stream(w_pad.reshape(768, L.DIM_PAD)).to(dev), b_pad.reshape(-1).contiguous().to(dev)


@torch.no_grad()
def pack_bias_table_rect(pos: torch.Tensor, hs: int, ws: int, sy: int) -> torch.Tensor:
    """Dynamic position bias (offsets (2hs-1)(2ws-1), heads 3) (dat_arch.py:219-225) -> [4][(2hs-1) * sy] floats * log2(e),
    entry dy * sy + dx for (dy, dx) = (yi - yj + hs - 1, xi - xj + ws - 1); head slot 3 is padding."""
    nh = pos.shape[1]
    t = pos.detach().cpu().float().t().reshape(nh, 2 * hs - 1, 2 * ws - 1) * LOG2E
    out = torch.zeros(4, 2 * hs - 1, sy)
    out[:nh, :, :2 * ws - 1] = t
    return out.reshape(4, -1).contiguous().to(pos.device)


# ---------------------------------------------------------------------------------------------------------------
# 3x3 convolutions (csrc/conv_kernel.cu)
# ---------------------------------------------------------------------------------------------------------------
def swizzle_slab_any(mat: torch.Tensor) -> torch.Tensor:
    """swizzle_slab for any 2-byte dtype (the conv slabs are fp16)."""
    return swizzle_slab(mat.view(torch.int16)).view(mat.dtype)


@torch.no_grad()
def pack_conv3x3(weight, bias, *, split_first: bool = False, pixel_shuffle: bool = False, out_scale: float = 1.0, out_shift=None,
                 split: bool = False):
    """nn.Conv2d(C_in, C_out, 3, 1, 1) -> (wstream uint8, bias fp32 [np], meta) for srk_conv3x3_fwd.

    wstream = k_atoms x 3 (dx) x 3 (dy) slabs of np rows x 64 input channels, fp16, 128-byte swizzled, in the order the kernel's
    weight producer streams them.  np = C_out rounded up to 32 (16 when C_out <= 4).
    split_first (conv_first, C_in <= 3): input channels [hi(w), hi(w), w - hi(w)] against srk_image_to_f16_split's
        [hi(v), v - hi(v), hi(v)], so the fp16 roundings of the image and of the weights cancel to second order.
    pixel_shuffle (Upsample: conv + nn.PixelShuffle(2), network_swinir.py:584-585): output row (2i + j) * C_out/4 + c <- original
        row 4c + 2i + j, so the kernel's SRK_CONV_OUT_SHUFFLE2_F16 epilogue writes whole 64-channel pixels.
    out_scale / out_shift: y = conv(x) * out_scale + out_shift folded into weights and bias (conv_last: x / img_range + mean).
    split (tight mode, srk_rows_to_f16_split): the weights as the fp16 pair hi(w), lo(w) = w - hi(w), streamed as the k-atoms
        [hi(w) | lo(w) | hi(w)] (each part padded to whole k-atoms) against the activation image [lo(x) | hi(x)]: ONE launch with
        k_atoms = 3 * ceil(C_in / 64) k-steps over a_atoms = 2 * ceil(C_in / 64) input atoms (the last third re-reads hi(x)); the
        two small products come first.
    """
    dev = weight.device
    w = weight.detach().cpu().double() * out_scale
    cout, cin = w.shape[:2]
    b = torch.zeros(cout, dtype=torch.float64) if bias is None else bias.detach().cpu().double() * out_scale
    if out_shift is not None:
        b = b + torch.as_tensor(out_shift, dtype=torch.float64).reshape(-1)
    if tuple(w.shape[2:]) != (3, 3):
        raise RuntimeError(f"pack_conv3x3: 3x3 kernels only, got {tuple(w.shape)}")
    w = w.float()
    if split_first:
        if 3 * cin > 64:
            raise RuntimeError("pack_conv3x3(split_first): at most 21 input channels")
        hi = w.half().float()
        w = torch.cat([hi, hi, w - hi], dim=1)
        cin = 3 * cin
    if pixel_shuffle:
        if cout % 4 or cout != 256:
            raise RuntimeError("pack_conv3x3(pixel_shuffle): C_out must be 4 x 64")
        c4 = cout // 4
        perm = torch.tensor([4 * c + s for s in range(4) for c in range(c4)])       # new row s * c4 + c <- old row 4c + s
        w, b = w[perm], b[perm]
    k_atoms = (cin + 63) // 64
    np_ = 16 if cout <= 4 else ((cout + 31) // 32) * 32
    if np_ > 256 or k_atoms > 4:
        raise RuntimeError(f"pack_conv3x3: unsupported geometry C_in {cin} C_out {cout}")
    if split and split_first:
        raise RuntimeError("pack_conv3x3: split_first already is a hi / lo split")
    wp = torch.zeros(np_, 64 * k_atoms, 3, 3)
    wp[:cout, :cin] = w

    def stream(wf):
        wh = wf.half()
        slabs = [swizzle_slab_any(wh[:, 64 * ka:64 * ka + 64, dy, dx].contiguous())
                 for ka in range(wf.shape[1] // 64) for dx in range(3) for dy in range(3)]
        return torch.cat([s.reshape(-1) for s in slabs]).contiguous().view(torch.uint8).to(dev)

    bp = torch.zeros(np_, dtype=torch.float32)
    bp[:cout] = b.float()
    meta = {"k_atoms": k_atoms, "np": np_, "cout": cout}
    if not split:
        return stream(wp), bp.to(dev), meta
    hi = wp.half().float()
    meta.update(k_atoms=3 * k_atoms, a_atoms=2 * k_atoms)
    return stream(torch.cat([hi, wp - hi, hi], dim=1)), bp.to(dev), meta
