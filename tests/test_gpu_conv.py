"""GPU parity of the tcgen05 implicit-GEMM 3x3 convolution (csrc/conv_kernel.cu, srk_conv3x3_fwd) against torch's fp32 CPU
convolution -- a floating-point kernel, so the reference is the plain PyTorch fp32 op (the reference's own nn.Conv2d,
network_swinir.py:465/720/729/742-745, hat_arch.py:67-72).  fp16 MMA operands, fp32 accumulation: gate 3e-3 of the output's
max magnitude (measured ~5e-4), 2e-5 for the hi/lo-split conv_first."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tpu_superresolution_b200 import _lib as L, packing

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    L.load()
    with torch.no_grad():
        yield


def _rand(shape, seed, scale=1.0):
    return torch.from_numpy(np.random.default_rng(seed).normal(0, scale, size=shape).astype(np.float32))


def _rel(a, b):
    return ((a.double().cpu() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12)).item()


def _to_f16_nhwc(x_nchw, cp):
    """(B, C, H, W) fp32 CPU -> fp16 NHWC (B*H*W, cp) on the GPU through srk_rows_to_f16."""
    B, C, H, W = x_nchw.shape
    rows = x_nchw.permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous()
    ld = (C + 3) // 4 * 4
    if ld != C:
        rows = F.pad(rows, (0, ld - C))
    out = torch.empty(B * H * W, cp, dtype=torch.float16, device="cuda")
    L.rows_to_f16(rows.cuda(), out, channels=C, ld_in=ld, pixels=B * H * W)
    return out


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (1, 20, 27), (1, 8, 8), (3, 40, 72)])
def test_conv180_rows_with_residual(B, H, W):
    """RSTB conv / conv_after_body: 180 -> 180, fp32 rows, out = conv + bias + residual (in place and out of place)."""
    x, w, b = _rand((B, 180, H, W), 1), _rand((180, 180, 3, 3), 2, 0.03), _rand((180,), 3, 0.1)
    res = _rand((B, H * W, 180), 4)
    ref = (F.conv2d(x, w, b, padding=1).permute(0, 2, 3, 1).reshape(B, H * W, 180) + res)
    ws, bp, meta = packing.pack_conv3x3(w.cuda(), b.cuda())
    x16 = _to_f16_nhwc(x, 192)
    assert torch.equal(x16[:, 180:], torch.zeros_like(x16[:, 180:])) and _rel(x16[:, :180].float().cpu(), x.permute(0, 2, 3, 1).reshape(-1, 180)) < 1e-3
    out = torch.empty(B, H * W, 180, device="cuda")
    L.conv3x3(x16, ws, bp, out, batch=B, height=H, width=W, k_atoms=3, np_=192, cout=180, out_mode=L.CONV_OUT_ROWS_F32, ld_out=180,
              residual=res.cuda())
    assert _rel(out, ref) < 3e-3
    inplace = res.cuda().clone()
    L.conv3x3(x16, ws, bp, inplace, batch=B, height=H, width=W, k_atoms=3, np_=192, cout=180, out_mode=L.CONV_OUT_ROWS_F32, ld_out=180,
              residual=inplace)
    assert torch.equal(inplace, out)
    out2 = torch.empty_like(out)
    L.conv3x3(x16, ws, bp, out2, batch=B, height=H, width=W, k_atoms=3, np_=192, cout=180, out_mode=L.CONV_OUT_ROWS_F32, ld_out=180)
    assert _rel(out2, ref - res) < 3e-3


def test_tail_chain_matches_torch():
    """conv_before_upsample (180 -> 64, LeakyReLU 0.01) -> [conv 64 -> 256 + PixelShuffle(2)] x 2 -> conv_last (64 -> 3) with
    x / img_range + mean folded in: network_swinir.py:742-745, :816-817, :838-840."""
    B, H, W = 2, 16, 24
    x = _rand((B, 180, H, W), 11)
    w0, b0 = _rand((64, 180, 3, 3), 12, 0.03), _rand((64,), 13, 0.1)
    w1, b1 = _rand((256, 64, 3, 3), 14, 0.05), _rand((256,), 15, 0.1)
    w2, b2 = _rand((256, 64, 3, 3), 16, 0.05), _rand((256,), 17, 0.1)
    w3, b3 = _rand((3, 64, 3, 3), 18, 0.05), _rand((3,), 19, 0.1)
    mean = torch.tensor([0.4488, 0.4371, 0.4040]).view(1, 3, 1, 1)
    t = F.leaky_relu(F.conv2d(x, w0, b0, padding=1), 0.01)
    t = F.pixel_shuffle(F.conv2d(t, w1, b1, padding=1), 2)
    t = F.pixel_shuffle(F.conv2d(t, w2, b2, padding=1), 2)
    ref = F.conv2d(t, w3, b3, padding=1) / 2.0 + mean

    dev = "cuda"
    p0, p1, p2 = packing.pack_conv3x3(w0.to(dev), b0.to(dev)), packing.pack_conv3x3(w1.to(dev), b1.to(dev), pixel_shuffle=True), \
        packing.pack_conv3x3(w2.to(dev), b2.to(dev), pixel_shuffle=True)
    p3 = packing.pack_conv3x3(w3.to(dev), b3.to(dev), out_scale=0.5, out_shift=mean.reshape(-1))
    x16 = _to_f16_nhwc(x, 192)
    a = torch.empty(B * H * W, 64, dtype=torch.float16, device=dev)
    L.conv3x3(x16, p0[0], p0[1], a, batch=B, height=H, width=W, k_atoms=3, np_=64, cout=64, out_mode=L.CONV_OUT_NHWC_F16, ld_out=64,
              act=L.ACT_LEAKY_RELU, slope=0.01)
    u1 = torch.empty(B * 2 * H * 2 * W, 64, dtype=torch.float16, device=dev)
    L.conv3x3(a, p1[0], p1[1], u1, batch=B, height=H, width=W, k_atoms=1, np_=256, cout=256, out_mode=L.CONV_OUT_SHUFFLE2_F16, ld_out=64)
    u2 = torch.empty(B * 4 * H * 4 * W, 64, dtype=torch.float16, device=dev)
    L.conv3x3(u1, p2[0], p2[1], u2, batch=B, height=2 * H, width=2 * W, k_atoms=1, np_=256, cout=256, out_mode=L.CONV_OUT_SHUFFLE2_F16, ld_out=64)
    y = torch.empty(B, 4 * H, 4 * W, 3, device=dev)
    L.conv3x3(u2, p3[0], p3[1], y, batch=B, height=4 * H, width=4 * W, k_atoms=1, np_=16, cout=3, out_mode=L.CONV_OUT_IMAGE, ld_out=3)
    # stage by stage (so a failure names its layer), then end to end
    r0 = F.leaky_relu(F.conv2d(x, w0, b0, padding=1), 0.01).permute(0, 2, 3, 1).reshape(-1, 64)
    assert _rel(a.float(), r0) < 3e-3
    r1 = F.pixel_shuffle(F.conv2d(r0.reshape(B, H, W, 64).permute(0, 3, 1, 2), w1, b1, padding=1), 2).permute(0, 2, 3, 1).reshape(-1, 64)
    assert _rel(u1.float(), r1) < 4e-3
    assert _rel(y.permute(0, 3, 1, 2), ref) < 5e-3


def test_conv_first_split_is_near_fp32():
    """conv_first (3 -> 180) on (x - mean) * range with the hi / lo split of image and weights: fp32-grade result."""
    B, H, W = 2, 24, 40
    img = torch.from_numpy(np.random.default_rng(5).random((B, 3, H, W), dtype=np.float32))
    mean = [0.4488, 0.4371, 0.4040]
    w, b = _rand((180, 3, 3, 3), 6, 0.2), _rand((180,), 7, 0.1)
    ref = F.conv2d((img - torch.tensor(mean).view(1, 3, 1, 1)) * 1.0, w, b, padding=1).permute(0, 2, 3, 1).reshape(B, H * W, 180)
    ws, bp, meta = packing.pack_conv3x3(w.cuda(), b.cuda(), split_first=True)
    x16 = torch.empty(B * H * W, 64, dtype=torch.float16, device="cuda")
    L.image_to_f16_split(img.cuda().contiguous(memory_format=torch.channels_last), x16, mean, 1.0)       # any strides
    out = torch.empty(B, H * W, 180, device="cuda")
    L.conv3x3(x16, ws, bp, out, batch=B, height=H, width=W, k_atoms=1, np_=192, cout=180, out_mode=L.CONV_OUT_ROWS_F32, ld_out=180)
    assert _rel(out, ref) < 2e-5


def test_split_convs_are_fp32_grade():
    """Tight mode (convs.SplitConv3x3 / SplitPixelShuffleTail): hi / lo fp16 pairs of activations and weights on the same kernel.
    Reference: torch float64 convolutions; gate 3e-6 of the output's max magnitude for one layer, 2e-5 for the four-layer tail
    (plain fp16 operands: ~5e-4 per layer)."""
    import torch.nn as nn
    from tpu_superresolution_b200 import convs
    B, H, W = 2, 24, 40
    dev = "cuda"
    g = torch.Generator().manual_seed(3)
    body = nn.Conv2d(180, 180, 3, 1, 1)
    before = nn.Sequential(nn.Conv2d(180, 64, 3, 1, 1), nn.LeakyReLU(0.01))
    ups = nn.Sequential(nn.Conv2d(64, 256, 3, 1, 1), nn.PixelShuffle(2), nn.Conv2d(64, 256, 3, 1, 1), nn.PixelShuffle(2))
    last = nn.Conv2d(64, 3, 3, 1, 1)
    for mod in (body, before, ups, last):
        for prm in mod.parameters():
            prm.data = torch.randn(prm.shape, generator=g) * (0.03 if prm.dim() == 4 else 0.1)
    x, res = _rand((B, 180, H, W), 31), _rand((B, H * W, 180), 32)
    mean = torch.tensor([0.4488, 0.4371, 0.4040])
    with torch.no_grad():
        d = lambda m: m.double()
        ref_body = d(body)(x.double()).permute(0, 2, 3, 1).reshape(B, H * W, 180) + res.double()
        ref_tail = d(last)(d(ups)(d(before)(x.double()))) / 2.0 + mean.double().view(1, 3, 1, 1)
        for mod in (body, before, ups, last):
            mod.float().to(dev)
    rows = x.permute(0, 2, 3, 1).reshape(B * H * W, 180).contiguous().to(dev)
    xs = convs.rows_split(rows, 180)
    assert xs.shape == (B * H * W, 384) and not xs[:, 180:192].any() and not xs[:, 372:].any()          # [lo | hi], zero padded
    assert _rel(xs[:, :192].float() + xs[:, 192:].float(), F.pad(rows, (0, 12)).cpu()) < 1e-6
    out = res.to(dev).clone()
    convs.SplitConv3x3(body)(xs, B, H, W, out=out, mode=L.CONV_OUT_ROWS_F32, ld_out=180, residual=out)        # in place += like the RSTB tail
    assert _rel(out, ref_body) < 3e-6
    y = convs.SplitPixelShuffleTail(before, ups, last, 2.0, mean)(xs, B, H, W)
    assert y.shape == (B, 3, 4 * H, 4 * W) and _rel(y, ref_tail) < 2e-5            # four layers deep (measured 8e-6)


def test_cab_pair_with_gelu():
    """hat_arch.py:67-72: conv 180 -> 60, GELU, conv 60 -> 180."""
    B, H, W = 1, 32, 48
    x = _rand((B, 180, H, W), 21)
    w1, b1, w2, b2 = _rand((60, 180, 3, 3), 22, 0.03), _rand((60,), 23, 0.1), _rand((180, 60, 3, 3), 24, 0.05), _rand((180,), 25, 0.1)
    ref = F.conv2d(F.gelu(F.conv2d(x, w1, b1, padding=1)), w2, b2, padding=1).permute(0, 2, 3, 1).reshape(B, H * W, 180)
    p1, p2 = packing.pack_conv3x3(w1.cuda(), b1.cuda()), packing.pack_conv3x3(w2.cuda(), b2.cuda())
    assert p1[2] == {"k_atoms": 3, "np": 64, "cout": 60} and p2[2] == {"k_atoms": 1, "np": 192, "cout": 180}
    x16 = _to_f16_nhwc(x, 192)
    mid = torch.empty(B * H * W, 64, dtype=torch.float16, device="cuda")
    L.conv3x3(x16, p1[0], p1[1], mid, batch=B, height=H, width=W, k_atoms=3, np_=64, cout=60, out_mode=L.CONV_OUT_NHWC_F16, ld_out=64, act=L.ACT_GELU)
    assert torch.equal(mid[:, 60:], torch.zeros_like(mid[:, 60:]))          # padded channels stay exactly zero (the next conv's K padding)
    out = torch.empty(B, H * W, 180, device="cuda")
    L.conv3x3(mid, p2[0], p2[1], out, batch=B, height=H, width=W, k_atoms=1, np_=192, cout=180, out_mode=L.CONV_OUT_ROWS_F32, ld_out=180)
    assert _rel(out, ref) < 4e-3


def test_bad_arguments_are_errors():
    x16 = torch.zeros(64 * 64, 192, dtype=torch.float16, device="cuda")
    ws, bp, _ = packing.pack_conv3x3(torch.zeros(180, 180, 3, 3, device="cuda"), None)
    out = torch.empty(1, 4096, 180, device="cuda")
    with pytest.raises(RuntimeError):
        L.conv3x3(x16, ws, bp, out, batch=1, height=64, width=64, k_atoms=3, np_=192, cout=180, out_mode=7, ld_out=180)
    with pytest.raises(RuntimeError):
        L.conv3x3(x16, ws, bp, out, batch=1, height=64, width=64, k_atoms=3, np_=192, cout=180, out_mode=L.CONV_OUT_ROWS_F32, ld_out=178)
    with pytest.raises(RuntimeError):
        L.conv3x3(x16.float(), ws, bp, out, batch=1, height=64, width=64, k_atoms=3, np_=192, cout=180, out_mode=L.CONV_OUT_ROWS_F32, ld_out=180)
    with pytest.raises(RuntimeError):      # k-steps may wrap over the input atoms once only (a_atoms <= k_atoms <= 2 a_atoms)
        L.conv3x3(x16[:, :64].contiguous(), ws, bp, out, batch=1, height=64, width=64, k_atoms=3, a_atoms=1, np_=192, cout=180,
                  out_mode=L.CONV_OUT_ROWS_F32, ld_out=180)
