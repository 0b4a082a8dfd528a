"""GPU parity of the HAT path (hat_arch.py) through the C ABI / drop-in modules against fixtures made from the unmodified
reference and against the CPU oracle.  Gates: whole-model max-abs <= 2e-3 on [0,1] pixels and |dPSNR| <= 0.01 dB (BASELINE.json,
bf16 MMA operands with fp32 accumulation); module KATs (stress weights, logit std ~ 3) use relative gates stated per test."""
import os

import numpy as np
import pytest
import torch

import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import _lib as L, packing, hat as H
from oracle import synth
from oracle import hat_oracle as HO
from oracle import swinir_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    L.load()
    torch.backends.cudnn.allow_tf32 = True
    with torch.no_grad():
        yield


@pytest.fixture(scope="module")
def sd():
    return synth.make_hat_state_dict(synth.HAT_CONFIGS["hat_x4_d2"], seed=99, kind="stress")


def _rel(a, b):
    a, b = a.double().cpu(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _sub(sd, pre):
    return {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}


@pytest.mark.parametrize("ntok", [1, 128, 300, 128 * 149 + 77])
def test_linear_qkv_planes(sd, ntok):
    """srk_linear_fwd, fp32 rows + LayerNorm -> 9 bf16 planes, both swizzle phases, ragged last tile."""
    pre = "layers.0.residual_group.blocks.1."
    x = synth.make_tokens(1, 1, ntok, 180, seed=3)[0]
    ref = O.layer_norm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"]) @ sd[pre + "attn.qkv.weight"].T + sd[pre + "attn.qkv.bias"]
    ref[:, :180] *= 30 ** -0.5 * packing.LOG2E
    qw, qb = packing.pack_qkv_planes(sd[pre + "attn.qkv.weight"].cuda(), sd[pre + "attn.qkv.bias"].cuda(),
                                     sd[pre + "norm1.weight"].cuda(), sd[pre + "norm1.bias"].cuda())
    for mask in (0, H._KV_PHASE4):
        planes = torch.zeros(9, ntok, 64, dtype=torch.bfloat16, device="cuda")
        L.linear(x.cuda(), qw, qb, planes, num_tokens=ntok, a_mode=L.LIN_A_ROWS, ld_in=180, apply_ln=True, n_chunks=3,
                 out_mode=L.LIN_OUT_PLANES, plane_phase_mask=mask)
        got = torch.cat([packing.unswizzle_planes(planes[p:p + 1], 4 if (mask >> p) & 1 else 0)[0] for p in range(9)], 1).float()
        pad = got.view(ntok, 3, 6, 32)[..., 30:].clone()
        assert bool((pad[:, 2, :, 0] == 1.0).all())                # ones column of v (softmax row sum via the P v GEMM)
        pad[:, 2, :, 0] = 0.0
        assert float(pad.abs().max()) == 0.0                       # the other padded head dims stay exactly zero
        assert _rel(got.view(ntok, 3, 6, 32)[..., :30].reshape(ntok, 540), ref) < 1.5e-2


@pytest.mark.parametrize("ntok", [5, 128, 128 * 150 + 1])
def test_linear_proj_rows(sd, ntok):
    """srk_linear_fwd, bf16 planes -> fp32 rows, plain and bulk reduce-add residual."""
    pre = "layers.0.residual_group.blocks.1."
    o = torch.from_numpy(np.random.default_rng(5).normal(0, 1, size=(ntok, 6, 30)).astype(np.float32))
    ref = o.reshape(ntok, 180) @ sd[pre + "attn.proj.weight"].T + sd[pre + "attn.proj.bias"]
    opad = torch.zeros(ntok, 6, 32)
    opad[..., :30] = o
    planes = packing.unswizzle_planes(opad.reshape(ntok, 3, 64).permute(1, 0, 2).contiguous().to(torch.bfloat16).cuda(), 0)
    pw, pb = packing.pack_proj_planes(sd[pre + "attn.proj.weight"].cuda(), sd[pre + "attn.proj.bias"].cuda())
    base = synth.make_tokens(1, 1, ntok, 180, seed=8)[0]
    for res in (False, True):
        y = base.clone().cuda()
        L.linear(planes, pw, pb, y, num_tokens=ntok, a_mode=L.LIN_A_PLANES, n_chunks=1, out_mode=L.LIN_OUT_ROWS, ld_out=180,
                 add_residual=res)
        assert _rel(y, ref + (base if res else 0)) < 1.5e-2


def test_window_attention_kat(sd):
    """hat_arch.py:166-197 with stress weights, with the rpi argument and an explicit (nW, 256, 256) mask."""
    g = np.load(os.path.join(GOLDEN, "kat_hat_window_attention.npz"))
    attn = H.WindowAttention(180, (16, 16), 6).eval()
    attn.load_state_dict(_sub(sd, "layers.0.residual_group.blocks.1.attn."), strict=True)
    attn.cuda()
    xw = synth.make_tokens(4, 16, 16, 180, seed=5).cuda()
    assert _rel(attn(xw, H.calculate_rpi_sa(16))[:, ::3], g["y_nomask"]) < 3e-2
    mask = torch.from_numpy(g["mask"].astype(np.float32)).cuda()
    assert _rel(attn(xw, H.calculate_rpi_sa(16).cuda(), mask)[:, ::3], g["y_mask"]) < 3e-2
    with pytest.raises(RuntimeError):
        attn(xw, H.calculate_rpi_sa(16) + 1)                       # a non-standard index is an error, not silently ignored


def test_hab_ocab_rhag_kats(sd):
    cfg = synth.HAT_CONFIGS["hat_x4_d2"]
    xt = synth.make_tokens(2, 32, 48, 180, seed=11).cuda()
    g = np.load(os.path.join(GOLDEN, "kat_hat_hab.npz"))
    for b, shift, key in ((0, 0, "y_unshifted"), (1, 8, "y_shifted")):
        blk = H.HAB(180, (64, 64), 6, window_size=16, shift_size=shift, mlp_ratio=2.0).eval()
        blk.load_state_dict(_sub(sd, f"layers.0.residual_group.blocks.{b}."), strict=True)
        blk.cuda()
        assert _rel(blk(xt, (32, 48), None, None)[:, ::5], g[key]) < 1e-2
    blk = H.OCAB(180, (64, 64), 16, 0.5, 6, mlp_ratio=2).eval()
    blk.load_state_dict(_sub(sd, "layers.0.residual_group.overlap_attn."), strict=True)
    blk.cuda()
    assert _rel(blk(xt, (32, 48), H.calculate_rpi_oca(16, 0.5))[:, ::5], np.load(os.path.join(GOLDEN, "kat_hat_ocab.npz"))["y"]) < 1e-2
    r = H.RHAG(180, (64, 64), 2, 6, 16, 3, 30, 0.01, 0.5, mlp_ratio=2.0, img_size=64, patch_size=1).eval()
    r.load_state_dict(_sub(sd, "layers.1."), strict=True)
    r.cuda()
    params = {"rpi_sa": None, "rpi_oca": None, "attn_mask": None}
    assert _rel(r(xt[:1], (32, 48), params)[:, ::3], np.load(os.path.join(GOLDEN, "kat_hat_rhag.npz"))["y"]) < 1e-2


def test_shifted_windows_edge_sizes(sd):
    """One window per axis (16 x 16 image: every key region differs) and a wide strip, shifted, vs the oracle."""
    cfg = synth.HAT_CONFIGS["hat_x4_d2"]
    for (h, w) in [(16, 16), (16, 64), (48, 16)]:
        xt = synth.make_tokens(1, h, w, 180, seed=21)
        blk = H.HAB(180, (64, 64), 6, window_size=16, shift_size=8, mlp_ratio=2.0).eval()
        blk.load_state_dict(_sub(sd, "layers.0.residual_group.blocks.1."), strict=True)
        blk.cuda()
        ref = HO.hab(xt, (h, w), sd, "layers.0.residual_group.blocks.1.", 6, 16, 8, cfg.conv_scale)
        assert _rel(blk(xt.cuda(), (h, w)), ref) < 1e-2


@pytest.mark.parametrize("name,kind,seed,B,h,w", [("hat_x4_d2", "init", 1234, 1, 64, 64), ("hat_x4_d2", "stress", 4321, 1, 32, 48),
                                                   ("hat_x2_d2", "stress", 77, 1, 20, 27)])
def test_whole_model_vs_reference_golden(name, kind, seed, B, h, w):
    cfg = synth.HAT_CONFIGS[name]
    m = srk.HAT(**cfg.as_kwargs()).eval()
    m.load_state_dict(synth.make_hat_state_dict(cfg, seed=seed, kind=kind), strict=True)
    m.cuda()
    lr = synth.make_lr_batch(B, h, w, seed=seed + 1)
    before = L.launch_count()
    y = m(lr.cuda()).cpu()
    assert L.launch_count() > before                              # the CUDA path ran
    ref = torch.from_numpy(np.load(os.path.join(GOLDEN, f"{name}_{kind}_{B}x{h}x{w}.npz"))["y"])
    assert (y - ref).abs().max().item() <= 2e-3
    hr = torch.nn.functional.interpolate(lr, scale_factor=cfg.upscale, mode="bicubic", align_corners=False)
    assert abs(O.batch_psnr(y, hr).item() - O.batch_psnr(ref, hr).item()) <= 0.01


def test_batch8_determinism_and_batch_independence():
    """BASELINE configs[2] shape (B = 8 of 64x64): two runs are bit-identical and each tile equals its own B = 1 run."""
    cfg = synth.HAT_CONFIGS["hat_x4_d2"]
    m = srk.HAT(**cfg.as_kwargs()).eval()
    m.load_state_dict(synth.make_hat_state_dict(cfg, seed=5, kind="stress"), strict=True)
    m.cuda()
    lr = synth.make_lr_batch(8, 64, 64, seed=9).cuda()
    y1, y2 = m(lr), m(lr)
    assert torch.isfinite(y1).all() and torch.equal(y1, y2)
    y0 = m(lr[3:4])
    assert (y0 - y1[3:4]).abs().max().item() <= 1e-3               # cuDNN picks other (TF32) conv algorithms for B = 1; measured 2e-4
