"""GPU parity of the DAT path (dat_arch.py) through the C ABI / drop-in modules against fixtures made from the unmodified
reference.  Gates: whole-model max-abs <= 2e-3 on [0,1] pixels and |dPSNR| <= 0.01 dB (BASELINE.json); module KATs (stress
weights) use relative gates stated per test (bf16 MMA operands, fp32 accumulation)."""
import json
import os

import numpy as np
import pytest
import torch

import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import _lib as L, dat as D
from oracle import synth
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
def model():
    cfg = synth.DAT_CONFIGS["dat_x2_d3"]
    m = srk.DAT(**cfg.as_kwargs()).eval()
    m.load_state_dict(synth.make_dat_state_dict(cfg, seed=99, kind="stress"), strict=True)
    return m.cuda()


def _rel(a, b):
    a, b = a.double().cpu(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _g(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def test_spatial_attention_kats(model):
    """dat_arch.py:363-438: 8x32 / 32x8 windows on the two channel halves, un-shifted and shifted (4,16)/(16,4) with masks."""
    xt = synth.make_tokens(1, 64, 64, 180, seed=11).cuda()
    g = _g("kat_dat_spatial")
    blocks = model.layers[0].blocks
    assert _rel(blocks[0].attn(xt, 64, 64)[:, ::11], g["y_unshifted"]) < 2e-2
    assert _rel(blocks[2].attn(xt, 64, 64)[:, ::11], g["y_shifted"]) < 2e-2


def test_channel_attention_sgfn_block_group_kats(model):
    xt = synth.make_tokens(1, 64, 64, 180, seed=11).cuda()
    blocks = model.layers[0].blocks
    assert _rel(blocks[1].attn(xt, 64, 64)[:, ::11], _g("kat_dat_channel")["y"]) < 2e-2
    assert _rel(blocks[0].ffn(xt, 64, 64)[:, ::11], _g("kat_dat_sgfn")["y"]) < 2e-2
    gb = _g("kat_dat_block")
    for b in range(3):
        assert _rel(blocks[b](xt, (64, 64))[:, ::11], gb[f"y{b}"]) < 1e-2
    assert _rel(model.layers[1](xt, (64, 64))[:, ::11], _g("kat_dat_rg")["y"]) < 1e-2


@pytest.mark.parametrize("name,kind,seed,B,h,w", [("dat_x2_d3", "init", 1234, 1, 64, 64), ("dat_x2_d3", "stress", 4321, 1, 32, 96),
                                                   ("dat_x2_d3", "stress", 77, 1, 40, 72)])      # 40x72: the zero-padded path
def test_whole_model_vs_reference_golden(name, kind, seed, B, h, w):
    cfg = synth.DAT_CONFIGS[name]
    m = srk.DAT(**cfg.as_kwargs()).eval()
    m.load_state_dict(synth.make_dat_state_dict(cfg, seed=seed, kind=kind), strict=True)
    m.cuda()
    lr = synth.make_lr_batch(B, h, w, seed=seed + 1)
    before = L.launch_count()
    y = m(lr.cuda()).cpu()
    assert L.launch_count() > before
    ref = torch.from_numpy(_g(f"{name}_{kind}_{B}x{h}x{w}")["y"])
    assert (y - ref).abs().max().item() <= 2e-3
    hr = torch.nn.functional.interpolate(lr, scale_factor=cfg.upscale, mode="bicubic", align_corners=False)
    assert abs(O.batch_psnr(y, hr).item() - O.batch_psnr(ref, hr).item()) <= 0.01


def test_spatial_attention_padded_path(model):
    """dat_arch.py:376-407: H, W not multiples of 32 -> the projected q, k, v are zero-padded to 64 x 96, the windows and masks are
    those of the padded size and the result is cropped; un-shifted and shifted blocks vs the unmodified reference."""
    xp = synth.make_tokens(2, 40, 72, 180, seed=12).cuda()
    g = _g("kat_dat_spatial_padded")
    blocks = model.layers[0].blocks
    assert _rel(blocks[0].attn(xp, 40, 72)[:, ::7], g["y_unshifted"]) < 2e-2
    assert _rel(blocks[2].attn(xp, 40, 72)[:, ::7], g["y_shifted"]) < 2e-2


def test_wrong_token_count_is_an_error(model):
    xt = synth.make_tokens(1, 48, 64, 180, seed=1).cuda()
    with pytest.raises(RuntimeError):
        model.layers[0].blocks[0].attn(xt, 48, 32)


def test_determinism(model):
    """Two runs are bit-identical (the channel-attention statistics are reduced in a fixed order, no float atomics)."""
    lr = synth.make_lr_batch(2, 64, 64, seed=9).cuda()
    y1, y2 = model(lr), model(lr)
    assert torch.isfinite(y1).all() and torch.equal(y1, y2)
