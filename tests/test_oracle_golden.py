"""Pin the CPU oracle restatement to outputs of the unmodified reference (tests/golden)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import swinir_oracle as O
from oracle import synth
from oracle.reference_loader import reference_available, load_reference_module

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-5   # fp32 vs fp32 with a different op order; fp32-vs-fp64 noise floor is ~4e-7


def _close(a, b, tol=TOL):
    a = a.detach().numpy() if torch.is_tensor(a) else a
    err = np.abs(a.astype(np.float64) - b.astype(np.float64)).max()
    assert err <= tol, f"max abs err {err:.3e} > {tol}"


@pytest.mark.parametrize("name", ["swinir_x2", "swinir_x4"])
def test_manifest_matches_reference_state_dict(name):
    with open(os.path.join(GOLDEN, f"{name}_manifest.json")) as f:
        man = json.load(f)
    cfg = synth.CONFIGS[name]
    sd = synth.make_swinir_state_dict(cfg, seed=1, kind="init")
    assert [m[0] for m in man] == list(sd.keys())
    for k, shape, dtype in man:
        assert list(sd[k].shape) == shape, k
        assert str(sd[k].dtype).replace("torch.", "") == dtype, k


def test_closed_forms_match_reference_buffers(golden):
    g = golden("kat_buffers")
    assert np.array_equal(O.relative_position_index(8).numpy(), g["rpi"])
    assert np.array_equal(O.shift_attention_mask(64, 64, 8, 4).numpy().astype(np.int8), g["mask_64"])
    assert np.array_equal(O.shift_attention_mask(48, 40, 8, 4).numpy().astype(np.int8), g["mask_48x40"])
    m = O.shift_attention_mask(64, 64, 8, 4)
    nz = sorted(int(i) for i in torch.nonzero(m.flatten(1).abs().sum(1)).flatten())
    assert nz == sorted(set(range(7, 64, 8)) | set(range(56, 64)))


def test_partition_reverse_roundtrip():
    x = torch.arange(2 * 16 * 24 * 3, dtype=torch.float32).reshape(2, 16, 24, 3)
    w = O.window_partition(x, 8)
    assert w.shape == (12, 8, 8, 3)
    assert torch.equal(w[1], x[0, 0:8, 8:16])
    assert torch.equal(O.window_reverse(w, 8, 16, 24), x)
    pix = O.window_token_pixels(16, 24, 8, 4)
    rolled = torch.roll(x, shifts=(-4, -4), dims=(1, 2))
    assert torch.equal(O.window_partition(rolled, 8).reshape(2, -1, 3), x.reshape(2, -1, 3)[:, pix.reshape(-1)])


def test_pixel_shuffle_and_tail(golden):
    g = golden("kat_tail")
    ps_in = torch.from_numpy(np.random.default_rng(33).normal(0, 1, size=(2, 16, 5, 7)).astype(np.float32))
    assert np.array_equal(O.pixel_shuffle(ps_in, 2).numpy(), g["ps"])
    cfg = synth.CONFIGS["swinir_x4_d2"]
    sd = synth.make_swinir_state_dict(cfg, seed=31, kind="stress")
    feat = torch.from_numpy(np.random.default_rng(32).normal(0, 1, size=(1, 180, 12, 10)).astype(np.float32))
    _close(O.upsample_tail(feat, sd, cfg), g["y"], 1e-4)


def test_window_attention_kat(golden):
    g = golden("kat_window_attention")
    cfg = synth.CONFIGS["swinir_x2_d2"]
    sd = synth.make_swinir_state_dict(cfg, seed=99, kind="stress")
    pre = "layers.0.residual_group.blocks.1.attn."
    xw = synth.make_tokens(8, 8, 8, 180, seed=5)
    _close(O.window_attention(xw, sd, pre, 6, 8, None), g["y_nomask"])
    _close(O.window_attention(xw, sd, pre, 6, 8, torch.from_numpy(g["mask"])), g["y_mask"])
    # stress weights really are peaky: a uniform softmax would not notice a wrong RPB index
    q = (xw @ sd[pre + "qkv.weight"].T)[..., :180].reshape(8, 64, 6, 30) * 30 ** -0.5
    k = (xw @ sd[pre + "qkv.weight"].T)[..., 180:360].reshape(8, 64, 6, 30)
    assert torch.einsum("bnhd,bmhd->bhnm", q, k).std() > 1.0


@pytest.mark.parametrize("tag,b_idx,shift,x_size", [
    ("unshifted", 0, 0, (16, 24)), ("shifted", 1, 4, (16, 24)),
    ("shifted_nonnative", 1, 4, (24, 16)), ("shifted_64", 1, 4, (64, 64))])
def test_block_kat(golden, tag, b_idx, shift, x_size):
    g = golden(f"kat_block_{tag}")
    cfg = synth.CONFIGS["swinir_x2_d2"]
    sd = synth.make_swinir_state_dict(cfg, seed=99, kind="stress")
    B = 2 if x_size != (64, 64) else 1
    xt = synth.make_tokens(B, x_size[0], x_size[1], 180, seed=11)
    y = O.swin_block(xt, x_size, sd, f"layers.0.residual_group.blocks.{b_idx}.", 6, 8, shift)
    if x_size == (64, 64):
        y = y[:, ::7]
    _close(y, g["y"])


def test_rstb_kat(golden):
    g = golden("kat_rstb")
    cfg = synth.CONFIGS["swinir_x2_d2"]
    sd = synth.make_swinir_state_dict(cfg, seed=99, kind="stress")
    xt = synth.make_tokens(1, 16, 16, 180, seed=12)
    _close(O.rstb(xt, (16, 16), sd, "layers.1.", 2, 6, 8), g["y"])


@pytest.mark.parametrize("name,kind,seed,B,h,w", [
    ("swinir_x2", "init", 1234, 1, 64, 64),
    ("swinir_x2", "stress", 4321, 1, 64, 64),
    ("swinir_x4_d2", "stress", 4321, 1, 32, 40),
    ("swinir_x4_d2", "init", 1234, 2, 64, 64),
    ("swinir_x2_d2", "stress", 77, 1, 20, 27)])
def test_whole_model_golden(golden, name, kind, seed, B, h, w):
    g = golden(f"{name}_{kind}_{B}x{h}x{w}")
    cfg = synth.CONFIGS[name]
    sd = synth.make_swinir_state_dict(cfg, seed=seed, kind=kind)
    y = O.swinir_forward(synth.make_lr_batch(B, h, w, seed=seed + 1), sd, cfg)
    assert y.shape == g["y"].shape
    _close(y, g["y"], 5e-5)


def test_fp64_mode_agrees():
    cfg = synth.CONFIGS["swinir_x2_d2"]
    sd = synth.make_swinir_state_dict(cfg, seed=3, kind="stress")
    lr = synth.make_lr_batch(1, 16, 16, seed=4)
    y32 = O.swinir_forward(lr, sd, cfg)
    y64 = O.swinir_forward(lr.double(), O.to_dtype(sd, torch.float64), cfg)
    assert (y32.double() - y64).abs().max() < 2e-5


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_against_live_reference_with_its_own_init():
    """SURVEY A.6 anchors: reference init under torch.manual_seed(1234); oracle on the same state_dict."""
    with open(os.path.join(GOLDEN, "swinir_x2_anchors.json")) as f:
        anchors = json.load(f)
    ns = load_reference_module("network_swinir")
    torch.manual_seed(1234)
    cfg = synth.CONFIGS["swinir_x2"]
    model = ns.SwinIR(**cfg.as_kwargs()).eval()
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        y_ref = model(x)
    assert abs(float(y_ref.double().sum()) - 20705.493491) < 0.05
    assert abs(float(y_ref.double().sum()) - anchors["sum_f64"]) < 0.05
    y = O.swinir_forward(x, dict(model.state_dict()), cfg)
    assert (y - y_ref).abs().max() < 2e-5
