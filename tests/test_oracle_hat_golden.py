"""The HAT oracle restatement (oracle/hat_oracle.py) against fixtures made from the unmodified reference
(oracle/make_golden_hat.py), plus the closed forms against the reference's own buffers."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import hat_oracle as HO
from oracle import swinir_oracle as O
from oracle import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-5


def _g(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def _err(a, b):
    return float((a.double() - torch.from_numpy(np.asarray(b)).double()).abs().max())


@pytest.fixture(scope="module")
def stress():
    cfg = synth.HAT_CONFIGS["hat_x4_d2"]
    return cfg, synth.make_hat_state_dict(cfg, seed=99, kind="stress")


def test_manifest_matches_reference():
    man = json.load(open(os.path.join(GOLDEN, "hat_x4_manifest.json")))
    ours = synth.hat_manifest(synth.HAT_CONFIGS["hat_x4"])
    assert [m[0] for m in man] == [k for k, _, _ in ours]
    assert [tuple(m[1]) for m in man] == [tuple(s) for _, s, _ in ours]


def test_closed_forms_match_reference_buffers():
    g = _g("kat_hat_buffers")
    assert np.array_equal(HO.rpi_sa(16).numpy(), g["rpi_sa"].astype(np.int64))
    assert np.array_equal(HO.rpi_oca(16, 0.5).numpy(), g["rpi_oca"].astype(np.int64))
    assert int(g["rpi_oca"].min()) == -880 and int(g["rpi_oca"].max()) == 640
    assert np.array_equal(O.shift_attention_mask(32, 48, 16, 8).numpy().astype(np.int8), g["mask_32x48"])


def test_window_attention(stress):
    cfg, sd = stress
    g = _g("kat_hat_window_attention")
    xw = synth.make_tokens(4, 16, 16, 180, seed=5)
    pre = "layers.0.residual_group.blocks.1.attn."
    assert _err(HO.hat_window_attention(xw, sd, pre, 6, 16, None)[:, ::3], g["y_nomask"]) < TOL
    mask = torch.from_numpy(g["mask"].astype(np.float32))
    assert _err(HO.hat_window_attention(xw, sd, pre, 6, 16, mask)[:, ::3], g["y_mask"]) < TOL


def test_hab_ocab_rhag_cab(stress):
    cfg, sd = stress
    x_size = (32, 48)
    xt = synth.make_tokens(2, 32, 48, 180, seed=11)
    g = _g("kat_hat_hab")
    y0 = HO.hab(xt, x_size, sd, "layers.0.residual_group.blocks.0.", 6, 16, 0, cfg.conv_scale)
    y1 = HO.hab(xt, x_size, sd, "layers.0.residual_group.blocks.1.", 6, 16, 8, cfg.conv_scale)
    assert _err(y0[:, ::5], g["y_unshifted"]) < TOL and _err(y1[:, ::5], g["y_shifted"]) < TOL
    yo = HO.ocab(xt, x_size, sd, "layers.0.residual_group.overlap_attn.", 6, 16, cfg.overlap_ratio)
    assert _err(yo[:, ::5], _g("kat_hat_ocab")["y"]) < TOL
    yr = HO.rhag(xt[:1], x_size, sd, "layers.1.", 2, 6, cfg)
    assert _err(yr[:, ::3], _g("kat_hat_rhag")["y"]) < 5e-5
    ximg = torch.from_numpy(np.random.default_rng(13).normal(0, 1, size=(2, 180, 12, 20)).astype(np.float32))
    assert _err(HO.cab(ximg, sd, "layers.0.residual_group.blocks.0.conv_block."), _g("kat_hat_cab")["y"]) < TOL


@pytest.mark.parametrize("name,kind,seed,B,h,w", [("hat_x4_d2", "init", 1234, 1, 64, 64),
                                                   ("hat_x4_d2", "stress", 4321, 1, 32, 48),
                                                   ("hat_x2_d2", "stress", 77, 1, 20, 27)])
def test_whole_model(name, kind, seed, B, h, w):
    cfg = synth.HAT_CONFIGS[name]
    sd = synth.make_hat_state_dict(cfg, seed=seed, kind=kind)
    lr = synth.make_lr_batch(B, h, w, seed=seed + 1)
    y = HO.hat_forward(lr, sd, cfg)
    assert _err(y, _g(f"{name}_{kind}_{B}x{h}x{w}")["y"]) < 5e-5
