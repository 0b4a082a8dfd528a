"""The DAT oracle restatement (oracle/dat_oracle.py) against fixtures made from the unmodified reference
(oracle/make_golden_dat.py), plus the closed forms against the reference's own buffers."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import dat_oracle as DO
from oracle import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 3e-5


def _g(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def _err(a, b):
    return float((a.double() - torch.from_numpy(np.asarray(b)).double()).abs().max())


@pytest.fixture(scope="module")
def stress():
    cfg = synth.DAT_CONFIGS["dat_x2_d3"]
    return cfg, synth.make_dat_state_dict(cfg, seed=99, kind="stress")


def test_manifest_matches_reference():
    man = json.load(open(os.path.join(GOLDEN, "dat_x2_manifest.json")))
    ours = synth.dat_manifest(synth.DAT_CONFIGS["dat_x2"])
    assert [m[0] for m in man] == [k for k, _, _ in ours]
    assert [tuple(m[1]) for m in man] == [tuple(s) for _, s, _ in ours]


def test_closed_forms_match_reference_buffers(stress):
    cfg, sd = stress
    g = _g("kat_dat_buffers")
    assert np.array_equal(DO.rect_relative_position_index(8, 32).numpy(), g["rpi0"].astype(np.int64))
    assert np.array_equal(DO.rect_relative_position_index(32, 8).numpy(), g["rpi1"].astype(np.int64))
    assert np.array_equal(DO.rect_shift_mask(64, 64, 8, 32, 4, 16).numpy().astype(np.int8)[::3], g["mask0"])
    assert np.array_equal(DO.rect_shift_mask(64, 64, 32, 8, 16, 4).numpy().astype(np.int8)[::3], g["mask1"])
    pre = "layers.0.blocks.2.attn.attns."
    assert _err(DO.dynamic_pos_bias_table(sd, pre + "0.pos.", 8, 32), g["pos0"]) < 1e-5
    assert _err(DO.dynamic_pos_bias_table(sd, pre + "1.pos.", 32, 8), g["pos1"]) < 1e-5
    assert DO.is_shifted(0, 2) and DO.is_shifted(1, 0) and DO.is_shifted(1, 4) and not DO.is_shifted(0, 0) and not DO.is_shifted(1, 2)


def test_modules(stress):
    cfg, sd = stress
    H = W = 64
    xt = synth.make_tokens(1, H, W, 180, seed=11)
    g = _g("kat_dat_spatial")
    y0 = DO.adaptive_spatial_attention(xt, H, W, sd, "layers.0.blocks.0.attn.", 6, cfg.split_size, False)
    y2 = DO.adaptive_spatial_attention(xt, H, W, sd, "layers.0.blocks.2.attn.", 6, cfg.split_size, True)
    assert _err(y0[:, ::11], g["y_unshifted"]) < TOL and _err(y2[:, ::11], g["y_shifted"]) < TOL
    assert _err(DO.adaptive_channel_attention(xt, H, W, sd, "layers.0.blocks.1.attn.", 6)[:, ::11], _g("kat_dat_channel")["y"]) < TOL
    assert _err(DO.sgfn(xt, H, W, sd, "layers.0.blocks.0.ffn.")[:, ::11], _g("kat_dat_sgfn")["y"]) < TOL
    gb = _g("kat_dat_block")
    for b in range(3):
        assert _err(DO.datb(xt, (H, W), sd, f"layers.0.blocks.{b}.", 6, cfg.split_size, 0, b)[:, ::11], gb[f"y{b}"]) < 1e-4
    assert _err(DO.residual_group(xt, (H, W), sd, "layers.1.", 2, 6, cfg.split_size, 1)[:, ::11], _g("kat_dat_rg")["y"]) < 1e-4


def test_spatial_attention_padded(stress):
    """dat_arch.py:376-407: H, W not multiples of 32 -> projected q, k, v zero-padded, masks of the padded size, crop."""
    cfg, sd = stress
    xp = synth.make_tokens(2, 40, 72, 180, seed=12)
    g = _g("kat_dat_spatial_padded")
    y0 = DO.adaptive_spatial_attention(xp, 40, 72, sd, "layers.0.blocks.0.attn.", 6, cfg.split_size, False)
    y2 = DO.adaptive_spatial_attention(xp, 40, 72, sd, "layers.0.blocks.2.attn.", 6, cfg.split_size, True)
    assert _err(y0[:, ::7], g["y_unshifted"]) < TOL and _err(y2[:, ::7], g["y_shifted"]) < TOL


@pytest.mark.parametrize("name,kind,seed,B,h,w", [("dat_x2_d3", "init", 1234, 1, 64, 64), ("dat_x2_d3", "stress", 4321, 1, 32, 96),
                                                   ("dat_x2_d3", "stress", 77, 1, 40, 72)])
def test_whole_model(name, kind, seed, B, h, w):
    cfg = synth.DAT_CONFIGS[name]
    sd = synth.make_dat_state_dict(cfg, seed=seed, kind=kind)
    y = DO.dat_forward(synth.make_lr_batch(B, h, w, seed=seed + 1), sd, cfg)
    assert _err(y, _g(f"{name}_{kind}_{B}x{h}x{w}")["y"]) < 5e-5
