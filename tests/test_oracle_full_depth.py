"""The oracle restatements at the FULL depth of BASELINE.json configs[1], [2], [3] against outputs of the unmodified
reference (oracle/make_golden_full.py): 6 groups x 6 blocks (HAT: + OCAB per group), init-scale and stress weights."""
import os

import numpy as np
import pytest
import torch

from oracle import dat_oracle as DO
from oracle import hat_oracle as HO
from oracle import swinir_oracle as O
from oracle import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

FAMILIES = {
    "swinir_x4": (synth.CONFIGS, synth.make_swinir_state_dict, O.swinir_forward),
    "hat_x4": (synth.HAT_CONFIGS, synth.make_hat_state_dict, HO.hat_forward),
    "dat_x2": (synth.DAT_CONFIGS, synth.make_dat_state_dict, DO.dat_forward),
}


@pytest.mark.parametrize("name", sorted(FAMILIES))
@pytest.mark.parametrize("kind,seed", [("init", 1234), ("stress", 4321)])
def test_full_depth_oracle_matches_reference(name, kind, seed):
    table, make_sd, fwd = FAMILIES[name]
    cfg = table[name]
    sd = make_sd(cfg, seed=seed, kind=kind)
    lr = synth.make_lr_batch(1, 64, 64, seed=seed + 1)
    with torch.no_grad():
        y = fwd(lr, sd, cfg)
    ref = torch.from_numpy(np.load(os.path.join(GOLDEN, f"{name}_{kind}_1x64x64.npz"))["y"])
    assert y.shape == ref.shape
    # fp32 restatement vs fp32 reference: summation-order noise only (measured <= 3e-5 over 36 blocks)
    assert (y.double() - ref.double()).abs().max().item() < 1e-4
