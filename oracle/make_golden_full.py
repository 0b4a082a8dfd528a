"""Generate the FULL-DEPTH whole-model fixtures from the UNMODIFIED reference (build container only).

    python -m oracle.make_golden_full

BASELINE.json configs[1], [2], [3] at their real depth (6 groups x 6 blocks; HAT 6 x (6 HAB + OCAB)), one 64x64 LR tile
each, with the init-scale weights the benchmark uses (seed 1234) and with stress weights (seed 4321: peaky softmax, bias
tables of std 1, non-trivial BatchNorm statistics).  Same rules as oracle/make_golden.py: the reference's own
constructors (network_swinir.py:618, hat_arch.py:738-764, dat_arch.py:717-737), synthetic numpy-RNG state_dicts loaded
strict=True, fp32 on the CPU.  Outputs are stored as float32 (npz, compressed).
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import synth  # noqa: E402
from oracle.make_golden import GOLDEN, _save  # noqa: E402
from oracle.reference_loader import load_reference_module  # noqa: E402

CASES = [
    # (reference module, class, config table, config name, state_dict maker)
    ("network_swinir", "SwinIR", "CONFIGS", "swinir_x4", "make_swinir_state_dict"),
    ("hat_arch", "HAT", "HAT_CONFIGS", "hat_x4", "make_hat_state_dict"),
    ("dat_arch", "DAT", "DAT_CONFIGS", "dat_x2", "make_dat_state_dict"),
]


@torch.no_grad()
def main() -> None:
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    for modname, clsname, table, name, maker in CASES:
        mod = load_reference_module(modname)
        cfg = getattr(synth, table)[name]
        for kind, seed in (("init", 1234), ("stress", 4321)):
            model = getattr(mod, clsname)(**cfg.as_kwargs()).eval()
            sd = getattr(synth, maker)(cfg, seed=seed, kind=kind)
            model.load_state_dict(sd, strict=True)
            lr = synth.make_lr_batch(1, 64, 64, seed=seed + 1)
            y = model(lr)
            assert torch.isfinite(y).all()
            print(f"{name} {kind}: out range [{y.min().item():.4f}, {y.max().item():.4f}] std {y.std().item():.4f}")
            _save(f"{name}_{kind}_1x64x64", y=y, seed=seed, lr_seed=seed + 1)


if __name__ == "__main__":
    main()
