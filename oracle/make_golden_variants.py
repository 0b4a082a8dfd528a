"""Generate tests/golden/swinir_variant_*.npz from the UNMODIFIED reference (build container only): the constructor variants of
network_swinir.py beyond the one configuration the repo uses -- upsampler 'pixelshuffledirect' (:746-749, :818-822), 'nearest+conv'
(:750-759, :823-831), '' (denoising tail, :760-762, :832-836), resi_connection '3conv' (:466-471, :731-738), ape=True (:694-696,
:793-794) and the x3 pixelshuffle stage (:586-588).  Weights: synth.generic_state_dict (numpy RNG keyed by parameter name).

    python -m oracle.make_golden_variants
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.make_golden import GOLDEN, _save  # noqa: E402
from oracle.reference_loader import load_reference_module  # noqa: E402
from tpu_superresolution_b200 import synth  # noqa: E402


@torch.no_grad()
def main() -> None:
    ns = load_reference_module("network_swinir")
    torch.set_num_threads(os.cpu_count() or 1)
    for name, kw in synth.SWINIR_VARIANTS.items():
        model = ns.SwinIR(**kw).eval()
        sd = synth.generic_state_dict(model.state_dict(), seed=7)
        model.load_state_dict(sd, strict=True)
        lr = synth.make_lr_batch(2, 16, 16, seed=21)
        y = model(lr)
        assert torch.isfinite(y).all()
        print(f"{name}: out {tuple(y.shape)} range [{y.min().item():.3f}, {y.max().item():.3f}]")
        _save(f"swinir_variant_{name}", y=y, keys="\n".join(sd.keys()))


if __name__ == "__main__":
    main()
