"""Generate tests/golden/dat_* and kat_dat_* from the UNMODIFIED reference dat_arch.py (build container only).

    python -m oracle.make_golden_dat
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import synth  # noqa: E402
from oracle.make_golden import GOLDEN, _block_state, _save  # noqa: E402
from oracle.reference_loader import load_reference_module  # noqa: E402


@torch.no_grad()
def main() -> None:
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    dat = load_reference_module("dat_arch")

    cfg = synth.DAT_CONFIGS["dat_x2"]
    model = dat.DAT(**cfg.as_kwargs()).eval()
    man = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()]
    with open(os.path.join(GOLDEN, "dat_x2_manifest.json"), "w") as f:
        json.dump(man, f, indent=0)
    print(f"dat_x2: {len(man)} state_dict entries, {sum(p.numel() for p in model.parameters())} params")

    # whole-model outputs (64x64 native; 32x96: x_size != img_size -> masks rebuilt per forward, dat_arch.py:396-399)
    # 40x72: neither side a multiple of 32 -> projected q, k, v zero-padded to 64x96, masks of the padded size, crop (:376-407)
    for name, kind, seed, B, h, w in [("dat_x2_d3", "init", 1234, 1, 64, 64), ("dat_x2_d3", "stress", 4321, 1, 32, 96),
                                      ("dat_x2_d3", "stress", 77, 1, 40, 72)]:
        cfg = synth.DAT_CONFIGS[name]
        model = dat.DAT(**cfg.as_kwargs()).eval()
        model.load_state_dict(synth.make_dat_state_dict(cfg, seed=seed, kind=kind), strict=True)
        lr = synth.make_lr_batch(B, h, w, seed=seed + 1)
        y = model(lr)
        assert torch.isfinite(y).all()
        _save(f"{name}_{kind}_{B}x{h}x{w}", y=y, seed=seed, lr_seed=seed + 1)

    # module KATs, stress weights
    cfg = synth.DAT_CONFIGS["dat_x2_d3"]
    sd = synth.make_dat_state_dict(cfg, seed=99, kind="stress")
    model = dat.DAT(**cfg.as_kwargs()).eval()
    model.load_state_dict(sd, strict=True)
    H, W = 64, 64
    xt = synth.make_tokens(1, H, W, 180, seed=11)
    blocks = model.layers[0].blocks
    _save("kat_dat_spatial", y_unshifted=blocks[0].attn(xt, H, W)[:, ::11], y_shifted=blocks[2].attn(xt, H, W)[:, ::11])
    xp = synth.make_tokens(2, 40, 72, 180, seed=12)
    _save("kat_dat_spatial_padded", y_unshifted=blocks[0].attn(xp, 40, 72)[:, ::7], y_shifted=blocks[2].attn(xp, 40, 72)[:, ::7])
    _save("kat_dat_channel", y=blocks[1].attn(xt, H, W)[:, ::11])
    _save("kat_dat_sgfn", y=blocks[0].ffn(xt, H, W)[:, ::11])
    _save("kat_dat_block", y0=blocks[0](xt, (H, W))[:, ::11], y1=blocks[1](xt, (H, W))[:, ::11], y2=blocks[2](xt, (H, W))[:, ::11])
    _save("kat_dat_rg", y=model.layers[1](xt, (H, W))[:, ::11])
    a = blocks[2].attn
    _save("kat_dat_buffers", mask0=a.attn_mask_0.to(torch.int8)[::3], mask1=a.attn_mask_1.to(torch.int8)[::3],
          rpi0=a.attns[0].relative_position_index.to(torch.int16), rpi1=a.attns[1].relative_position_index.to(torch.int16),
          pos0=a.attns[0].pos(a.attns[0].rpe_biases), pos1=a.attns[1].pos(a.attns[1].rpe_biases))


if __name__ == "__main__":
    main()
