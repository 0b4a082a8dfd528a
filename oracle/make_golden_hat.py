"""Generate tests/golden/hat_* and kat_hat_* from the UNMODIFIED reference hat_arch.py (build container only).

    python -m oracle.make_golden_hat

Same rules as oracle/make_golden.py: the reference's own constructors, synthetic numpy-RNG state_dicts loaded
strict=True, fp32 on the CPU.
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
    hat = load_reference_module("hat_arch")

    # ---- 1. manifest of the real reference (BASELINE configs[2]) ---------------------------------
    cfg = synth.HAT_CONFIGS["hat_x4"]
    model = hat.HAT(**cfg.as_kwargs()).eval()
    man = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()]
    with open(os.path.join(GOLDEN, "hat_x4_manifest.json"), "w") as f:
        json.dump(man, f, indent=0)
    print(f"hat_x4: {len(man)} state_dict entries, {sum(p.numel() for p in model.parameters())} params")

    # ---- 2. whole-model outputs ------------------------------------------------------------------
    for name, kind, seed, B, h, w in [
        ("hat_x4_d2", "init", 1234, 1, 64, 64),
        ("hat_x4_d2", "stress", 4321, 1, 32, 48),      # x_size != img_size: mask rebuilt per forward (hat_arch.py:955)
        ("hat_x2_d2", "stress", 77, 1, 20, 27),        # reflect pad to 32x32 + crop (hat_arch.py:971-976, :994)
    ]:
        cfg = synth.HAT_CONFIGS[name]
        model = hat.HAT(**cfg.as_kwargs()).eval()
        sd = synth.make_hat_state_dict(cfg, seed=seed, kind=kind)
        model.load_state_dict(sd, strict=True)
        lr = synth.make_lr_batch(B, h, w, seed=seed + 1)
        y = model(lr)
        assert torch.isfinite(y).all()
        _save(f"{name}_{kind}_{B}x{h}x{w}", y=y, seed=seed, lr_seed=seed + 1)

    # ---- 3. module KATs with stress weights -------------------------------------------------------
    cfg = synth.HAT_CONFIGS["hat_x4_d2"]
    sd = synth.make_hat_state_dict(cfg, seed=99, kind="stress")
    C, nh, ws = cfg.embed_dim, 6, cfg.window_size
    model = hat.HAT(**cfg.as_kwargs()).eval()
    model.load_state_dict(sd, strict=True)
    rpi_sa_buf, rpi_oca_buf = model.relative_position_index_SA, model.relative_position_index_OCA

    # 3a. WindowAttention.forward(x, rpi, mask) (hat_arch.py:166-197)
    attn = model.layers[0].residual_group.blocks[1].attn
    xw = synth.make_tokens(4, ws, ws, C, seed=5)
    rng = np.random.default_rng(6)
    mask = torch.from_numpy(np.where(rng.random((2, ws * ws, ws * ws)) < 0.3, -100.0, 0.0).astype(np.float32))
    _save("kat_hat_window_attention", y_nomask=attn(xw, rpi_sa_buf)[:, ::3], y_mask=attn(xw, rpi_sa_buf, mask)[:, ::3], mask=mask.to(torch.int8))

    # (fixtures keep every 3rd / 5th token to stay small)
    # 3b. HAB un-shifted / shifted (hat_arch.py:267-310) at 32x48; 3c. OCAB (:393-439); 3d. RHAG (:619-620)
    x_size = (32, 48)
    xt = synth.make_tokens(2, x_size[0], x_size[1], C, seed=11)
    amask = model.calculate_mask(x_size)
    params = {"attn_mask": amask, "rpi_sa": rpi_sa_buf, "rpi_oca": rpi_oca_buf}
    blocks = model.layers[0].residual_group.blocks
    _save("kat_hat_hab", y_unshifted=blocks[0](xt, x_size, rpi_sa_buf, amask)[:, ::5], y_shifted=blocks[1](xt, x_size, rpi_sa_buf, amask)[:, ::5])
    _save("kat_hat_ocab", y=model.layers[0].residual_group.overlap_attn(xt, x_size, rpi_oca_buf)[:, ::5])
    _save("kat_hat_rhag", y=model.layers[1](xt[:1], x_size, params)[:, ::3])
    # 3e. CAB alone (hat_arch.py:62-75)
    ximg = torch.from_numpy(np.random.default_rng(13).normal(0, 1, size=(2, C, 12, 20)).astype(np.float32))
    _save("kat_hat_cab", y=blocks[0].conv_block(ximg))
    # 3f. buffers for the closed forms
    _save("kat_hat_buffers", rpi_sa=rpi_sa_buf.to(torch.int16), rpi_oca=rpi_oca_buf.to(torch.int16),
          mask_32x48=amask.to(torch.int8))


if __name__ == "__main__":
    main()
