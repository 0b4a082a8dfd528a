"""Generate tests/golden/* from the UNMODIFIED reference (run in the build container only).

    python -m oracle.make_golden

Every fixture is an output of ``/root/reference/modules/network_swinir.py`` classes,
built with the reference's own constructors, loaded ``strict=True`` with the synthetic
state_dict of ``oracle/synth.py`` and run in fp32 on the CPU.  The oracle restatement
and the CUDA path are both tested against these files; nothing here runs on the GPU box.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import synth  # noqa: E402
from oracle.reference_loader import load_reference_module  # noqa: E402
from oracle.swinir_oracle import SwinIRConfig  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _save(name: str, **arrays) -> None:
    path = os.path.join(GOLDEN, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print(f"wrote {path}.npz  ({os.path.getsize(path + '.npz') / 1e3:.0f} KB)")


def _block_state(sd, pre):
    return {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}


@torch.no_grad()
def main() -> None:
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    ns = load_reference_module("network_swinir")

    # ---- 1. state_dict manifests of the real reference (keys / shapes / dtypes) -------------
    for name in ("swinir_x2", "swinir_x4"):
        cfg = synth.CONFIGS[name]
        model = ns.SwinIR(**cfg.as_kwargs()).eval()
        man = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()]
        with open(os.path.join(GOLDEN, f"{name}_manifest.json"), "w") as f:
            json.dump(man, f, indent=0)
        print(f"{name}: {len(man)} state_dict entries, {sum(p.numel() for p in model.parameters())} params")

    # ---- 2. SURVEY §A.6 anchors: reference init under torch.manual_seed(1234) ----------------
    torch.manual_seed(1234)
    cfg = synth.CONFIGS["swinir_x2"]
    model = ns.SwinIR(**cfg.as_kwargs()).eval()
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    y = model(x)
    anchors = {
        "x_first4": x[0, 0, 0, :4].tolist(),
        "qkv_w_first3": model.layers[0].residual_group.blocks[0].attn.qkv.weight[0, :3].tolist(),
        "sum_f64": float(y.double().sum()), "mean": float(y.mean()), "std": float(y.std()),
        "min": float(y.min()), "max": float(y.max()),
        "y_000": y[0, :, 0, 0].tolist(), "y_1_64_64": float(y[0, 1, 64, 64]),
        "torch": torch.__version__,
    }
    with open(os.path.join(GOLDEN, "swinir_x2_anchors.json"), "w") as f:
        json.dump(anchors, f, indent=1)
    print("anchors:", anchors)

    # ---- 3. whole-model outputs on synthetic (numpy-RNG) weights -----------------------------
    for name, kind, seed, B, h, w in [
        ("swinir_x2", "init", 1234, 1, 64, 64),          # BASELINE configs[0]
        ("swinir_x2", "stress", 4321, 1, 64, 64),
        ("swinir_x4_d2", "stress", 4321, 1, 32, 40),     # non-native x_size -> mask rebuilt (:259-262)
        ("swinir_x4_d2", "init", 1234, 2, 64, 64),
        ("swinir_x2_d2", "stress", 77, 1, 20, 27),       # not a multiple of ws -> reflect pad + crop (:783-788, :840)
    ]:
        cfg = synth.CONFIGS[name]
        model = ns.SwinIR(**cfg.as_kwargs()).eval()
        sd = synth.make_swinir_state_dict(cfg, seed=seed, kind=kind)
        model.load_state_dict(sd, strict=True)
        lr = synth.make_lr_batch(B, h, w, seed=seed + 1)
        y = model(lr)
        assert torch.isfinite(y).all()
        _save(f"{name}_{kind}_{B}x{h}x{w}", y=y, seed=seed, lr_seed=seed + 1)

    # ---- 4. module-level KATs (stress weights) ------------------------------------------------
    cfg = synth.CONFIGS["swinir_x2_d2"]
    sd = synth.make_swinir_state_dict(cfg, seed=99, kind="stress")
    C, nh, ws = cfg.embed_dim, 6, cfg.window_size

    # 4a. WindowAttention with and without an explicit mask (network_swinir.py:114-145)
    attn = ns.WindowAttention(C, (ws, ws), nh).eval()
    pre = "layers.0.residual_group.blocks.1.attn."
    attn.load_state_dict(_block_state(sd, pre), strict=True)
    xw = synth.make_tokens(8, ws, ws, C, seed=5)
    rng = np.random.default_rng(6)
    mask = torch.from_numpy(np.where(rng.random((4, ws * ws, ws * ws)) < 0.3, -100.0, 0.0).astype(np.float32))
    _save("kat_window_attention", y_nomask=attn(xw), y_mask=attn(xw, mask), mask=mask)

    # 4b. SwinTransformerBlock shifted / un-shifted, native and non-native x_size (:239-279)
    for tag, b_idx, shift, res, x_size in [
        ("unshifted", 0, 0, (16, 24), (16, 24)),
        ("shifted", 1, ws // 2, (16, 24), (16, 24)),
        ("shifted_nonnative", 1, ws // 2, (16, 24), (24, 16)),
        ("shifted_64", 1, ws // 2, (64, 64), (64, 64)),
    ]:
        blk = ns.SwinTransformerBlock(C, res, nh, window_size=ws, shift_size=shift, mlp_ratio=cfg.mlp_ratio).eval()
        pre = f"layers.0.residual_group.blocks.{b_idx}."
        st = _block_state(sd, pre)
        if shift > 0:
            st["attn_mask"] = blk.attn_mask.clone()     # buffer depends on input_resolution, not a weight
        blk.load_state_dict(st, strict=True)
        B = 2 if x_size != (64, 64) else 1
        xt = synth.make_tokens(B, x_size[0], x_size[1], C, seed=11)
        y = blk(xt, x_size)
        if x_size == (64, 64):
            y = y[:, ::7]          # keep the fixture small: every 7th token
        _save(f"kat_block_{tag}", y=y)

    # 4c. RSTB depth 2 (:481-482)
    r = ns.RSTB(C, (16, 16), 2, nh, ws, mlp_ratio=cfg.mlp_ratio, img_size=16, patch_size=1).eval()
    st = _block_state(sd, "layers.1.")
    st["residual_group.blocks.1.attn_mask"] = r.residual_group.blocks[1].attn_mask.clone()
    r.load_state_dict(st, strict=True)
    xt = synth.make_tokens(1, 16, 16, C, seed=12)
    _save("kat_rstb", y=r(xt, (16, 16)))

    # 4d. reference buffers for the closed forms
    blk = ns.SwinTransformerBlock(C, (64, 64), nh, window_size=ws, shift_size=ws // 2, mlp_ratio=2.0)
    blk2 = ns.SwinTransformerBlock(C, (48, 40), nh, window_size=ws, shift_size=ws // 2, mlp_ratio=2.0)
    _save("kat_buffers", rpi=blk.attn.relative_position_index, mask_64=blk.attn_mask.to(torch.int8),
          mask_48x40=blk2.attn_mask.to(torch.int8))

    # 4e. upsample tail / PixelShuffle (:572-591, :742-745)
    cfg4 = synth.CONFIGS["swinir_x4_d2"]
    m4 = ns.SwinIR(**cfg4.as_kwargs()).eval()
    sd4 = synth.make_swinir_state_dict(cfg4, seed=31, kind="stress")
    m4.load_state_dict(sd4, strict=True)
    feat = torch.from_numpy(np.random.default_rng(32).normal(0, 1, size=(1, C, 12, 10)).astype(np.float32))
    t = m4.conv_last(m4.upsample(m4.conv_before_upsample(feat)))
    ps_in = torch.from_numpy(np.random.default_rng(33).normal(0, 1, size=(2, 16, 5, 7)).astype(np.float32))
    _save("kat_tail", y=t, ps=torch.nn.PixelShuffle(2)(ps_in))


if __name__ == "__main__":
    main()
