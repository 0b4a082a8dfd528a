#!/usr/bin/env python
"""Checkpoint + evaluation loop of the reference on the fused path.

Mirrors finetune_swinir.py: the checkpoint is loaded like :283-285 (`ckpt["params"]` if present, `strict=True`), the images are
paired like sr_datasets.Shuffled2DPaired (:31-74: LR `<stem>[_-]x<s>` <-> HR `<stem>`), and `validate()` (:182-207) is run with the
same metrics -- L1 loss, `batch_psnr` (:69-74: clamped to [0, 1], +1e-8) -- plus SSIM (11x11 Gaussian, sigma 1.5, the
pytorch_msssim definition the repo's environment pins, restated here because the package is absent).  The model is the drop-in
`tpu_superresolution_b200` class; whole images run through the model (it reflect-pads to the window size itself) or, with `--tile`,
through `TiledSuperResolver` (overlapping tiles, one CUDA-graph replay per batch).

    python tools/validate.py --arch swinir --scale 4 --weights 001_classicalSR_DF2K_s64w8_SwinIR-M_x4.pth \
        --lr-dir data/Set5/LRbicx4 --hr-dir data/Set5/GTmod12 [--tile 64 --overlap 8] [--save-dir out/]
    python tools/validate.py --selftest          # synthetic checkpoint + images in a temp dir (what tests/ runs)
"""
from __future__ import annotations

import argparse
import os
import re
import sys
import tempfile
import time
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

EXTS = (".png", ".jpg", ".jpeg", ".tif", ".tiff")


def load_checkpoint(model: torch.nn.Module, path: str) -> None:
    """finetune_swinir.py:283-285: `params` sub-dict of BasicSR checkpoints (also `params_ema`), strict key match."""
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    state = ckpt
    for k in ("params_ema", "params"):
        if isinstance(ckpt, dict) and k in ckpt:
            state = ckpt[k]
            break
    missing, unexpected = model.load_state_dict(state, strict=True)
    print(f"[weights] loaded {path}: missing={len(missing)} unexpected={len(unexpected)}")


def pair_files(lr_dir: str, hr_dir: str, scale: int):
    """sr_datasets.py:45-60: HR by stem; LR stem with a trailing `_x<s>` / `-x<s>` / `x<s>` removed."""
    hr = {p.stem: p for p in sorted(Path(hr_dir).iterdir()) if p.suffix.lower() in EXTS}
    pairs = []
    for p in sorted(Path(lr_dir).iterdir()):
        if p.suffix.lower() not in EXTS:
            continue
        stem = re.sub(rf"([_-]?)x{scale}$", "", p.stem, flags=re.IGNORECASE)
        if stem in hr:
            pairs.append((p, hr[stem]))
    if not pairs:
        raise RuntimeError(f"no LR/HR pairs between {lr_dir} and {hr_dir}")
    return pairs


def to_tensor(path: Path) -> torch.Tensor:
    """(1, 3, H, W) float32 in [0, 1]; grey images are repeated to 3 channels (finetune_swinir.py:80-90 `_ensure_3ch`)."""
    from PIL import Image
    with Image.open(path) as im:
        a = np.asarray(im.convert("RGB"), dtype=np.float32) / 255.0
    return torch.from_numpy(a).permute(2, 0, 1).unsqueeze(0).contiguous()


def batch_psnr(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    pred, target = pred.clamp(0, 1), target.clamp(0, 1)
    mse = F.mse_loss(pred, target, reduction="none").flatten(1).mean(1)
    return 20.0 * torch.log10(1.0 / torch.sqrt(mse + 1e-8))


def ssim(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """SSIM of pytorch_msssim.ssim(data_range=1, win 11, sigma 1.5, valid convolution), mean over channels and pixels."""
    pred, target = pred.clamp(0, 1).double(), target.clamp(0, 1).double()
    C = pred.shape[1]
    g = torch.arange(11, dtype=torch.float64, device=pred.device) - 5
    g = torch.exp(-(g ** 2) / (2 * 1.5 ** 2))
    g = (g / g.sum())
    win = (g[:, None] * g[None, :]).expand(C, 1, 11, 11).contiguous()

    def blur(x):
        return F.conv2d(x, win, groups=C)
    mu1, mu2 = blur(pred), blur(target)
    s11, s22, s12 = blur(pred * pred) - mu1 * mu1, blur(target * target) - mu2 * mu2, blur(pred * target) - mu1 * mu2
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s11 + s22 + c2))
    return m.flatten(1).mean(1)


def build_model(arch: str, scale: int):
    import tpu_superresolution_b200 as srk
    from tpu_superresolution_b200 import synth
    if arch == "swinir":
        cfg = synth.CONFIGS["swinir_x4"].as_kwargs()
        cfg["upscale"] = scale
        return srk.SwinIR(**cfg)
    if arch == "hat":
        cfg = synth.HAT_CONFIGS["hat_x4"].as_kwargs()
        cfg["upscale"] = scale
        return srk.HAT(**cfg)
    cfg = synth.DAT_CONFIGS["dat_x2"].as_kwargs()
    cfg["upscale"] = scale
    return srk.DAT(**cfg)


@torch.no_grad()
def validate(model, pairs, scale: int, device, tile: int = 0, overlap: int = 8, batch: int = 16, save_dir: str = ""):
    """finetune_swinir.py:182-207 on whole images: mean L1, mean PSNR, mean SSIM, seconds."""
    from tpu_superresolution_b200 import tiling
    tiler = tiling.TiledSuperResolver(model, scale=scale, tile=tile, overlap=overlap, batch=batch) if tile > 0 else None
    tot_l1 = tot_psnr = tot_ssim = 0.0
    t0 = time.time()
    for lr_path, hr_path in pairs:
        lr, hr = to_tensor(lr_path).to(device), to_tensor(hr_path).to(device)
        sr = tiler(lr) if tiler is not None and min(lr.shape[2:]) >= 8 else model(lr)
        h, w = min(sr.shape[2], hr.shape[2]), min(sr.shape[3], hr.shape[3])
        sr, hr = sr[:, :, :h, :w].float(), hr[:, :, :h, :w]
        tot_l1 += F.l1_loss(sr, hr).item()
        tot_psnr += batch_psnr(sr, hr).sum().item()
        tot_ssim += ssim(sr, hr).sum().item()
        if save_dir:
            from PIL import Image
            os.makedirs(save_dir, exist_ok=True)
            Image.fromarray((sr[0].clamp(0, 1) * 255).round().byte().permute(1, 2, 0).cpu().numpy()).save(os.path.join(save_dir, lr_path.stem + "_sr.png"))
    n = len(pairs)
    return tot_l1 / n, tot_psnr / n, tot_ssim / n, time.time() - t0


def selftest(device) -> dict:
    """A synthetic BasicSR-style checkpoint ({"params": state_dict}) and PNG pairs in a temp dir through the whole loop; returns the
    metrics of the fused path and of the tiled path (they must agree closely: same model, overlapping-tile stitch)."""
    from PIL import Image
    import tpu_superresolution_b200 as srk
    from tpu_superresolution_b200 import synth
    cfg = synth.CONFIGS["swinir_x4_d2"]
    sd = synth.make_swinir_state_dict(cfg, seed=1234, kind="init")
    with tempfile.TemporaryDirectory() as d:
        torch.save({"params": sd}, os.path.join(d, "net.pth"))
        os.makedirs(os.path.join(d, "lr")); os.makedirs(os.path.join(d, "hr"))
        rng = np.random.default_rng(0)
        for k, (h, w) in enumerate([(40, 56), (72, 64)]):
            hr = rng.integers(0, 256, size=(h * 4, w * 4, 3), dtype=np.uint8)
            lr = np.asarray(Image.fromarray(hr).resize((w, h), Image.BICUBIC))
            Image.fromarray(hr).save(os.path.join(d, "hr", f"img{k}.png"))
            Image.fromarray(lr).save(os.path.join(d, "lr", f"img{k}_x4.png"))
        model = srk.SwinIR(**cfg.as_kwargs()).eval()
        load_checkpoint(model, os.path.join(d, "net.pth"))
        model.to(device)
        pairs = pair_files(os.path.join(d, "lr"), os.path.join(d, "hr"), 4)
        whole = validate(model, pairs, 4, device)
        tiled = validate(model, pairs, 4, device, tile=32, overlap=8, batch=4)
    return {"pairs": len(pairs), "whole": whole[:3], "tiled": tiled[:3]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="swinir", choices=["swinir", "hat", "dat"])
    ap.add_argument("--scale", type=int, default=4)
    ap.add_argument("--weights", default="")
    ap.add_argument("--lr-dir", default="")
    ap.add_argument("--hr-dir", default="")
    ap.add_argument("--tile", type=int, default=0, help="> 0: overlapping-tile inference (TiledSuperResolver) with this LR tile side")
    ap.add_argument("--overlap", type=int, default=8)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--save-dir", default="")
    ap.add_argument("--selftest", action="store_true")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("validate.py: a CUDA device is required (the fused path has no CPU fallback)")
    device = torch.device("cuda")
    if args.selftest:
        print(selftest(device))
        return
    model = build_model(args.arch, args.scale).eval()
    if args.weights:
        load_checkpoint(model, args.weights)
    else:
        print("[weights] none given: random init")
    model.to(device)
    pairs = pair_files(args.lr_dir, args.hr_dir, args.scale)
    l1, psnr, ssim_v, secs = validate(model, pairs, args.scale, device, args.tile, args.overlap, args.batch, args.save_dir)
    print(f"[valid] images={len(pairs)} l1={l1:.6f} psnr={psnr:.3f} dB ssim={ssim_v:.4f} time={secs:.1f}s")


if __name__ == "__main__":
    main()
