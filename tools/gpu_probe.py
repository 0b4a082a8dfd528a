"""Step-by-step bring-up probe for a B200 box: each step runs in its own process with a timeout so a trap or a
deadlock in one kernel does not hide the others.  Writes gpurun_out/probe_*.npz on mismatch for offline analysis.

    python tools/gpu_probe.py            # all steps
    python tools/gpu_probe.py mlp_raw    # one step (child mode)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")

STEPS = ["aux", "mlp_raw", "mlp_full", "attn_win", "attn_mask", "block_unshifted", "block_shifted", "model_small"]


def _report(name, y, ref, tol):
    import numpy as np
    import torch
    y, ref = y.detach().float().cpu(), torch.as_tensor(ref).float()
    err = (y - ref).abs()
    rel = err.max().item() / max(ref.abs().max().item(), 1e-12)
    ok = bool(torch.isfinite(y).all()) and rel < tol
    print(f"[{name}] max_abs={err.max().item():.4e} rel={rel:.4e} ref_max={ref.abs().max().item():.3f} "
          f"finite={bool(torch.isfinite(y).all())} -> {'OK' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        os.makedirs(OUT, exist_ok=True)
        np.savez_compressed(os.path.join(OUT, f"probe_{name}.npz"), y=y.numpy(), ref=ref.numpy())
    return ok


def child(step):
    import numpy as np
    import torch
    import tpu_superresolution_b200 as srk
    from tpu_superresolution_b200 import _lib as L
    from tpu_superresolution_b200 import synth
    from oracle import swinir_oracle as O
    torch.set_grad_enabled(False)
    sd = synth.make_swinir_state_dict(synth.CONFIGS["swinir_x2_d2"], seed=99, kind="stress")
    bsd = lambda pre: {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    ok = True
    if step == "aux":
        x = synth.make_tokens(3, 5, 7, 180, seed=1).cuda()
        w, b = torch.rand(180, device="cuda") + 0.5, torch.randn(180, device="cuda")
        y = torch.empty_like(x)
        L.layernorm(x, y, w, b, num_tokens=105, ld_in=180, ld_out=180)
        ok &= _report("layernorm", y, O.layer_norm(x.cpu(), w.cpu(), b.cpu()), 1e-5)
        xs = torch.randn(2, 256, 9, 4, device="cuda").contiguous(memory_format=torch.channels_last)
        ok &= _report("pixelshuffle", srk.PixelShuffle(2)(xs), O.pixel_shuffle(xs.cpu(), 2), 1e-9)
    elif step in ("mlp_raw", "mlp_full"):
        pre = "layers.0.residual_group.blocks.0."
        if step == "mlp_raw":
            m = srk.Mlp(180, 360).eval()
            m.load_state_dict(bsd(pre + "mlp."), strict=True)
            m.cuda()
            for ntok in (128, 300, 128 * 200):
                x = synth.make_tokens(1, 1, ntok, 180, seed=7)[0]
                ok &= _report(f"mlp_raw_{ntok}", m(x.cuda()), O.mlp(x, sd, pre + "mlp."), 1.5e-2)
        else:
            blk = srk.SwinTransformerBlock(180, (16, 16), 6, window_size=8, shift_size=0, mlp_ratio=2.0).eval()
            blk.load_state_dict(bsd(pre), strict=True)
            blk.cuda()
            x = synth.make_tokens(1, 16, 16, 180, seed=8)
            w, v = blk.mlp._packed(blk.norm2)
            y = torch.empty_like(x).cuda()
            L.swin_mlp(x.cuda(), y, w, v, num_tokens=256, ld_in=180, ld_out=180)
            ref = x + O.mlp(O.layer_norm(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"]), sd, pre + "mlp.")
            ok &= _report("mlp_full", y, ref, 1e-2)
    elif step in ("attn_win", "attn_mask"):
        g = np.load(os.path.join(ROOT, "tests", "golden", "kat_window_attention.npz"))
        pre = "layers.0.residual_group.blocks.1.attn."
        attn = srk.WindowAttention(180, (8, 8), 6).eval()
        attn.load_state_dict(bsd(pre), strict=True)
        attn.cuda()
        xw = synth.make_tokens(8, 8, 8, 180, seed=5).cuda()
        if step == "attn_win":
            ok &= _report("attn_nomask", attn(xw), g["y_nomask"], 2e-2)
            ok &= _report("attn_nomask_odd", attn(xw[:7]), g["y_nomask"][:7], 2e-2)
        else:
            ok &= _report("attn_mask", attn(xw, torch.from_numpy(g["mask"]).cuda()), g["y_mask"], 2e-2)
    elif step in ("block_unshifted", "block_shifted"):
        for tag, b_idx, shift, res, x_size in ([("unshifted", 0, 0, (16, 24), (16, 24))] if step == "block_unshifted" else
                                               [("shifted", 1, 4, (16, 24), (16, 24)), ("shifted_nonnative", 1, 4, (16, 24), (24, 16)),
                                                ("shifted_64", 1, 4, (64, 64), (64, 64))]):
            g = np.load(os.path.join(ROOT, "tests", "golden", f"kat_block_{tag}.npz"))
            blk = srk.SwinTransformerBlock(180, res, 6, window_size=8, shift_size=shift, mlp_ratio=2.0).eval()
            st = bsd(f"layers.0.residual_group.blocks.{b_idx}.")
            if shift:
                st["attn_mask"] = blk.attn_mask.clone()
            blk.load_state_dict(st, strict=True)
            blk.cuda()
            B = 2 if x_size != (64, 64) else 1
            xt = synth.make_tokens(B, x_size[0], x_size[1], 180, seed=11).cuda()
            y = blk(xt, x_size)
            ok &= _report(f"block_{tag}", y[:, ::7] if x_size == (64, 64) else y, g["y"], 1.5e-2)
    elif step == "model_small":
        for name, kind, seed, B, h, w in [("swinir_x4_d2", "init", 1234, 2, 64, 64), ("swinir_x2", "init", 1234, 1, 64, 64)]:
            g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}_{kind}_{B}x{h}x{w}.npz"))
            cfg = synth.CONFIGS[name]
            m = srk.SwinIR(**cfg.as_kwargs()).eval()
            m.load_state_dict(synth.make_swinir_state_dict(cfg, seed=seed, kind=kind), strict=True)
            m.cuda()
            y = m(synth.make_lr_batch(B, h, w, seed=seed + 1).cuda())
            err = (y.cpu() - torch.from_numpy(g["y"])).abs().max().item()
            print(f"[model {name}] max_abs={err:.3e} (gate 2e-3)", flush=True)
            ok &= err <= 2e-3
    torch.cuda.synchronize()
    print(f"STEP {step}: {'PASS' if ok else 'FAIL'}", flush=True)
    return 0 if ok else 1


def main():
    if len(sys.argv) > 1 and sys.argv[1] in STEPS:
        sys.exit(child(sys.argv[1]))
    os.makedirs(OUT, exist_ok=True)
    results = {}
    for s in STEPS:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), s], timeout=240, capture_output=True, text=True)
            rc, out = p.returncode, p.stdout + p.stderr[-3000:]
        except subprocess.TimeoutExpired as e:
            rc, out = -9, f"TIMEOUT\n{(e.stdout or b'')[-2000:]}\n{(e.stderr or b'')[-2000:]}"
        results[s] = rc
        print(f"===== {s}: rc={rc} ({time.time() - t0:.1f}s)\n{out}", flush=True)
    print("SUMMARY", results)


if __name__ == "__main__":
    main()
