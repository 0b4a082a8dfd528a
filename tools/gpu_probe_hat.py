"""Bring-up probe for the HAT path: every stage against the CPU oracle, printed (never asserts) so one GPU call shows all."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import _lib as L, packing, hat as H
from tpu_superresolution_b200 import synth
from oracle import hat_oracle as HO, swinir_oracle as O

torch.set_grad_enabled(False)
torch.backends.cudnn.allow_tf32 = True
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
cfg = synth.HAT_CONFIGS["hat_x4_d2"]
sd = synth.make_hat_state_dict(cfg, seed=99, kind="stress")
fails = 0


def rep(name, got, ref, tol):
    global fails
    got, ref = got.double().cpu(), torch.as_tensor(ref).double()
    err = (got - ref).abs().max().item()
    rel = err / max(ref.abs().max().item(), 1e-12)
    ok = bool(torch.isfinite(got).all()) and rel < tol
    fails += 0 if ok else 1
    print(f"[{name}] max_abs={err:.4e} rel={rel:.4e} ref_max={ref.abs().max():.3f} finite={bool(torch.isfinite(got).all())} -> {'OK' if ok else 'FAIL'}", flush=True)


def sub(pre):
    return {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}


def stage(fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as e:      # noqa
        global fails
        fails += 1
        print(f"[{fn.__name__}] EXCEPTION {type(e).__name__}: {e}", flush=True)


def linear_qkv():
    pre = "layers.0.residual_group.blocks.1."
    for ntok in (128, 300, 2048):
        x = synth.make_tokens(1, 1, ntok, 180, seed=3)[0]
        xn = O.layer_norm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
        ref = xn @ sd[pre + "attn.qkv.weight"].T + sd[pre + "attn.qkv.bias"]
        ref[:, :180] *= 30 ** -0.5 * packing.LOG2E
        qw, qb = packing.pack_qkv_planes(sd[pre + "attn.qkv.weight"].cuda(), sd[pre + "attn.qkv.bias"].cuda(),
                                         sd[pre + "norm1.weight"].cuda(), sd[pre + "norm1.bias"].cuda())
        for mask in (0, H._KV_PHASE4):
            planes = torch.zeros(9, ntok, 64, dtype=torch.bfloat16, device="cuda")
            L.linear(x.cuda(), qw, qb, planes, num_tokens=ntok, a_mode=L.LIN_A_ROWS, ld_in=180, apply_ln=True, n_chunks=3,
                     out_mode=L.LIN_OUT_PLANES, plane_phase_mask=mask)
            got = torch.cat([packing.unswizzle_planes(planes[p:p + 1], 4 if (mask >> p) & 1 else 0)[0] for p in range(9)], 1).float()   # (T, 576)
            got = got.view(ntok, 3, 6, 32)[..., :30].reshape(ntok, 540)
            rep(f"linear_qkv_{ntok}_m{mask}", got, ref, 1.5e-2)


def linear_proj():
    pre = "layers.0.residual_group.blocks.1."
    for ntok in (128, 300, 4096):
        rng = np.random.default_rng(5)
        o = torch.from_numpy(rng.normal(0, 1, size=(ntok, 6, 30)).astype(np.float32))
        ref = o.reshape(ntok, 180) @ sd[pre + "attn.proj.weight"].T + sd[pre + "attn.proj.bias"]
        opad = torch.zeros(ntok, 6, 32)
        opad[..., :30] = o
        planes = opad.reshape(ntok, 3, 64).permute(1, 0, 2).contiguous().to(torch.bfloat16).cuda()
        planes = packing.unswizzle_planes(planes, 0)      # the permutation is an involution: this swizzles
        pw, pb = packing.pack_proj_planes(sd[pre + "attn.proj.weight"].cuda(), sd[pre + "attn.proj.bias"].cuda())
        base = synth.make_tokens(1, 1, ntok, 180, seed=8)[0]
        for res in (False, True):
            y = base.clone().cuda()
            L.linear(planes, pw, pb, y, num_tokens=ntok, a_mode=L.LIN_A_PLANES, n_chunks=1, out_mode=L.LIN_OUT_ROWS, ld_out=180,
                     add_residual=res)
            rep(f"linear_proj_{ntok}_res{int(res)}", y, ref + (base if res else 0), 1.5e-2)


def window_attention():
    g = np.load(os.path.join(G, "kat_hat_window_attention.npz"))
    attn = H.WindowAttention(180, (16, 16), 6).eval()
    attn.load_state_dict(sub("layers.0.residual_group.blocks.1.attn."), strict=True)
    attn.cuda()
    xw = synth.make_tokens(4, 16, 16, 180, seed=5).cuda()
    rep("hat_attn_nomask", attn(xw, H.calculate_rpi_sa(16))[:, ::3], g["y_nomask"], 3e-2)
    mask = torch.from_numpy(g["mask"].astype(np.float32)).cuda()
    rep("hat_attn_mask", attn(xw, None, mask)[:, ::3], g["y_mask"], 3e-2)


def hab():
    g = np.load(os.path.join(G, "kat_hat_hab.npz"))
    xt = synth.make_tokens(2, 32, 48, 180, seed=11).cuda()
    for b, shift, key in ((0, 0, "y_unshifted"), (1, 8, "y_shifted")):
        blk = H.HAB(180, (64, 64), 6, window_size=16, shift_size=shift, mlp_ratio=2.0).eval()
        blk.load_state_dict(sub(f"layers.0.residual_group.blocks.{b}."), strict=True)
        blk.cuda()
        rep(f"hab_{key}", blk(xt, (32, 48))[:, ::5], g[key], 1e-2)


def ocab():
    g = np.load(os.path.join(G, "kat_hat_ocab.npz"))
    xt = synth.make_tokens(2, 32, 48, 180, seed=11).cuda()
    blk = H.OCAB(180, (64, 64), 16, 0.5, 6, mlp_ratio=2).eval()
    blk.load_state_dict(sub("layers.0.residual_group.overlap_attn."), strict=True)
    blk.cuda()
    rep("ocab", blk(xt, (32, 48))[:, ::5], g["y"], 1e-2)


def model():
    for name, kind, seed, B, h, w in [("hat_x4_d2", "init", 1234, 1, 64, 64), ("hat_x4_d2", "stress", 4321, 1, 32, 48),
                                      ("hat_x2_d2", "stress", 77, 1, 20, 27)]:
        c = synth.HAT_CONFIGS[name]
        m = srk.HAT(**c.as_kwargs()).eval()
        m.load_state_dict(synth.make_hat_state_dict(c, seed=seed, kind=kind), strict=True)
        m.cuda()
        lr = synth.make_lr_batch(B, h, w, seed=seed + 1)
        y = m(lr.cuda())
        g = np.load(os.path.join(G, f"{name}_{kind}_{B}x{h}x{w}.npz"))
        got, ref = y.double().cpu(), torch.from_numpy(g["y"]).double()
        err = (got - ref).abs().max().item()
        print(f"[model {name} {kind} {h}x{w}] max_abs={err:.4e} (gate 2e-3) -> {'OK' if err < 2e-3 else 'FAIL'}", flush=True)


for fn in (linear_qkv, linear_proj, window_attention, hab, ocab, model):
    if len(sys.argv) > 1 and fn.__name__ not in sys.argv[1:]:
        continue
    stage(fn)
print("FAILS", fails)
