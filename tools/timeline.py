"""Print CTA 0's phase timeline (clock64 deltas) of one swin_attn / swin_mlp launch at the bench workload size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import _lib as L
from tpu_superresolution_b200 import synth

torch.set_grad_enabled(False)
sd = synth.make_swinir_state_dict(synth.CONFIGS["swinir_x2_d2"], seed=99, kind="init")
pre = "layers.0.residual_group.blocks.1."
blk = srk.SwinTransformerBlock(180, (64, 64), 6, window_size=8, shift_size=4, mlp_ratio=2.0).eval()
st = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
st["attn_mask"] = blk.attn_mask.clone()
blk.load_state_dict(st, strict=True)
blk.cuda()
x = synth.make_tokens(16, 64, 64, 180, seed=1).cuda()
y = torch.empty_like(x)
for _ in range(3):
    blk.forward_into(x, (64, 64), y)
lib = L.load()
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
aw, av = blk.attn._packed(blk.norm1)
mw, mv = blk.mlp._packed(blk.norm2)
names1 = {0: "tile start", 2: "VTF seen", 3: "VT epi done", 23: "OF seen", 24: "O epi done", 25: "LN(next) done", 26: "PJF seen",
          27: "store done", 32: "MMA tile start", 33: "MMA XA seen", 34: "MMA VT issued", 41: "MMA PV5 issued", 42: "MMA OR seen",
          43: "MMA proj issued"}
for hh in range(3):
    names1[4 + 3 * hh] = f"g0 QKep(h{2*hh}) done"; names1[5 + 3 * hh] = f"g0 SF(h{2*hh}) seen"; names1[6 + 3 * hh] = f"g0 softmax(h{2*hh}) done"
for h in range(6):
    names1[35 + h] = f"MMA QKR{h} seen"; names1[44 + h] = f"MMA S{h} issued"; names1[50 + h] = f"U QKF{h} seen"; names1[56 + h] = f"U QKep{h} done"
names1[62] = "U LN(next) done"; names1[1] = "MMA weights of qk(4) landed"; names1[30] = "MMA gemm_qk(4) issued"; names1[31] = "MMA PR(1) seen"; names1[63] = "MMA PV(1) issued"; names1[28] = "rows staged, copies issued"; names1[29] = "copies read smem"
names2 = {0: "tile start", 1: "F1c0 seen", 2: "gelu0 done", 3: "F1c1 seen", 4: "gelu1 done", 5: "F1c2 seen", 6: "gelu2 done", 7: "LN next done", 8: "F2 seen", 9: "store done"}
for nm in (names1, names2):
    nm.update({13: "kernel entry", 14: "prologue done (TMEM, barriers)", 15: "PDL wait done", 16: "role loop done", 17: "all warps done"})
for which, names in (("attn", names1), ("mlp", names2)):
    buf.zero_()
    lib.srk_debug_set_timeline(buf.data_ptr())
    if which == "attn":
        L.swin_attn(x, y, aw, av, mode=L.MODE_IMAGE, batch=16, height=64, width=64, ld_in=180, ld_out=180, shift=4, mask_mode=L.MASK_SHIFT)
    else:
        L.swin_mlp(y, y, mw, mv, num_tokens=16 * 4096, ld_in=180, ld_out=180)
    torch.cuda.synchronize()
    lib.srk_debug_set_timeline(0)
    t = buf.cpu()[:512].view(8, 64)
    if which == "attn":
        w = buf.cpu()[1600:1700].tolist()
        if w[8]:
            nt = max(w[9], 1)
            print(f"--- attn CTA0 wait profile (cycles per tile, averaged over {w[9]} tiles incl. the first)")
            lab = ["wait XA", "GEMM issue (V, q|k, proj: incl. weight-slab waits)", "wait VTD", "wait QKA/QKB free", "wait OALL"]
            print("  GEMM issuer : loop %d | " % (w[8] // nt) + ", ".join(f"{lab[c]} {w[c] // nt}" for c in range(5)))
            lab = ["wait VTD", "wait QKR (6)", "wait PR (6)", "wait OR (2)", "issue S (6)", "issue PV (6)"]
            print("  attn issuer : loop %d | " % (w[88] // nt) + ", ".join(f"{lab[c]} {w[80 + c] // nt}" for c in range(6)))
            lab = ["wait VTF", "wait SF (3)", "wait OF (3)", "wait PJF", "store drain + bar"]
            for name, o in (("group 0     ", 20), ("group 1     ", 40)):
                print(f"  {name}: loop %d | " % (w[o + 8] // nt) + ", ".join(f"{lab[c]} {w[o + c] // nt}" for c in range(5)))
            lab = ["wait DRAIN", "wait QKF (3)", "wait SF (6)", "LN next tile"]
            print("  utility     : loop %d | " % (w[68] // nt) + ", ".join(f"{lab[c]} {w[60 + c] // nt}" for c in range(4)))
    for it in range(4):
        ev = sorted((int(t[it, i]), i) for i in names if int(t[it, i]) != 0)
        if not ev:
            continue
        t0 = ev[0][0]
        print(f"--- {which} CTA0 tile iteration {it}")
        prev = t0
        for c, i in ev:
            print(f"  {c - t0:8d} (+{c - prev:6d})  {names[i]}")
            prev = c
