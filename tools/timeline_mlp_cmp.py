"""MMA-warp and row-warp stamps of MLP work: swin_mlp_kernel alone vs the MLP items of swin_layer_kernel (CTA 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import _lib as L, synth, swinir
torch.set_grad_enabled(False)
cfg = synth.CONFIGS["swinir_x4"]
m = srk.SwinIR(**cfg.as_kwargs()).eval()
m.load_state_dict(synth.make_swinir_state_dict(cfg, seed=1234, kind="init"), strict=True)
m.cuda()
layer = m.layers[0].residual_group
x = synth.make_tokens(16, 64, 64, 180, seed=1).cuda()
lib = L.load()
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
names = {0: "row: item start", 1: "row: F1c0 seen", 2: "row: gelu0 done", 3: "row: F1c1 seen", 4: "row: gelu1 done", 5: "row: F1c2 seen", 6: "row: gelu2 done",
         8: "row: F2 seen", 9: "row: rows staged", 32: "MMA: item start", 33: "MMA: XA seen", 34: "MMA: fc1 c0,c1 issued", 35: "MMA: HR0 seen",
         36: "MMA: fc2(0), fc1 c2 issued", 37: "MMA: HR1 seen", 38: "MMA: fc2(1) issued", 39: "MMA: HR2 seen", 40: "MMA: fc2(2) issued"}
for mode in ("layer", "mlp alone"):
    swinir.USE_LAYER_KERNEL = mode == "layer"
    for _ in range(2):
        layer(x, (64, 64))
    torch.cuda.synchronize()
    buf.zero_()
    lib.srk_debug_set_timeline(buf.data_ptr())
    if mode == "layer":
        layer(x, (64, 64))
        rows = (5, 6)          # CTA 0: items 4..6 are MLP items of block 0
    else:
        blk = layer.blocks[0]
        y = x.clone()
        mw, mv = blk.mlp._packed(blk.norm2)
        L.swin_mlp(y, y, mw, mv, num_tokens=16 * 4096, ld_in=180, ld_out=180)
        rows = (1, 2)
    torch.cuda.synchronize()
    lib.srk_debug_set_timeline(0)
    t = buf.cpu()[:512].view(8, 64)
    for n in rows:
        ev = sorted((int(t[n, i]), i) for i in names if int(t[n, i]) != 0)
        if not ev:
            continue
        base = int(t[n, 0]) if int(t[n, 0]) else ev[0][0]
        print(f"--- {mode}: CTA0 item/tile {n}")
        prev = ev[0][0]
        for c, i in ev:
            print(f"  {c - base:8d} (+{c - prev:6d})  {names[i]}")
            prev = c
