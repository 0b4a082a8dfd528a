"""CTA 0's per-head-step timeline (clock64 deltas) of one HAT window-attention launch at the bench size (B = 8, 64x64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tpu_superresolution_b200 import _lib as L, packing, hat as H
from tpu_superresolution_b200 import synth

torch.set_grad_enabled(False)
kind = L.WA_HAT_OCAB if len(sys.argv) > 1 and sys.argv[1] == "ocab" else L.WA_HAT_WMSA
shift = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = synth.HAT_CONFIGS["hat_x4_d2"]
sd = synth.make_hat_state_dict(cfg, seed=99, kind="stress")
pre = "layers.0.residual_group.blocks.1."
B, Hh, W = 8, 64, 64
T = B * Hh * W
x = synth.make_tokens(B, Hh, W, 180, seed=1).cuda()
qw, qb = packing.pack_qkv_planes(sd[pre + "attn.qkv.weight"].cuda(), sd[pre + "attn.qkv.bias"].cuda(), sd[pre + "norm1.weight"].cuda(), sd[pre + "norm1.bias"].cuda())
qkv = torch.empty(9, T, 64, dtype=torch.bfloat16, device="cuda")
L.linear(x, qw, qb, qkv, num_tokens=T, a_mode=L.LIN_A_ROWS, ld_in=180, apply_ln=True, n_chunks=3, out_mode=L.LIN_OUT_PLANES,
         plane_phase_mask=H._KV_PHASE4 if kind == L.WA_HAT_OCAB else 0)
if kind == L.WA_HAT_OCAB:
    tab = packing.pack_bias_table_ocab(sd["layers.0.residual_group.overlap_attn.relative_position_bias_table"].cuda())
else:
    tab = packing.pack_bias_table_wmsa(sd[pre + "attn.relative_position_bias_table"].cuda())
o = torch.empty(3, T, 64, dtype=torch.bfloat16, device="cuda")
def run():
    L.window_attention(qkv[0:3], qkv[3:6], qkv[6:9], tab, o, kind=kind, batch=B, height=Hh, width=W, shift=(shift, shift), mask_shift=shift > 0)
lib = L.load()
for st in (0, 1500, 3000, 4000, 5000, 7000):
    lib.srk_debug_set_winattn_stagger(st)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record(); torch.cuda.synchronize()
    print(f"kind={kind} shift={shift} stagger={st}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch")
lib.srk_debug_set_winattn_stagger(4000)
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
lib.srk_debug_set_timeline(buf.data_ptr())
run()
torch.cuda.synchronize()
lib.srk_debug_set_timeline(0)
t = buf.cpu()[:512].view(8, 64)
names = {0: "step start", 1: "SF seen c0", 2: "sweep1 c0 done", 3: "sweep2 c0 done (PR)", 5: "SF seen c1", 6: "sweep1 c1", 7: "PR c1", 9: "SF c2", 10: "sweep1 c2", 11: "PR c2",
         20: "OF seen", 21: "drain done"}
t0 = min(int(v) for v in t.flatten() if int(v) != 0)
for slot in range(8):
    ev = sorted((int(t[slot, i]), i) for i in names if int(t[slot, i]) != 0)
    if not ev:
        continue
    print(f"--- slot {slot}: head step {slot // 2}, group {slot % 2}")
    prev = ev[0][0]
    for c, i in ev:
        print(f"  {c - t0:8d} (+{c - prev:6d})  {names[i]}")
        prev = c
