"""CTA 0's timeline of one token_linear launch (qkv planes from fp32 rows + LN) at a given token count."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tpu_superresolution_b200 import _lib as L, packing
from tpu_superresolution_b200 import synth

torch.set_grad_enabled(False)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
cfg = synth.HAT_CONFIGS["hat_x4_d2"]
sd = synth.make_hat_state_dict(cfg, seed=99, kind="stress")
pre = "layers.0.residual_group.blocks.1."
x = synth.make_tokens(1, 1, T, 180, seed=1)[0].cuda()
qw, qb = packing.pack_qkv_planes(sd[pre + "attn.qkv.weight"].cuda(), sd[pre + "attn.qkv.bias"].cuda(), sd[pre + "norm1.weight"].cuda(), sd[pre + "norm1.bias"].cuda())
pw, pb = packing.pack_proj_planes(sd[pre + "attn.proj.weight"].cuda(), sd[pre + "attn.proj.bias"].cuda())
qkv = torch.empty(9, T, 64, dtype=torch.bfloat16, device="cuda")
y = x.clone()
lib = L.load()
def qkv_run():
    L.linear(x, qw, qb, qkv, num_tokens=T, a_mode=L.LIN_A_ROWS, ld_in=180, apply_ln=True, n_chunks=3, out_mode=L.LIN_OUT_PLANES)
def proj_run():
    L.linear(qkv[0:3], pw, pb, y, num_tokens=T, a_mode=L.LIN_A_PLANES, n_chunks=1, out_mode=L.LIN_OUT_ROWS, ld_out=180, add_residual=True)
for name, fn in (("qkv", qkv_run), ("proj", proj_run)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name} T={T}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
    buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
    lib.srk_debug_set_timeline(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.srk_debug_set_timeline(0)
    t = buf.cpu()[:512].view(8, 64)
    t0 = int(t[0, 62])
    print(f"  kernel-body start -> first LN done: {int(t[0, 63]) - t0}")
    for it in range(8):
        ev = [(int(t[it, i]), i) for i in range(0, 13) if int(t[it, i]) != 0]
        if not ev:
            continue
        print(f"  tile {it}: " + "  ".join(f"[{i}] {c - t0}" for c, i in ev))

# DAT: fp32 rows (no LayerNorm) -> 180 fp32 rows added into the residual stream (proj / fc2 of the DAT blocks)
rw, rb = packing.pack_rows_linear(sd[pre + "attn.proj.weight"].cuda(), sd[pre + "attn.proj.bias"].cuda())
src = x.clone()
def rows_run():
    L.linear(src, rw, rb, y, num_tokens=T, a_mode=L.LIN_A_ROWS, ld_in=180, apply_ln=False, n_chunks=1, out_mode=L.LIN_OUT_ROWS, ld_out=180, add_residual=True)
for _ in range(3):
    rows_run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    rows_run()
e1.record(); torch.cuda.synchronize()
print(f"rows->rows+res T={T}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
lib.srk_debug_set_timeline(buf.data_ptr())
rows_run()
torch.cuda.synchronize()
lib.srk_debug_set_timeline(0)
t = buf.cpu()[:512].view(8, 64)
t0 = int(t[0, 62])
print(f"  kernel-body start -> first LN done: {int(t[0, 63]) - t0}")
for it in range(8):
    ev = [(int(t[it, i]), i) for i in range(0, 13) if int(t[it, i]) != 0]
    if ev:
        print(f"  tile {it}: " + "  ".join(f"[{i}] {c - t0}" for c, i in ev))
