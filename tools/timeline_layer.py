"""CTA 0's clock64 timeline of one srk_swin_layer_fwd launch (BasicLayer of `depth` blocks, B x 64 x 64): its first 8 work items."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import _lib as L, synth

torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = synth.CONFIGS["swinir_x4"]
m = srk.SwinIR(**cfg.as_kwargs()).eval()
m.load_state_dict(synth.make_swinir_state_dict(cfg, seed=1234, kind="init"), strict=True)
m.cuda()
layer = m.layers[0].residual_group
x = synth.make_tokens(B, 64, 64, 180, seed=1).cuda()
for _ in range(3):
    layer(x, (64, 64))
lib = L.load()
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
lib.srk_debug_set_timeline(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); layer(x, (64, 64)); e1.record()
torch.cuda.synchronize()
lib.srk_debug_set_timeline(0)
print(f"layer call (clone + memset + kernel): {e0.elapsed_time(e1) * 1e3:.1f} us")
starts = buf.cpu()[1024:1152]
t = buf.cpu()[:512].view(8, 64)
names = {0: "item start", 2: "VTF seen", 3: "V epi done", 23: "OF seen", 24: "O epi done", 26: "PJF seen", 28: "rows staged (attn)", 29: "copies read smem",
         27: "item end (all row warps)", 63: "item end", 1: "F1c0 seen", 7: "gelu2 done'", 8: "F2 seen", 9: "rows staged (mlp)", 32: "MMA: waits XA", 33: "MMA: XA seen"}
for hh in range(3):
    names[5 + 3 * hh] = f"SF(h{2*hh}) seen"; names[6 + 3 * hh] = f"softmax(h{2*hh}) done"
mlp_names = {0: "item start", 1: "F1c0 seen", 2: "gelu0 done", 3: "F1c1 seen", 4: "gelu1 done", 5: "F1c2 seen", 6: "gelu2 done", 8: "F2 seen", 9: "rows staged",
             29: "copies read smem", 27: "item end (all row warps)", 63: "item end", 32: "MMA: waits XA", 33: "MMA: XA seen"}
base = int(t[0, 0])
G = min(B * 32, 148)
for n in range(8):
    if int(t[n, 0]) == 0:
        continue
    item = n * G
    is_mlp = (item // (B * 32)) & 1
    nm = mlp_names if is_mlp else names
    ev = sorted((int(t[n, i]), i) for i in nm if int(t[n, i]) != 0)
    print(f"--- CTA0 item {n} (global {item}, {'mlp' if is_mlp else 'attn'}): start +{int(t[n, 0]) - base}")
    prev = int(t[n, 0])
    for c, i in ev:
        print(f"  {c - int(t[n, 0]):8d} (+{c - prev:6d})  {nm[i]}")
        prev = c

G = min(B * 32, 148)
st = [int(v) for v in starts if int(v) != 0]
print("item durations of CTA 0 (cycles), '|' = phase change:")
line = []
for k in range(len(st) - 1):
    ph0, ph1 = (k * G) // (B * 32), ((k + 1) * G) // (B * 32)
    line.append(f"{st[k + 1] - st[k]}{' |' if ph1 != ph0 else ''}")
print(" ".join(line))
if len(st) > 1:
    print(f"total {st[-1] - st[0]} cycles over {len(st) - 1} items")
