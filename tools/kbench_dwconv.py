"""Time the two DAT depthwise-conv configurations (v branch: strided input + GELU; SpatialGate: LayerNorm-ed input * gate)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tpu_superresolution_b200 import _lib as L

torch.manual_seed(0)
B, H, W, C = 16, 64, 64, 180
T = B * H * W
qkv = torch.randn(T, 540, device="cuda")
h = torch.randn(T, 360, device="cuda")
w9 = torch.randn(9, C, device="cuda")
sc, sh = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
g, bt = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
stats = torch.empty(T, 2, device="cuda")
out = torch.empty(T, C, device="cuda")
L.row_stats(h, stats, ld_in=360, c_in=180, channels=C, tokens=T, eps=1e-5) if hasattr(L, "row_stats") else None


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


a = timed(lambda: L.dwconv3x3_rows(qkv, w9, sc, sh, out, ld_in=540, c_in=360, ld_out=C, channels=C, batch=B, height=H, width=W, act_gelu=True))
b = timed(lambda: L.dwconv3x3_rows(h, w9, sc, sh, out, ld_in=360, c_in=180, ld_out=C, channels=C, batch=B, height=H, width=W,
                                   ln_stats=stats, ln_gamma=g, ln_beta=bt, gate=h, ld_gate=360, c_gate=0))
print(f"dwconv v-branch (GELU): {a:.1f} us   spatial-gate (LN, gate): {b:.1f} us")

att = torch.randn(T, C, device="cuda"); conv = torch.randn(T, C, device="cuda"); mix = torch.empty(T, C, device="cuda")
cmap = torch.randn(B, C, device="cuda")
w1 = torch.randn(11, C, device="cuda") * 0.1; b1 = torch.randn(11, device="cuda"); w2 = torch.randn(11, device="cuda")
for mode in (0, 1):
    t = timed(lambda: L.dat_mix(att, conv, cmap, w1, b1, w2, 0.1, mix, mode=mode, tokens=T, tokens_per_image=H * W))
    print(f"dat_mix mode {mode}: {t:.1f} us")
