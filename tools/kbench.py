"""Micro-benchmark of swin_attn / swin_mlp launches at the bench workload size (B=16, 64x64), with optional start skew sweep."""
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
B = int(os.environ.get("KB_BATCH", "16"))
NX = int(os.environ.get("KB_NX", "3"))
xs = [synth.make_tokens(B, 64, 64, 180, seed=i).cuda() for i in range(NX)]
y = torch.empty_like(xs[0])
INPLACE = os.environ.get('KB_INPLACE', '1') == '1'
aw, av = blk.attn._packed(blk.norm1)
mw, mv = blk.mlp._packed(blk.norm2)
lib = L.load()

def run(kind, n=30):
    for i in range(5):
        (L.swin_attn(xs[i % NX], xs[i % NX] if INPLACE else y, aw, av, mode=L.MODE_IMAGE, batch=B, height=64, width=64, ld_in=180, ld_out=180, shift=4, mask_mode=L.MASK_SHIFT)
         if kind == "attn" else L.swin_mlp(xs[i % NX], xs[i % NX] if INPLACE else y, mw, mv, num_tokens=B * 4096, ld_in=180, ld_out=180))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        (L.swin_attn(xs[i % NX], xs[i % NX] if INPLACE else y, aw, av, mode=L.MODE_IMAGE, batch=B, height=64, width=64, ld_in=180, ld_out=180, shift=4, mask_mode=L.MASK_SHIFT)
         if kind == "attn" else L.swin_mlp(xs[i % NX], xs[i % NX] if INPLACE else y, mw, mv, num_tokens=B * 4096, ld_in=180, ld_out=180))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for sk in [0]:
    lib.srk_debug_set_stagger(sk, sk)
    print(f"stagger {sk:5d}: attn {run('attn'):7.1f} us   mlp {run('mlp'):7.1f} us", flush=True)
