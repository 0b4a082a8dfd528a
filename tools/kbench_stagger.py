import os, sys
sys.path.insert(0, "/root/repo")
sys.path.insert(0, os.getcwd())
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
B = 16
aw, av = blk.attn._packed(blk.norm1)
mw, mv = blk.mlp._packed(blk.norm2)
lib = L.load()
x = synth.make_tokens(B, 64, 64, 180, seed=0).cuda()
def attn(): L.swin_attn(x, x, aw, av, mode=L.MODE_IMAGE, batch=B, height=64, width=64, ld_in=180, ld_out=180, shift=4, mask_mode=L.MASK_SHIFT)
def mlp(): L.swin_mlp(x, x, mw, mv, num_tokens=B * 4096, ld_in=180, ld_out=180, operands=blk.mlp.operands)
def timed(fn, n=40):
    for i in range(6): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for sa in (0, 1000, 2500, 5000):
    for sm in (0, 1000, 3000):
        lib.srk_debug_set_stagger(sa, sm)
        print(f"stagger attn {sa:5d} mlp {sm:5d}: attn {timed(attn):6.1f} mlp {timed(mlp):6.1f} pair {timed(lambda: (attn(), mlp())):6.1f} us", flush=True)
