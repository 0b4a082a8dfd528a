"""A/B helper: swin_attn / swin_mlp timed alone and alternating (as in the model), shift 0 and 4, x L2-resident or not."""
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
B = 16
aw, av = blk.attn._packed(blk.norm1)
mw, mv = blk.mlp._packed(blk.norm2)
lib = L.load()


def attn(x, shift):
    L.swin_attn(x, x, aw, av, mode=L.MODE_IMAGE, batch=B, height=64, width=64, ld_in=180, ld_out=180, shift=shift,
                mask_mode=L.MASK_SHIFT if shift else L.MASK_NONE)


OPS = os.environ.get("SRK_OPS", "bf16")


def mlp(x):
    L.swin_mlp(x, x, mw, mv, num_tokens=B * 4096, ld_in=180, ld_out=180, operands=OPS)


def timed(fn, n=40):
    for i in range(6):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for nx in (1, 3):
    xs = [synth.make_tokens(B, 64, 64, 180, seed=i).cuda() for i in range(nx)]
    for shift in (0, 4):
        a = timed(lambda i: attn(xs[i % nx], shift))
        m = timed(lambda i: mlp(xs[i % nx]))
        p = timed(lambda i: (attn(xs[i % nx], shift), mlp(xs[i % nx])))
        print(f"nx {nx} shift {shift}: attn {a:6.1f}  mlp {m:6.1f}  attn+mlp alternating {p:6.1f} us", flush=True)
