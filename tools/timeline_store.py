import os, sys
sys.path.insert(0, "/root/repo")
import torch
import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import _lib as L, synth
torch.set_grad_enabled(False)
cfg = synth.CONFIGS["swinir_x4"]
m = srk.SwinIR(**cfg.as_kwargs()).eval()
m.load_state_dict(synth.make_swinir_state_dict(cfg, seed=1234, kind="init"), strict=True)
m.cuda()
layer = m.layers[0].residual_group
x = synth.make_tokens(16, 64, 64, 180, seed=1).cuda()
lib = L.load()
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
from tpu_superresolution_b200 import swinir
for mode in (True, False):
    swinir.USE_LAYER_KERNEL = mode
    for _ in range(2):
        layer(x, (64, 64))
    torch.cuda.synchronize()
    buf.zero_()
    lib.srk_debug_set_timeline(buf.data_ptr())
    if mode:
        layer(x, (64, 64))
    else:
        blk = layer.blocks[0]
        y = x.clone()
        aw, av = blk.attn._packed(blk.norm1)
        L.swin_attn(y, y, aw, av, mode=L.MODE_IMAGE, batch=16, height=64, width=64, ld_in=180, ld_out=180, shift=0, mask_mode=L.MASK_NONE)
    torch.cuda.synchronize()
    lib.srk_debug_set_timeline(0)
    t = buf.cpu()[:512].view(8, 64)
    print("layer kernel" if mode else "swin_attn_kernel alone")
    for n in range(4):
        r = [int(v) for v in t[n]]
        if r[26]:
            print(f"  item {n}: PJF->chunks written {r[50]-r[26]}, ->quadrant barrier {r[51]-r[50]}, ->copies issued (staged) {r[28]-r[51]}")
