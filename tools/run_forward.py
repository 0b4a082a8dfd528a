"""Run N SwinIR x4 forwards on B=16 64x64 tiles (the bench workload); used under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = synth.CONFIGS["swinir_x4"]
m = srk.SwinIR(**cfg.as_kwargs()).eval()
m.load_state_dict(synth.make_swinir_state_dict(cfg, seed=1234, kind="init"), strict=True)
m.cuda()
x = synth.make_lr_batch(16, 64, 64, seed=2).cuda()
with torch.no_grad():
    for _ in range(n):
        y = m(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.mean()))
