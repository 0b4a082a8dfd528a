"""Stress the image-progress-counter path: many SwinIR x4 B=16 forwards (eager and CUDA-graph replay) must be bit-identical to
the first one and to the whole-grid-ordered run (SRK_BLOCK_SYNC=0 semantics, selected at run time through swinir.USE_BLOCK_SYNC)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import swinir
from tpu_superresolution_b200 import synth

torch.set_grad_enabled(False)
cfg = synth.CONFIGS["swinir_x4"]
m = srk.SwinIR(**cfg.as_kwargs()).eval()
m.load_state_dict(synth.make_swinir_state_dict(cfg, seed=1234, kind="init"), strict=True)
m.cuda()
x = synth.make_lr_batch(16, 64, 64, seed=2).cuda()
swinir.USE_BLOCK_SYNC = False
ref = m(x).clone()
swinir.USE_BLOCK_SYNC = True
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
bad = 0
for i in range(n):
    y = m(x)
    if not torch.equal(y, ref):
        bad += 1
        print("eager mismatch at", i, float((y - ref).abs().max()))
g = srk.GraphedModel(m)
for i in range(n):
    y = g(x)
    if not torch.equal(y, ref):
        bad += 1
        print("graph mismatch at", i, float((y - ref).abs().max()))
torch.cuda.synchronize()
print("stress_sync:", "OK" if bad == 0 else f"{bad} MISMATCHES", "over", 2 * n, "forwards")
