"""Run forwards of a bench workload; the LAST one is bracketed by cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "hat_x4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W = bench.WORKLOADS[name]
from tpu_superresolution_b200 import synth
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
cfg, sd, cls = bench._build(W["family"], W["cfg"])
m = cls(**cfg.as_kwargs()).eval()
m.load_state_dict(sd, strict=True)
m.cuda()
x = synth.make_lr_batch(W["tiles"], 64, 64, seed=2).cuda()
with torch.no_grad():
    for i in range(n):
        if i == n - 1:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
        y = m(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok", tuple(y.shape), float(y.mean()))
