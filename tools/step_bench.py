"""CUDA-graph replay step time of bench workloads (inputs resident, no L2 flush): a quick A/B number for kernel experiments.
   [SRK_LIB=path/to/libsrk.so] python tools/step_bench.py [workload ...]        (default: swinir_x4 hat_x4 dat_x2)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import synth

torch.set_grad_enabled(False)
for name in (sys.argv[1:] or ["swinir_x4", "hat_x4", "dat_x2"]):
    W = bench.WORKLOADS[name]
    cfg, sd, cls = bench._build(W["family"], W["cfg"])
    m = cls(**cfg.as_kwargs()).eval()
    m.load_state_dict(sd, strict=True)
    m.cuda()
    x = synth.make_lr_batch(W["tiles"], 64, 64, seed=1).cuda()
    g = srk.GraphedModel(m)
    for _ in range(5):
        y = g(x)
    torch.cuda.synchronize()
    best = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g(x)
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / 20)
    print(f"{name}: graph replay {min(best):.3f} ms/step (3 x 20 steps: {' '.join(f'{b:.3f}' for b in best)})  checksum {y.double().sum().item():.6f}")
    del g, m
