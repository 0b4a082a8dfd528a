"""GPU-kernel time breakdown of one eager step of a bench workload (torch.profiler / CUPTI): all kernels, not only libsrk's."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from tpu_superresolution_b200 import synth

torch.set_grad_enabled(False)
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
name = sys.argv[1] if len(sys.argv) > 1 else "swinir_x4"
W = bench.WORKLOADS[name]
cfg, sd, cls = bench._build(W["family"], W["cfg"])
m = cls(**cfg.as_kwargs()).eval()
m.load_state_dict(sd, strict=True)
m.cuda()
x = synth.make_lr_batch(W["tiles"], 64, 64, seed=1).cuda()
for _ in range(3):
    m(x)
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        m(x)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / N, e.count // N) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"{name}: sum of GPU kernel time per step {tot / 1e3:.3f} ms")
for k, t, n in rows[:25]:
    print(f"  {t / 1e3:8.3f} ms  n={n:4d}  {k[:110]}")
