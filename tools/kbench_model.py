"""Step time and per-libsrk-kernel CUDA-event timings of a bench workload (eager)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tpu_superresolution_b200 import _lib as L
from tpu_superresolution_b200 import synth

torch.set_grad_enabled(False)
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
name = sys.argv[1] if len(sys.argv) > 1 else "dat_x2"
W = bench.WORKLOADS[name]
cfg, sd, cls = bench._build(W["family"], W["cfg"])
m = cls(**cfg.as_kwargs()).eval()
m.load_state_dict(sd, strict=True)
m.cuda()
x = synth.make_lr_batch(W["tiles"], 64, 64, seed=1).cuda()
for _ in range(3):
    y = m(x)
torch.cuda.synchronize()
evs = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m(x); e1.record()
    evs.append((e0, e1))
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in evs)
print(f"{name}: eager step median {ms[len(ms)//2]:.3f} ms")
prof = {}
L.PROFILE = prof
L.PROFILE_DETAIL = True
for _ in range(3):
    m(x)
torch.cuda.synchronize()
L.PROFILE = None
tot = 0.0
for k, v in sorted(prof.items(), key=lambda kv: -sum(a.elapsed_time(b) for a, b in kv[1])):
    t = [a.elapsed_time(b) for a, b in v]
    print(f"  {k:18s} n/step={len(t)//3:4d}  avg {sum(t)/len(t)*1e3:8.1f} us   per step {sum(t)/3:7.3f} ms")
    tot += sum(t) / 3
print(f"  libsrk total per step {tot:.3f} ms")
