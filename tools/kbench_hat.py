"""HAT x4 (BASELINE configs[2]: B = 8 tiles of 64x64) step time and per-kernel CUDA-event timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import _lib as L
from tpu_superresolution_b200 import synth

torch.set_grad_enabled(False)
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
name = sys.argv[1] if len(sys.argv) > 1 else "hat_x4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = synth.HAT_CONFIGS[name]
m = srk.HAT(**cfg.as_kwargs()).eval()
m.load_state_dict(synth.make_hat_state_dict(cfg, seed=1234, kind="init"), strict=True)
m.cuda()
x = synth.make_lr_batch(B, 64, 64, seed=1).cuda()
for _ in range(3):
    y = m(x)
torch.cuda.synchronize()
assert torch.isfinite(y).all()
evs = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m(x); e1.record()
    evs.append((e0, e1))
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in evs)
mpix = B * (64 * cfg.upscale) ** 2 / 1e6
print(f"{name} B={B}: step median {ms[len(ms)//2]:.3f} ms  best {ms[0]:.3f} ms  -> {mpix / ms[len(ms)//2] * 1e3:.1f} Mpix/s")
prof = {}
L.PROFILE = prof
for _ in range(3):
    m(x)
torch.cuda.synchronize()
L.PROFILE = None
tot = 0.0
for k, v in prof.items():
    t = [a.elapsed_time(b) for a, b in v]
    print(f"  {k:18s} n/step={len(t)//3:4d}  avg {sum(t)/len(t)*1e3:8.1f} us   per step {sum(t)/3:7.3f} ms")
    tot += sum(t) / 3
print(f"  libsrk total per step {tot:.3f} ms")
