"""Tight mode (SwinIR.set_precision("fp16")) at the BASELINE workload (SwinIR x4, 16 x 64 x 64): ms per step with the split
convolutions (convs.SplitConv3x3) and, for A/B, with fp32 library convolutions (SRK_TIGHT_CONV=library); eager and CUDA-graph
replay; plus the max abs difference of the two and of the default mode.   python tools/tight_bench.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import tpu_superresolution_b200 as srk                                   # noqa: E402
from tpu_superresolution_b200 import synth                               # noqa: E402


def timed(fn, x, steps):
    for _ in range(3):
        fn(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        y = fn(x)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, y


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    cfg = synth.CONFIGS["swinir_x4"]
    m = srk.SwinIR(**cfg.as_kwargs()).eval()
    m.load_state_dict(synth.make_swinir_state_dict(cfg, seed=1), strict=True)
    m.cuda()
    x = torch.rand(16, 3, 64, 64, device="cuda")
    mpix = 16 * 256 * 256 / 1e6
    with torch.no_grad():
        outs = {}
        for label, prec, env in (("default", "bf16", "split"), ("tight/split", "fp16", "split"), ("tight/library", "fp16", "library")):
            os.environ["SRK_TIGHT_CONV"] = env
            m.set_precision(prec)
            ms_e, y = timed(m, x, steps)
            g = srk.GraphedModel(m)
            ms_g, yg = timed(g, x, steps)
            outs[label] = y.clone()
            print(f"{label:14s} eager {ms_e:7.3f} ms ({mpix / ms_e * 1e3:6.1f} Mpix/s)   graph {ms_g:7.3f} ms ({mpix / ms_g * 1e3:6.1f} Mpix/s)"
                  f"   graph == eager: {torch.equal(y, yg)}")
        ref = outs["tight/library"]
        for k in ("default", "tight/split"):
            print(f"max abs {k} - tight/library: {(outs[k] - ref).abs().max().item():.3e}")


if __name__ == "__main__":
    main()
