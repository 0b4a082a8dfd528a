"""Where does the remaining error come from?  SwinIR full depth vs the reference golden for operand type x conv path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import convs, synth

torch.set_grad_enabled(False)
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for cfg_name in ("swinir_x2", "swinir_x4"):
    for kind, seed in (("init", 1234), ("stress", 4321)):
        cfg = synth.CONFIGS[cfg_name]
        sd = synth.make_swinir_state_dict(cfg, seed=seed, kind=kind)
        m = srk.SwinIR(**cfg.as_kwargs()).eval()
        m.load_state_dict(sd, strict=True)
        m.cuda()
        lr = synth.make_lr_batch(1, 64, 64, seed=seed + 1).cuda()
        ref = torch.from_numpy(np.load(os.path.join(GOLDEN, f"{cfg_name}_{kind}_1x64x64.npz"))["y"])
        for ops in ("bf16", "fp16"):
            m.set_precision(ops)
            m.precision = "probe"          # the flags below pick the convolution path, not SwinIR.forward's tight-mode branch
            for conv in ("fused", "split", "cudnn-fp32", "cudnn-tf32"):      # fused: plain fp16 operands; split: hi / lo pairs (tight mode)
                convs.USE_FUSED_CONV = conv in ("fused", "split")
                for mod in [m] + list(m.layers):
                    mod.split_conv = conv == "split"
                torch.backends.cudnn.allow_tf32 = conv == "cudnn-tf32"
                torch.backends.cuda.matmul.allow_tf32 = False
                y = m(lr).cpu()
                print(f"{cfg_name} {kind:6s} operands {ops} conv {conv:10s}: max abs err {(y - ref).abs().max().item():.3e}  rms {(y - ref).pow(2).mean().sqrt().item():.3e}", flush=True)
        convs.USE_FUSED_CONV = True
