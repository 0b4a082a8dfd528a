"""CTA 0's clock64 timeline of one srk_conv3x3_fwd launch (180 -> 180, B = 16, 64 x 64): per patch the MMA warp's waits for
every input box, the total weight-slab wait, and the epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tpu_superresolution_b200 import _lib as L, packing

torch.set_grad_enabled(False)
lib = L.load()
B, H, W = 16, 64, 64
cin, cout = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (180, 180)
w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.03
ws, bp, meta = packing.pack_conv3x3(w, None)
x16 = (torch.randn(B * H * W, 64 * meta["k_atoms"], device="cuda") * 0.5).half()
out = torch.zeros(B * H * W, cout, device="cuda")
run = lambda: L.conv3x3(x16, ws, bp, out, batch=B, height=H, width=W, k_atoms=meta["k_atoms"], np_=meta["np"], cout=cout,
                        out_mode=L.CONV_OUT_ROWS_F32, ld_out=cout)
for _ in range(3):
    run()
torch.cuda.synchronize()
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
lib.srk_debug_set_timeline(buf.data_ptr())
run()
torch.cuda.synchronize()
lib.srk_debug_set_timeline(0)
t = buf.cpu()[:512].view(8, 64)
base = int(t[0, 0])
for it in range(4):
    if int(t[it, 0]) == 0:
        continue
    r = [int(v) for v in t[it]]
    print(f"--- patch iteration {it}: start +{r[0] - base}, acc_empty wait {r[1] - r[0]}")
    for s in range(3 * meta["k_atoms"]):
        if s < 12:
            nxt = r[2 + 2 * (s + 1)] if s + 1 < min(12, 3 * meta["k_atoms"]) else r[30]
            print(f"   step {s}: box wait {r[3 + 2 * s] - r[2 + 2 * s]:6d}, step total {nxt - r[2 + 2 * s]:6d}")
    print(f"   mainloop {r[30] - r[1]} cycles, of which weight waits {r[31]}; acc_full seen +{r[40] - r[30]} after the last issue; epilogue {r[41] - r[40]}")
