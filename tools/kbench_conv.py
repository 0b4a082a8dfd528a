#!/usr/bin/env python
"""Back-to-back timing of srk_conv3x3_fwd at the SwinIR x4 B=16 shapes (CUDA events, L2-resident and flushed), next to cuDNN."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tpu_superresolution_b200 import _lib as L, packing

L.load()
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20, do_flush=False):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


cases = [  # name, B, H, W, cin, cout, mode
    ("body 180->180 @64", 16, 64, 64, 180, 180, L.CONV_OUT_ROWS_F32),
    ("before_up 180->64 @64", 16, 64, 64, 180, 64, L.CONV_OUT_NHWC_F16),
    ("up1 64->256+PS @64", 16, 64, 64, 64, 256, L.CONV_OUT_SHUFFLE2_F16),
    ("up2 64->256+PS @128", 16, 128, 128, 64, 256, L.CONV_OUT_SHUFFLE2_F16),
    ("last 64->3 @256", 16, 256, 256, 64, 3, L.CONV_OUT_IMAGE),
    ("cab1 180->60 @64 (B=8)", 8, 64, 64, 180, 60, L.CONV_OUT_NHWC_F16),
    ("cab2 60->180 @64 (B=8)", 8, 64, 64, 60, 180, L.CONV_OUT_ROWS_F32),
]
with torch.no_grad():
    for name, B, H, W, cin, cout, mode in cases:
        w = torch.randn(cout, cin, 3, 3, device=dev) * 0.03
        b = torch.randn(cout, device=dev) * 0.1
        ws, bp, meta = packing.pack_conv3x3(w, b, pixel_shuffle=(mode == L.CONV_OUT_SHUFFLE2_F16))
        cp = 64 * meta["k_atoms"]
        x16 = (torch.randn(B * H * W, cp, device=dev) * 0.5).half()
        if mode == L.CONV_OUT_ROWS_F32:
            out, ld = torch.zeros(B * H * W, cout, device=dev), cout
        elif mode == L.CONV_OUT_NHWC_F16:
            out, ld = torch.zeros(B * H * W, 64, dtype=torch.float16, device=dev), 64
        elif mode == L.CONV_OUT_SHUFFLE2_F16:
            out, ld = torch.zeros(B * 4 * H * W, 64, dtype=torch.float16, device=dev), 64
        else:
            out, ld = torch.zeros(B * H * W, 3, device=dev), 3
        fn = lambda: L.conv3x3(x16, ws, bp, out, batch=B, height=H, width=W, k_atoms=meta["k_atoms"], np_=meta["np"], cout=cout, out_mode=mode, ld_out=ld)
        xin = torch.randn(B, cin, H, W, device=dev).contiguous(memory_format=torch.channels_last)
        wcl = w.contiguous(memory_format=torch.channels_last)
        ref = lambda: F.conv2d(xin, wcl, None, padding=1)
        gf = 2 * 9 * cin * cout * B * H * W / 1e9
        t, tf, tc = timeit(fn), timeit(fn, do_flush=True), timeit(ref)
        print(f"{name:26s} {gf:7.1f} GF  srk {t:7.1f} us ({gf / t * 1e3:6.1f} TF/s)  flushed {tf:7.1f} us   cuDNN tf32 {tc:7.1f} us")
    # the fp32 rows -> fp16 NHWC conversion in front of the body convolutions
    x = torch.randn(16 * 4096, 180, device=dev)
    o = torch.empty(16 * 4096, 192, dtype=torch.float16, device=dev)
    t = timeit(lambda: L.rows_to_f16(x, o, channels=180, ld_in=180, pixels=16 * 4096))
    print(f"rows_to_f16 65536 x 180: {t:.1f} us ({(x.numel() * 4 + o.numel() * 2) / t / 1e3:.0f} GB/s)")
