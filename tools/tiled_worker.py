#!/usr/bin/env python
"""torchrun worker: TiledSuperResolver over WORLD_SIZE real GPUs (NCCL), gathered image on rank 0 compared bit for bit with
the single-GPU result computed on rank 0.  Used by tests/test_gpu_full_configs.py (2 GPUs) and by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/tiled_worker.py --height 200 --width 136
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=200)
    ap.add_argument("--width", type=int, default=136)
    ap.add_argument("--config", default="swinir_x4_d2")
    ap.add_argument("--out-dtype", default="float32", choices=["float32", "bfloat16", "uint8"])
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import tpu_superresolution_b200 as srk
    from tpu_superresolution_b200 import tiling, synth

    cfg = synth.CONFIGS[args.config]
    model = srk.SwinIR(**cfg.as_kwargs()).eval()
    model.load_state_dict(synth.make_swinir_state_dict(cfg, seed=1234, kind="init"), strict=True)
    model.cuda()
    torch.backends.cudnn.allow_tf32 = True
    lr = synth.make_lr_batch(1, args.height, args.width, seed=9)            # host image: each rank copies only its rows
    res = tiling.TiledSuperResolver(model, scale=cfg.upscale, tile=64, overlap=8, batch=16, out_dtype=getattr(torch, args.out_dtype))
    with torch.no_grad():
        out = res(lr.pin_memory(), rank=rank, world=world, device=torch.device("cuda", local))
        torch.cuda.synchronize()
        if rank == 0:
            single = res(lr.cuda(), rank=0, world=1)
            torch.cuda.synchronize()
            assert out.shape == single.shape == (1, 3, args.height * cfg.upscale, args.width * cfg.upscale), out.shape
            assert torch.equal(out, single), f"sharded != single: max diff {(out.float() - single.float()).abs().max().item():.3e}"
            print(f"TILED_WORKER_OK world={world} shape={tuple(out.shape)} dtype={out.dtype} "
                  f"seam_recompute={tiling.seam_recompute_fraction(args.height, args.width, 64, 8, world):.3f}")
        else:
            assert out is None
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
