#!/usr/bin/env python
"""Benchmark of the fused window-attention SR path (BASELINE.json: "SwinIR/HAT x4 output Mpix/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

One "step" = one pass of the hot path over one batch of synthetic input: by default BASELINE.json configs[1],
SwinIR classical x4 on a batch of 16 LR tiles of 64x64 per GPU (16 x 3 x 256 x 256 = 1.049 output Mpix);
`--workload hat_x4` / `dat_x2` time configs[2] (HAT x4, batch 8) and configs[3] (DAT x2, batch 16) the same way.
With N > 1 (torchrun, one rank per GPU) every rank runs its own batch -- tiles are independent, there is
no collective on the data path -- so scaling is "weak" and `value` is the sum over ranks divided by the
slowest rank's time.

configs[4] -- tiled x4 SwinIR inference of a synthetic 4096x4096 LR image (5 329 overlapping 64-px tiles) sharded over
the ranks, timed from LR-resident to the stitched image resident on rank 0, the NCCL gather of the bands included -- is
measured (a) as the `tiled_4096` object of every default line (a few steps, so SCALE_rNN.json carries its strong
scaling at N = 1/2/4/8 beside the replica numbers) and (b) as the main metric with `--workload tiled_4096`
(`scaling: "strong"`, value = unique output Mpix/s).

Reference arms: `--impl reference` = the UNMODIFIED reference modules (staged into oracle/_ref by build(); else the
oracle port) on the host cores, full step; the default line also carries `reference_gpu` = the same unmodified modules'
eager CUDA forward (fp32 and bf16 autocast, finetune_swinir.py:194) on the same GPU.  JSON keys follow the driver contract.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

TILE = 64
# algorithmic FLOPs, 2/MAC, un-padded dims.  "kernels": libsrk launch name -> FLOP per token per launch.
WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on (default)
    "swinir_x4": dict(
        metric="SwinIR x4 output Mpix/s", family="swinir", cfg="swinir_x4", tiles=16, scale=4, gflop_per_tile=107.113,
        workload="SwinIR classical x4 (embed 180, 6 RSTB x 6, window 8, 6 heads, mlp_ratio 2), batch 16 of 64x64 LR tiles per GPU",
        dominant="swin_attn", kernel_name="swin_attn_kernel",
        kernels={"swin_attn": 305_280,      # qkv + qk^T + pv + proj (one swin_attn_kernel launch covers all tokens of the batch)
                 "swin_mlp": 259_200,       # fc1 + fc2
                 "conv3x3_c180": 583_200}), # 3x3 conv 180 -> 180 as an implicit GEMM
    # BASELINE.json configs[2]: HAT x4, window 16, OCAB overlap 0.5, CAB, batch 8
    "hat_x4": dict(
        metric="HAT x4 output Mpix/s", family="hat", cfg="hat_x4", tiles=8, scale=4, gflop_per_tile=207.761,
        workload="HAT x4 (embed 180, 6 RHAG x (6 HAB + OCAB), window 16, overlap 0.5, CAB 3/30, mlp_ratio 2), batch 8 of 64x64 LR tiles per GPU",
        dominant="window_attention", kernel_name="winattn_kernel",
        kernels={"window_attention": (36 * 184_320 + 6 * 414_720) / 42,    # qk^T + pv: 256 keys (36 W-MSA) / 576 keys (6 OCAB) per launch
                 "linear": (42 * 194_400 + 42 * 64_800) / 84,              # qkv and proj launches
                 "swin_mlp": 259_200}),
    # BASELINE.json configs[3]: DAT x2, split 8x32, expansion 4, batch 16
    "dat_x2": dict(
        metric="DAT x2 output Mpix/s", family="dat", cfg="dat_x2", tiles=16, scale=2, gflop_per_tile=131.635,
        workload="DAT x2 (embed 180, 6 RG x 6 DATB, split 8x32, 6 heads, expansion 4), batch 16 of 64x64 LR tiles per GPU",
        dominant="window_attention", kernel_name="winattn_kernel",
        kernels={"window_attention": 92_160,       # qk^T + pv of the 3 heads of one channel half (two launches per spatial block)
                 # per DATB: spatial qkv planes 194400 + v rows 64800 (or channel qkv 194400) + proj 64800 + fc1 259200 + fc2 2 x 64800
                 "linear": (18 * (194_400 + 64_800) + 18 * 194_400 + 36 * (64_800 + 259_200 + 129_600)) / (18 * 2 + 18 + 36 * 4)}),
}
# BASELINE.json configs[4]: same model as swinir_x4
TILED = dict(height=4096, width=4096, overlap=8, batch=16, tiles=5329, unique_mpix=268.435456,
             workload="tiled x4 SwinIR inference of a synthetic 4096x4096 LR image: 5329 overlapping 64-px tiles (overlap 8), HR row bands sharded over the GPUs, "
                      "seam tiles recomputed, one gather of the stitched bands to GPU 0")
WORKLOADS["tiled_4096"] = dict(WORKLOADS["swinir_x4"], metric="tiled SwinIR x4 unique output Mpix/s (4096x4096 LR)", workload=TILED["workload"])
W = WORKLOADS["swinir_x4"]      # replaced in main()
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def shared_config():
    """The `config` object of BOTH arms (identical, so the driver can pair them)."""
    return {"workload": W["workload"], "tiles_per_step_per_gpu": W["tiles"], "input": "synthetic fp32 LR tiles in [0,1), numpy RNG",
            "l2": "GPU arm: flushed between steps (256 MiB memset outside the timed events); CPU arm: not applicable"}


def _build(family, cfg_name):
    """-> (cfg, state_dict, drop-in model class) for a workload (synthetic weights, numpy RNG)."""
    from tpu_superresolution_b200 import synth
    import tpu_superresolution_b200 as srk
    if family == "swinir":
        cfg = synth.CONFIGS[cfg_name]
        return cfg, synth.make_swinir_state_dict(cfg, seed=1234, kind="init"), srk.SwinIR
    if family == "hat":
        cfg = synth.HAT_CONFIGS[cfg_name]
        return cfg, synth.make_hat_state_dict(cfg, seed=1234, kind="init"), srk.HAT
    cfg = synth.DAT_CONFIGS[cfg_name]
    return cfg, synth.make_dat_state_dict(cfg, seed=1234, kind="init"), srk.DAT


def reference_forward(family, cfg, sd, device="cpu"):
    """-> (callable lr -> sr, kind): the UNMODIFIED reference module (oracle/_ref, staged by build()) when present -- kind
    "reference" -- else the oracle restatement (CPU only) -- kind "port".  The only place bench.py executes oracle/."""
    have_ref = os.path.isfile(os.path.join(REF_DIR, "modules", "network_swinir.py")) or os.path.isdir("/root/reference/modules")
    if have_ref:
        if not os.path.isdir("/root/reference/modules"):
            os.environ["SRK_REFERENCE_ROOT"] = REF_DIR
        from oracle.reference_loader import load_reference_module
        modname, clsname = {"swinir": ("network_swinir", "SwinIR"), "hat": ("hat_arch", "HAT"), "dat": ("dat_arch", "DAT")}[family]
        model = getattr(load_reference_module(modname), clsname)(**cfg.as_kwargs()).eval()
        model.load_state_dict(sd, strict=True)
        model.to(device)
        return (lambda lr: model(lr)), "reference"
    if device != "cpu":
        return None, "unavailable"
    if family == "swinir":
        from oracle import swinir_oracle as O
        return (lambda lr: O.swinir_forward(lr, sd, cfg)), "port"
    if family == "hat":
        from oracle import hat_oracle as HO
        return (lambda lr: HO.hat_forward(lr, sd, cfg)), "port"
    from oracle import dat_oracle as DO
    return (lambda lr: DO.dat_forward(lr, sd, cfg)), "port"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained bf16, 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


def _dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: whatever NCCL_DEBUG level is in force, its log (incl. the version banner) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if "SRK_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["SRK_NCCL_DEBUG"]
        torch.cuda.set_device(local)
        # NCCL prints its version banner on stdout when the communicator is created: create it (first collective) with fd 1
        # pointed at stderr so that stdout carries nothing but the JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    return world, rank, local


def _max_over_ranks(x: float, world: int) -> float:
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(world: int):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def cpu_reference_rate(tiles: int, reps: int = 1, warm: bool = True):
    """The reference's CPU forward (unmodified modules if staged, else the oracle port; fp32) on the host cores:
    output Mpix/s on `tiles` tiles of the workload."""
    from tpu_superresolution_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, sd, _ = _build(W["family"], W["cfg"])
    fwd, kind = reference_forward(W["family"], cfg, sd)
    lr = synth.make_lr_batch(tiles, TILE, TILE, seed=2)
    with torch.no_grad():
        if warm:
            fwd(lr[:1])
        best = float("inf")
        for _ in range(reps):
            t0 = time.perf_counter()
            fwd(lr)
            best = min(best, time.perf_counter() - t0)
    return tiles * (TILE * W['scale']) ** 2 / best / 1e6, cores, best, kind


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the unmodified modules staged into oracle/_ref by
    build(); the oracle port if they are absent) on all host cores, the FULL step (all tiles of the batch) per timed step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tpu_superresolution_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tiles = W["tiles"]
    cfg, sd, _ = _build(W["family"], W["cfg"])
    fwd, kind = reference_forward(W["family"], cfg, sd)
    lr = synth.make_lr_batch(tiles, TILE, TILE, seed=2)
    with torch.no_grad():
        for _ in range(args.warmup):
            fwd(lr)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fwd(lr)
        dt = time.perf_counter() - t0
    mpix = tiles * (TILE * W['scale']) ** 2 / 1e6
    value = mpix * args.steps / dt
    sample = f"all {tiles} tiles of one step per timed step, fp32, {'unmodified reference modules (oracle/_ref)' if kind == 'reference' else 'oracle port'}, torch CPU ops, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": W["metric"], "value": value, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": shared_config(),
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def reference_gpu_rates(dev, lr_dev, steps):
    """The unmodified reference's EAGER CUDA forward on this GPU (what the reference does today: ~850 launches per SwinIR
    forward), fp32 and under bf16 autocast (finetune_swinir.py:194), same weights and inputs, CUDA events."""
    cfg, sd, _ = _build(W["family"], W["cfg"])
    fwd, kind = reference_forward(W["family"], cfg, sd, device=dev)
    if fwd is None:
        return {"unavailable": "oracle/_ref not staged (build() runs where /root/reference exists)"}
    out = {"kind": "unmodified reference modules (oracle/_ref), eager PyTorch CUDA forward, same weights / inputs / batch", "steps": steps}
    mpix = lr_dev.shape[0] * (TILE * W['scale']) ** 2 / 1e6
    stream = torch.cuda.current_stream()
    for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        with torch.no_grad(), ctx:
            for _ in range(2):
                fwd(lr_dev)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                fwd(lr_dev)
            e1.record(stream)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"value": mpix / (ms * 1e-3), "unit": "Mpix/s", "ms_per_step": ms}
    return out


def run_tiled(model, dev, world, rank, steps, warmup):
    """BASELINE.json configs[4]: strong scaling of the tiled 4096x4096 image over `world` ranks (tiling.TiledSuperResolver).
    value: LR resident on every rank's device -> stitched fp32 image resident on rank 0, gather included (CUDA events, max over
    ranks).  e2e: pinned host uint8 LR -> each rank copies its rows -> uint8 stitched image in pinned host memory on rank 0."""
    from tpu_superresolution_b200 import synth, tiling
    H, Wd = TILED["height"], TILED["width"]
    lr = synth.make_lr_batch(1, H, Wd, seed=0)
    lr_dev = lr.to(dev)
    res = tiling.TiledSuperResolver(model, scale=4, tile=TILE, overlap=TILED["overlap"], batch=TILED["batch"])
    res8 = tiling.TiledSuperResolver(model, scale=4, tile=TILE, overlap=TILED["overlap"], batch=TILED["batch"], out_dtype=torch.uint8)
    res8.run_tiles = res.run_tiles                      # share the captured graph
    stream = torch.cuda.current_stream()
    plan = res.plan(H, Wd, rank, world, dev)

    def timed(fn, n):
        ts = []
        for _ in range(n):
            _barrier(world)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            _barrier(world)
            ts.append(_max_over_ranks(e0.elapsed_time(e1) * 1e-3, world))
        return ts

    with torch.no_grad():
        for _ in range(warmup):
            out = res(lr_dev, rank=rank, world=world)
            del out
        t_full = timed(lambda: res(lr_dev, rank=rank, world=world), steps)
        t_band = timed(lambda: res(lr_dev, rank=rank, world=world, gather=False), max(1, min(steps, 2)))      # no gather: its cost by difference
        # end to end: host uint8 LR -> uint8 HR image in pinned host memory on rank 0
        lr8 = (lr * 255).round().to(torch.uint8).pin_memory()
        host_out = torch.empty(1, 3, H * 4, Wd * 4, dtype=torch.uint8).pin_memory() if rank == 0 else None

        def e2e_step():
            o = res8(lr8, rank=rank, world=world, device=dev)
            if rank == 0:
                host_out.copy_(o, non_blocking=True)
        e2e_step()
        t_e2e = timed(e2e_step, max(1, min(steps, 2)))
    t = sum(t_full) / len(t_full)
    tb = sum(t_band) / len(t_band)
    te = sum(t_e2e) / len(t_e2e)
    gather_bytes = 3 * (H * 4) * (Wd * 4) * 4 * (world - 1) / world
    band_rows = plan.band[1] - plan.band[0]
    return {
        "workload": TILED["workload"], "scaling": "strong", "n_gpus": world, "steps": steps, "warmup": warmup,
        "value": TILED["unique_mpix"] / t, "unit": "unique output Mpix/s", "ms_per_step": t * 1e3,
        "tiles_total": TILED["tiles"], "tiles_rank0": plan.n_tiles, "lr_rows_rank0": band_rows,
        "seam_recompute_fraction": tiling.seam_recompute_fraction(H, Wd, TILE, TILED["overlap"], world),
        "compute_and_stitch_ms": tb * 1e3, "gather_ms": max(0.0, (t - tb)) * 1e3,
        "gather_bytes_into_rank0": gather_bytes, "gather_gbs": (gather_bytes / max(t - tb, 1e-9) / 1e9) if world > 1 else None,
        "frac_of_sustained_bf16_peak": TILED["tiles"] * W["gflop_per_tile"] / t / 1e3 / _peaks()[0] / world,
        "e2e": {"value": TILED["unique_mpix"] / te, "unit": "unique output Mpix/s", "ms_per_step": te * 1e3,
                "h2d_bytes_per_step": 3 * H * Wd, "d2h_bytes_per_step": 3 * H * 4 * Wd * 4,
                "mode": "pinned host uint8 LR -> per-rank H2D of its rows -> uint8 stitched image -> pinned host on rank 0"},
        "timed": "CUDA events per step on every rank, barrier + synchronize both sides, max over ranks; LR resident, output resident on rank 0",
    }


def run_ours(args):
    from tpu_superresolution_b200 import synth    # synthetic weights / inputs (numpy RNG)
    from tpu_superresolution_b200 import _lib as L

    world, rank, local = _dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    L.load()
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True

    TILES_PER_STEP = W["tiles"]
    cfg, sd, model_cls = _build(W["family"], W["cfg"])
    model = model_cls(**cfg.as_kwargs()).eval()
    model.load_state_dict(sd, strict=True)
    model.to(dev)

    if args.workload == "tiled_4096":                   # configs[4] as the main metric
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = L.launch_count()
        tiled = run_tiled(model, dev, world, rank, args.steps, args.warmup)
        launches = L.launch_count() - l0
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            print(json.dumps({
                "metric": W["metric"], "value": tiled["value"], "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": tiled["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": {"workload": TILED["workload"], "tiles_per_step": TILED["tiles"],
                                                "l2": "inputs larger than L2 (201 MB LR, 3.2 GB output per step)"},
                "e2e": tiled["e2e"], "gpu_launches": int(launches), "clocks": clocks, "tiled_4096": tiled,
                "roofline": {"kernel": "whole step", "bound": "tensor", "achieved": tiled["frac_of_sustained_bf16_peak"] * _peaks()[0],
                             "peak": _peaks()[0], "unit": "TFLOP/s", "frac": tiled["frac_of_sustained_bf16_peak"], "traffic": None},
                "cpu_baseline": None}))
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    # one cudaGraphLaunch per step instead of ~1000 kernel launches: at this batch size the eager forward is bound by the
    # host's launch rate (tpu_superresolution_b200/graphs.py); --eager times the plain module instead
    from tpu_superresolution_b200.graphs import GraphedModel
    runner = model if args.eager else GraphedModel(model)
    n_in = 4                                        # rotate distinct input batches
    host_in = [synth.make_lr_batch(TILES_PER_STEP, TILE, TILE, seed=100 + rank * n_in + i).pin_memory() for i in range(n_in)]
    dev_in = [h.to(dev) for h in host_in]
    host_out = torch.empty(TILES_PER_STEP, 3, TILE * W['scale'], TILE * W['scale']).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    mpix_step = TILES_PER_STEP * (TILE * W['scale']) ** 2 / 1e6
    stream = torch.cuda.current_stream()

    def step_resident(i):
        return runner(dev_in[i % n_in])

    def step_e2e(i):
        if args.eager:
            y = runner(host_in[i % n_in].to(dev, non_blocking=True))
        else:
            y = runner(host_in[i % n_in])          # H2D straight into the graph's static input buffer
        host_out.copy_(y, non_blocking=True)
        return y

    def timed(fn, steps, warmup):
        with torch.no_grad():
            for i in range(warmup):
                fn(i)
            _barrier(world)
            evs = []
            for i in range(steps):
                flush.zero_()                                          # L2 flush, outside the timed events
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn(i)
                e1.record(stream)
                evs.append((e0, e1))
            _barrier(world)
            return sum(a.elapsed_time(b) for a, b in evs) / 1e3        # seconds of device time in the K steps

    with torch.no_grad():                       # a benchmark of a wrong result is worthless: check the output first
        y_chk = step_resident(0)
        torch.cuda.synchronize()
        if not bool(torch.isfinite(y_chk).all()) or not (0.0 < float(y_chk.mean()) < 1.0):
            raise SystemExit("bench.py: model output is not finite / out of range; refusing to time it")
        del y_chk

    with torch.no_grad():                       # libsrk launches of ONE forward (a graph replays exactly these)
        l0 = L.launch_count()
        model(dev_in[0])
        torch.cuda.synchronize()
        launches_per_step = L.launch_count() - l0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_res = _max_over_ranks(timed(step_resident, args.steps, args.warmup), world)
    launches = launches_per_step * args.steps
    if args.eager:
        t_e2e = _max_over_ranks(timed(step_e2e, args.steps, max(1, args.warmup // 2)), world)
        e2e_mode = "per step: H2D, forward, D2H on one stream"
    else:
        # host-to-host serving loop (srk.PipelinedRunner): H2D into the graph's input buffer, replay, D2H on a side stream, so the
        # 12.6 MB result copy of step i overlaps the forward of step i + 1.  ONE timed region around all K steps, closed only after
        # the last copy has landed; the L2 flushes between the steps are inside it (they cannot be cut out of a pipelined region).
        from tpu_superresolution_b200.graphs import PipelinedRunner
        host_outs = [host_out, torch.empty_like(host_out).pin_memory()]

        def timed_pipeline(steps, warmup):
            pr = PipelinedRunner(runner)
            with torch.no_grad():
                for i in range(warmup):
                    flush.zero_()
                    pr.submit(host_in[i % n_in], host_outs[i % 2])
                pr.drain()
                torch.cuda.synchronize()
                _barrier(world)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for i in range(steps):
                    flush.zero_()
                    pr.submit(host_in[i % n_in], host_outs[i % 2])
                pr.drain()
                e1.record(stream)
                _barrier(world)
                return e0.elapsed_time(e1) / 1e3

        t_e2e = _max_over_ranks(timed_pipeline(args.steps, max(2, args.warmup // 2)), world)
        e2e_mode = "pipelined host-to-host loop (PipelinedRunner): D2H of step i overlaps step i+1; one timed region over all steps, L2 flushes inside it"
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel timing pass (CUDA events around every libsrk launch) for the roofline of the dominant kernel
    prof = {}
    L.PROFILE = prof
    with torch.no_grad():
        for i in range(3):
            flush.zero_()
            model(dev_in[i % n_in])             # eager: events cannot be recorded around kernels inside a graph
    torch.cuda.synchronize()
    L.PROFILE = None
    kstats = {k: (sum(a.elapsed_time(b) for a, b in v) / len(v), len(v)) for k, v in prof.items()}

    # ---- the dominant kernel launched back to back (SwinIR only): 40 launches of one block's attention half in a row on the residual
    #      stream, programmatic dependent launch as inside the graph, one pair of events around all of them
    b2b = None
    if W["family"] == "swinir" and rank == 0:
        blk0, blk1 = model.layers[0].residual_group.blocks[0], model.layers[0].residual_group.blocks[1]
        xt = synth.make_tokens(TILES_PER_STEP, TILE, TILE, 180, seed=7).to(dev)
        b2b = {}
        with torch.no_grad():
            for blk, name in ((blk0, "shift0"), (blk1, "shift4")):
                aw, av = blk.attn._packed(blk.norm1)

                def one():
                    L.swin_attn(xt, xt, aw, av, mode=L.MODE_IMAGE, batch=TILES_PER_STEP, height=TILE, width=TILE, ld_in=180, ld_out=180,
                                shift=blk.shift_size, mask_mode=L.MASK_SHIFT if blk.shift_size else L.MASK_NONE, operands=blk.attn.operands)
                for _ in range(6):
                    one()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(40):
                    one()
                e1.record(stream)
                torch.cuda.synchronize()
                b2b[name] = e0.elapsed_time(e1) / 40
        del xt

    # ---- BASELINE.json configs[4] beside the replica number (SwinIR x4 only): a few steps of the tiled 4096x4096 image
    tiled = None
    if W["cfg"] == "swinir_x4" and not args.no_tiled:
        flush = None
        tiled = run_tiled(model, dev, world, rank, steps=max(1, min(args.steps, 3)), warmup=1)

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    peak_tf, peak_gbs, peak_src = _peaks()
    tokens = TILES_PER_STEP * TILE * TILE

    def kstat(name):
        ms, n = kstats.get(name, (float("nan"), 0))
        tf = W["kernels"][name] * tokens / (ms * 1e-3) / 1e12
        return {"avg_launch_ms": ms, "launches_timed": n, "achieved": tf, "frac": tf / peak_tf,
                "algorithmic_flop_per_launch": W["kernels"][name] * tokens}

    dom = kstat(W["dominant"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(W["kernel_name"] + "_dram_bytes_per_launch")

    cpu_baseline, ref_gpu = None, None          # reported at N = 1 only (driver contract)
    if world == 1:
        cpu_tiles = TILES_PER_STEP if (os.cpu_count() or 1) >= 16 else 4
        cpu_val, cores, cpu_s, kind = cpu_reference_rate(tiles=cpu_tiles)
        cpu_baseline = {"value": cpu_val, "unit": "Mpix/s", "cores": cores, "kind": kind,
                        "sample": f"{cpu_tiles} of the {TILES_PER_STEP} tiles of one step, fp32, "
                                  f"{'unmodified reference modules (oracle/_ref)' if kind == 'reference' else 'oracle port'} (torch CPU ops), {cpu_s:.1f} s"}
        ref_gpu = reference_gpu_rates(dev, dev_in[0], steps=max(2, min(args.steps, 10)))

    tight = None
    if world == 1 and hasattr(model, "set_precision"):
        # the tight precision mode (north star: "tighter for a TF32 mode"): fp16 operands in the fused attention / MLP kernels,
        # hi / lo split convolutions on the same tcgen05 kernel; eager launches, inputs resident, CUDA events
        model.set_precision("fp16")
        with torch.no_grad():
            for i in range(2):
                model(dev_in[i % n_in])
            n_t = max(2, min(args.steps, 5))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n_t):
                model(dev_in[i % n_in])
            e1.record()
            torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / n_t
        tight = {"value": mpix_step / (ms_t * 1e-3), "unit": "Mpix/s", "ms_per_step": ms_t, "steps": n_t,
                 "precision": "fp16 MMA operands (11-bit significand) in the fused attention / MLP kernels, fp32 accumulate; 3x3 convolutions as hi/lo fp16 pairs on the tcgen05 kernel (3 products per layer, ~22-bit operands)",
                 "gate": "max abs <= 2e-4 vs the reference fp32 forward (tests/test_gpu_full_configs.py::test_tight_mode_fp16_operands_vs_reference_golden)"}
        model.set_precision("bf16")

    others = None
    if world == 1 and args.workload == "swinir_x4":
        # the metric names HAT x4 too (BASELINE.json): the other two families of the path at their BASELINE configs in the same run --
        # CUDA-graph replay, inputs resident and rotated, L2 flushed between steps, CUDA events (full lines: --workload hat_x4 | dat_x2)
        others = {}
        fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev)         # L2 flush buffer (the main one was freed for the tiled run)
        for name in ("hat_x4", "dat_x2"):
            Wo = WORKLOADS[name]
            cfg_o, sd_o, cls_o = _build(Wo["family"], Wo["cfg"])
            mo = cls_o(**cfg_o.as_kwargs()).eval()
            mo.load_state_dict(sd_o, strict=True)
            mo.to(dev)
            go = GraphedModel(mo)
            xin = [synth.make_lr_batch(Wo["tiles"], TILE, TILE, seed=300 + i).to(dev) for i in range(2)]
            n_o = max(3, min(args.steps, 10))
            with torch.no_grad():
                yo = go(xin[0])
                torch.cuda.synchronize()
                if not bool(torch.isfinite(yo).all()) or not (0.0 < float(yo.mean()) < 1.0):
                    raise SystemExit(f"bench.py: {name} output is not finite / out of range")
                for i in range(3):
                    go(xin[i % 2])
                tot = 0.0
                for i in range(n_o):
                    fl.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    go(xin[i % 2])
                    e1.record(stream)
                    torch.cuda.synchronize()
                    tot += e0.elapsed_time(e1)
            ms_o = tot / n_o
            mp_o = Wo["tiles"] * (TILE * Wo["scale"]) ** 2 / 1e6
            tf_o = Wo["gflop_per_tile"] * Wo["tiles"] / ms_o
            others[name] = {"metric": Wo["metric"], "value": mp_o / (ms_o * 1e-3), "unit": "Mpix/s", "ms_per_step": ms_o, "steps": n_o,
                            "workload": Wo["workload"], "whole_model_frac_of_peak": tf_o / peak_tf}
            del go, mo, xin

    ms_step = t_res / args.steps * 1e3
    value = world * mpix_step * args.steps / t_res
    model_tf = world * W["gflop_per_tile"] * TILES_PER_STEP * args.steps / t_res / 1e3
    out = {
        "metric": W["metric"], "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": shared_config(),
        "detail": {"parallelism": f"tile-sharded x{world}, no collective",
                   "precision": "bf16 MMA operands in the attention kernels, fp16 in the MLP and the convolutions; fp32 accumulate, fp32 residual stream / LayerNorm / softmax",
                   "launch": "eager" if args.eager else "one CUDA graph replay per step",
                   "whole_model_tflops": model_tf, "whole_model_frac_of_peak": model_tf / world / peak_tf},
        "e2e": {"value": world * mpix_step * args.steps / t_e2e, "unit": "Mpix/s",
                "h2d_bytes_per_step": host_in[0].numel() * 4, "d2h_bytes_per_step": host_out.numel() * 4,
                "ms_per_step": t_e2e / args.steps * 1e3, "mode": e2e_mode},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": W["kernel_name"], "bound": "tensor", "achieved": dom["achieved"], "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": dom["frac"], "traffic": traffic, "peak_source": peak_src,
                     "avg_launch_ms": dom["avg_launch_ms"], "launches_timed": dom["launches_timed"],
                     "avg_launch_ms_note": "eager pass with CUDA events around every launch: no programmatic-dependent-launch / progress-counter "
                                           "overlap between kernels, so this is a conservative (upper) launch time",
                     "algorithmic_flop_per_launch": dom["algorithmic_flop_per_launch"],
                     "back_to_back": None if not b2b else {
                         "avg_launch_ms": sum(b2b.values()) / len(b2b), "per_shift_ms": b2b,
                         "frac": dom["algorithmic_flop_per_launch"] / (sum(b2b.values()) / len(b2b) * 1e-3) / 1e12 / peak_tf,
                         "note": "the same kernel launched 40 times in a row on the L2-resident residual stream (programmatic dependent launch "
                                 "between the launches, as inside the graph replay): its steady cost inside the step; `frac` above stays the "
                                 "conservative isolated-launch figure"},
                     "other_kernels": {k: kstat(k) for k in W["kernels"] if k != W["dominant"] and k in kstats},
                     "libsrk_ms_per_step": sum(v[0] * v[1] for v in kstats.values()) / 3.0,
                     "per_kernel_ms_per_step": {k: v[0] * v[1] / 3.0 for k, v in sorted(kstats.items())}},
        "cpu_baseline": cpu_baseline,
        "reference_gpu": ref_gpu,
        "tight_mode": tight,
        "other_workloads": others,
        "tiled_4096": tiled,
    }
    print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="swinir_x4", choices=sorted(WORKLOADS))
    ap.add_argument("--eager", action="store_true", help="time the plain module instead of the CUDA-graph replay")
    ap.add_argument("--no-tiled", action="store_true", help="skip the tiled_4096 object of the default line")
    args = ap.parse_args()
    global W
    W = WORKLOADS[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
