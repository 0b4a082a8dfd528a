"""GPU parity: the CUDA path (through the C ABI / drop-in modules) against the CPU oracle and the
reference-made golden fixtures.  Tolerances: the kernels use bf16 MMA operands with fp32 accumulation
and an fp32 residual stream; BASELINE.json's gate for the bf16 path is max-abs <= 2e-3 on [0,1] pixels
and |dPSNR| <= 0.01 dB for whole-model outputs.  Module-level KATs use a relative gate (bf16 operand
rounding, ~2^-8 relative per product) stated per test.
"""
import os

import numpy as np
import pytest
import torch

import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import _lib as L
from oracle import synth
from oracle import swinir_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    L.load()
    torch.backends.cudnn.allow_tf32 = True
    with torch.no_grad():
        yield


def _rel(a, b):
    a, b = a.double().cpu(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _block_sd(sd, pre):
    return {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}


def _stress_sd():
    return synth.make_swinir_state_dict(synth.CONFIGS["swinir_x2_d2"], seed=99, kind="stress")


def test_layernorm_kernel():
    x = synth.make_tokens(3, 5, 7, 180, seed=1).cuda()
    w = torch.rand(180, device="cuda") + 0.5
    b = torch.randn(180, device="cuda")
    y = torch.empty_like(x)
    L.layernorm(x, y, w, b, num_tokens=105, ld_in=180, ld_out=180)
    ref = O.layer_norm(x.cpu(), w.cpu(), b.cpu())
    assert (y.cpu() - ref).abs().max() < 1e-5


@pytest.mark.parametrize("r,shape", [(2, (2, 16, 5, 7)), (2, (1, 256, 9, 4)), (3, (1, 27, 4, 5))])
def test_pixelshuffle_kernel_bit_exact(r, shape):
    x = torch.randn(shape, device="cuda").contiguous(memory_format=torch.channels_last)
    y = srk.PixelShuffle(r)(x)
    assert torch.equal(y.cpu(), O.pixel_shuffle(x.cpu(), r))
    g = np.load(os.path.join(GOLDEN, "kat_tail.npz"))
    ps_in = torch.from_numpy(np.random.default_rng(33).normal(0, 1, size=(2, 16, 5, 7)).astype(np.float32)).cuda()
    assert np.array_equal(srk.PixelShuffle(2)(ps_in).cpu().numpy(), g["ps"])


@pytest.mark.parametrize("r,shape", [(2, (2, 16, 5, 7)), (2, (1, 256, 9, 4)), (3, (1, 27, 4, 5))])
def test_pixelshuffle_with_conv_bias_bit_exact(r, shape):
    # Upsample runs its conv bias-free and the shuffle adds the bias on the way: same fp32 add, moved
    x = torch.randn(shape, device="cuda").contiguous(memory_format=torch.channels_last)
    bias = torch.randn(shape[1], device="cuda")
    y = srk.PixelShuffle(r)(x, bias=bias)
    ref = O.pixel_shuffle((x + bias.view(1, -1, 1, 1)).cpu(), r)
    assert torch.equal(y.cpu(), ref)


@pytest.mark.parametrize("pixels,C", [(1, 4), (37, 180), (16 * 64 * 64, 180), (513, 3), (1000, 64)])
@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("use_bias,use_res", [(True, True), (True, False), (False, True)])
def test_bias_act_add_kernel_bit_exact(pixels, C, act, use_bias, use_res):
    from tpu_superresolution_b200 import _lib as L
    g = torch.Generator(device="cpu").manual_seed(pixels * 7 + C)
    x = torch.randn(pixels, C, generator=g).cuda()
    bias = torch.randn(C, generator=g).cuda() if use_bias else None
    res = torch.randn(pixels, C, generator=g).cuda() if use_res else None
    ref = x.clone()
    if bias is not None:
        ref = ref + bias
    if act == 1:
        ref = torch.where(ref > 0, ref, ref * 0.2)
    elif act == 2:
        ref = torch.nn.functional.gelu(ref)               # exact (erf) GELU; the libm erff may differ in the last ulp
    if res is not None:
        ref = ref + res
    same = torch.equal if act != 2 else (lambda a, b: torch.allclose(a, b, rtol=0, atol=2e-6))
    y = torch.empty_like(x)
    L.bias_act_add_nhwc(x, y, pixels=pixels, channels=C, bias=bias, residual=res, act=act, slope=0.2)
    assert same(y, ref)
    L.bias_act_add_nhwc(x, x, pixels=pixels, channels=C, bias=bias, residual=res, act=act, slope=0.2)      # in place
    assert same(x, ref)


@pytest.mark.parametrize("ntok", [64, 128, 1000, 128 * 149 + 5])
def test_mlp_kernel(ntok):
    sd = _stress_sd()
    pre = "layers.0.residual_group.blocks.0.mlp."
    m = srk.Mlp(180, 360).eval()
    m.load_state_dict(_block_sd(sd, pre), strict=True)
    m.cuda()
    x = synth.make_tokens(1, 1, ntok, 180, seed=7)[0]
    y = m(x.cuda())
    ref = O.mlp(x, sd, pre)
    assert _rel(y, ref) < 1.5e-2          # bf16 operands on two chained GEMMs


def test_window_attention_kat_vs_golden():
    """network_swinir.py:114-145 with stress weights (peaky softmax), with and without an explicit mask."""
    g = np.load(os.path.join(GOLDEN, "kat_window_attention.npz"))
    sd = _stress_sd()
    pre = "layers.0.residual_group.blocks.1.attn."
    attn = srk.WindowAttention(180, (8, 8), 6).eval()
    attn.load_state_dict(_block_sd(sd, pre), strict=True)
    attn.cuda()
    xw = synth.make_tokens(8, 8, 8, 180, seed=5).cuda()
    assert _rel(attn(xw), g["y_nomask"]) < 2e-2
    assert _rel(attn(xw, torch.from_numpy(g["mask"]).cuda()), g["y_mask"]) < 2e-2
    assert _rel(attn(xw[:7]), g["y_nomask"][:7]) < 2e-2          # odd window count -> half-empty last tile


@pytest.mark.parametrize("tag,b_idx,shift,res,x_size", [
    ("unshifted", 0, 0, (16, 24), (16, 24)), ("shifted", 1, 4, (16, 24), (16, 24)),
    ("shifted_nonnative", 1, 4, (16, 24), (24, 16)), ("shifted_64", 1, 4, (64, 64), (64, 64))])
def test_block_kat_vs_golden(tag, b_idx, shift, res, x_size):
    """SwinTransformerBlock (network_swinir.py:239-279): roll/partition/mask/reverse as index math."""
    g = np.load(os.path.join(GOLDEN, f"kat_block_{tag}.npz"))
    sd = _stress_sd()
    blk = srk.SwinTransformerBlock(180, res, 6, window_size=8, shift_size=shift, mlp_ratio=2.0).eval()
    st = _block_sd(sd, f"layers.0.residual_group.blocks.{b_idx}.")
    if shift:
        st["attn_mask"] = blk.attn_mask.clone()
    blk.load_state_dict(st, strict=True)
    blk.cuda()
    B = 2 if x_size != (64, 64) else 1
    xt = synth.make_tokens(B, x_size[0], x_size[1], 180, seed=11).cuda()
    y = blk(xt, x_size)
    if x_size == (64, 64):
        y = y[:, ::7]
    assert _rel(y, g["y"]) < 1.5e-2
    # in-place form used by the whole-model fast path gives the same bits
    y2 = blk.forward_into(xt.clone(), x_size, xt.clone())
    y2 = y2[:, ::7] if x_size == (64, 64) else y2
    assert torch.equal(y2, y)


@pytest.mark.parametrize("shift", [0, 4])
def test_block_at_bench_scale_many_tiles_per_cta(shift):
    """B=16 x 64x64 = 512 attention tiles / 512 MLP tiles on 148 persistent CTAs: exercises the cross-tile pipeline
    (next-tile normalisation and GEMMs overlapping the store phase).  Repeated to shake out timing-dependent races."""
    sd = _stress_sd()
    b_idx = 1 if shift else 0
    blk = srk.SwinTransformerBlock(180, (64, 64), 6, window_size=8, shift_size=shift, mlp_ratio=2.0).eval()
    st = _block_sd(sd, f"layers.0.residual_group.blocks.{b_idx}.")
    if shift:
        st["attn_mask"] = blk.attn_mask.clone()
    blk.load_state_dict(st, strict=True)
    blk.cuda()
    xt = synth.make_tokens(16, 64, 64, 180, seed=21)
    ref = O.swin_block(xt[[0, 7, 15]], (64, 64), sd, f"layers.0.residual_group.blocks.{b_idx}.", 6, 8, shift)
    xc = xt.cuda()
    y0 = blk(xc, (64, 64))
    assert torch.isfinite(y0).all()
    assert _rel(y0[[0, 7, 15]], ref) < 1.5e-2
    for _ in range(20):
        assert torch.equal(blk(xc, (64, 64)), y0), "non-deterministic result: race between tiles"


def test_rstb_kat_vs_golden():
    g = np.load(os.path.join(GOLDEN, "kat_rstb.npz"))
    sd = _stress_sd()
    r = srk.RSTB(180, (16, 16), 2, 6, 8, mlp_ratio=2.0, img_size=16, patch_size=1).eval()
    st = _block_sd(sd, "layers.1.")
    st["residual_group.blocks.1.attn_mask"] = r.residual_group.blocks[1].attn_mask.clone()
    r.load_state_dict(st, strict=True)
    r.cuda()
    xt = synth.make_tokens(1, 16, 16, 180, seed=12).cuda()
    assert _rel(r(xt, (16, 16)), g["y"]) < 2e-2


def _model(name, kind, seed):
    cfg = synth.CONFIGS[name]
    sd = synth.make_swinir_state_dict(cfg, seed=seed, kind=kind)
    m = srk.SwinIR(**cfg.as_kwargs()).eval()
    m.load_state_dict(sd, strict=True)
    return m.cuda(), cfg, sd


def _psnr_gate(y, ref, lr, scale):
    hr = torch.nn.functional.interpolate(lr, scale_factor=scale, mode="bicubic", align_corners=False).clamp(0, 1)
    return abs(O.batch_psnr(y, hr).item() - O.batch_psnr(ref, hr).item())


@pytest.mark.parametrize("name,kind,seed,B,h,w", [
    ("swinir_x2", "init", 1234, 1, 64, 64),            # BASELINE.json configs[0]
    ("swinir_x4_d2", "init", 1234, 2, 64, 64),
    ("swinir_x2_d2", "stress", 77, 1, 20, 27),         # reflect pad + crop
    ("swinir_x4_d2", "stress", 4321, 1, 32, 40)])      # non-native x_size
def test_whole_model_vs_golden(name, kind, seed, B, h, w):
    """North-star gate: max-abs <= 2e-3 on [0,1] pixels and |dPSNR| <= 0.01 dB vs the reference fp32 forward."""
    g = np.load(os.path.join(GOLDEN, f"{name}_{kind}_{B}x{h}x{w}.npz"))
    m, cfg, _ = _model(name, kind, seed)
    lr = synth.make_lr_batch(B, h, w, seed=seed + 1)
    y = m(lr.cuda()).cpu()
    ref = torch.from_numpy(g["y"])
    assert y.shape == ref.shape
    tol = 2e-3 if kind == "init" else 2e-3 * max(1.0, ref.abs().max().item())   # stress outputs leave [0,1]
    assert (y - ref).abs().max().item() <= tol
    if kind == "init":
        assert _psnr_gate(y, ref, lr, cfg.upscale) <= 0.01


def test_config2_batch16_x4_vs_oracle_sample():
    """BASELINE.json configs[1] (x4, B=16, 64x64): full size on the GPU; the oracle checks a 2-tile sample,
    the rest through batch independence (each tile's output must not depend on its batch neighbours)."""
    m, cfg, sd = _model("swinir_x4", "init", 1234)
    lr = synth.make_lr_batch(16, 64, 64, seed=2)
    y = m(lr.cuda()).cpu()
    assert y.shape == (16, 3, 256, 256) and torch.isfinite(y).all()
    ref = O.swinir_forward(lr[[0, 15]], sd, cfg)
    assert (y[[0, 15]] - ref).abs().max().item() <= 2e-3
    assert _psnr_gate(y[[0, 15]], ref, lr[[0, 15]], 4) <= 0.01
    y_perm = m(lr.flip(0).cuda()).cpu().flip(0)
    assert torch.equal(y_perm, y)


def test_unsupported_geometry_is_an_error():
    blk = srk.SwinTransformerBlock(180, (16, 16), 6, window_size=8, shift_size=0, mlp_ratio=2.0).eval().cuda()
    with pytest.raises(RuntimeError):
        blk(torch.zeros(1, 12 * 12, 180, device="cuda"), (12, 12))
    with pytest.raises(RuntimeError):
        blk(torch.zeros(1, 256, 180, device="cuda", dtype=torch.float16), (16, 16))


def test_tiled_inference_matches_per_tile_oracle_and_shards_bit_identically():
    """BASELINE.json configs[4] at a reduced size: overlapping 64-px tiles, E / W stitch, N logical ranks on one GPU."""
    from tpu_superresolution_b200 import tiling
    m, cfg, sd = _model("swinir_x4_d2", "init", 1234)
    lr = synth.make_lr_batch(1, 136, 120, seed=9)
    res = tiling.TiledSuperResolver(m, scale=4, tile=64, overlap=8, batch=16)
    full = res(lr.cuda())
    assert full.shape == (1, 3, 544, 480)
    E, Wt = torch.zeros(3, 544, 480), torch.zeros(544, 480)
    for t in tiling.plan_tiles(136, 120, 64, 8):
        y = O.swinir_forward(lr[:, :, t.y0:t.y0 + 64, t.x0:t.x0 + 64], sd, cfg)[0]
        E[:, 4 * t.y0:4 * t.y0 + 256, 4 * t.x0:4 * t.x0 + 256] += y
        Wt[4 * t.y0:4 * t.y0 + 256, 4 * t.x0:4 * t.x0 + 256] += 1
    assert (full[0].cpu() - E / Wt).abs().max().item() <= 2e-3
    for world in (2, 4):
        bands = [res.band(lr.cuda(), r, world)[0] for r in range(world)]
        assert torch.equal(torch.cat(bands, dim=1), full[0])


def test_block_sync_progress_counters_bit_identical():
    """Image progress counters (SrkBlockSync): the kernels of a BasicLayer order themselves per image instead of per grid.
    The result must be bit-identical to whole-grid ordering, run after run, eager and under CUDA-graph replay."""
    from tpu_superresolution_b200 import swinir
    cfg = synth.CONFIGS["swinir_x4_d2"] if "swinir_x4_d2" in synth.CONFIGS else synth.CONFIGS["swinir_x2_d2"]
    m = srk.SwinIR(**cfg.as_kwargs()).eval()
    m.load_state_dict(synth.make_swinir_state_dict(cfg, seed=77, kind="stress"), strict=True)
    m.cuda()
    x = synth.make_lr_batch(16, 64, 64, seed=3).cuda()
    old = swinir.USE_BLOCK_SYNC
    try:
        with torch.no_grad():
            swinir.USE_BLOCK_SYNC = False
            ref = m(x).clone()
            swinir.USE_BLOCK_SYNC = True
            for _ in range(10):
                assert torch.equal(m(x), ref)
            g = srk.GraphedModel(m)
            for _ in range(10):
                assert torch.equal(g(x), ref)
            # a size whose images are not whole 128-token tiles falls back to whole-grid ordering
            x2 = synth.make_lr_batch(2, 24, 40, seed=4).cuda()
            y2 = m(x2).clone()
            swinir.USE_BLOCK_SYNC = False
            assert torch.equal(m(x2), y2)
    finally:
        swinir.USE_BLOCK_SYNC = old


@pytest.mark.parametrize("B,h,w", [(16, 64, 64), (1, 64, 64), (3, 16, 32), (1, 16, 8), (2, 8, 16), (5, 48, 32)])
def test_layer_kernel_bit_identical_to_per_block_kernels(B, h, w):
    """srk_swin_layer_fwd (one persistent launch per BasicLayer: attention and MLP halves of all blocks as work items ordered by the
    image progress counters) must reproduce the one-launch-per-half-block path bit for bit, run after run -- including the
    degenerate grids where every CTA changes phase at every item (one tile per phase) and the last, ragged round of items."""
    from tpu_superresolution_b200 import swinir
    cfg = synth.CONFIGS["swinir_x4_d2"]
    m = srk.SwinIR(**cfg.as_kwargs()).eval()
    m.load_state_dict(synth.make_swinir_state_dict(cfg, seed=77, kind="stress"), strict=True)
    m.cuda()
    m.set_precision("bf16_strict")                      # the layer kernel implements bf16 operands only
    layer = m.layers[1].residual_group
    x = synth.make_tokens(B, h, w, 180, seed=B + h).cuda()
    old = swinir.USE_LAYER_KERNEL
    try:
        swinir.USE_LAYER_KERNEL = False
        ref = layer(x, (h, w)).clone()
        swinir.USE_LAYER_KERNEL = True
        before = L.launch_count()
        for _ in range(5):
            y = layer(x, (h, w))
            assert torch.equal(y, ref)
        assert L.launch_count() - before == 5           # one launch per call
    finally:
        swinir.USE_LAYER_KERNEL = old


@pytest.mark.parametrize("name", sorted(synth.SWINIR_VARIANTS))
def test_constructor_variants_vs_reference_golden(name):
    """The constructor variants beyond BASELINE's configuration (pixelshuffledirect, nearest+conv, the denoising tail, '3conv', ape,
    x3) against outputs of the unmodified reference.  Name-keyed synthetic weights of order-one logits; outputs reach +-10, so the
    gate is relative: 1e-2 of the output range (bf16 attention / MLP operands; these tails run on library convolutions)."""
    m = srk.SwinIR(**synth.SWINIR_VARIANTS[name]).eval()
    m.load_state_dict(synth.generic_state_dict(m.state_dict(), seed=7), strict=True)
    m.cuda()
    lr = synth.make_lr_batch(2, 16, 16, seed=21)
    before = L.launch_count()
    y = m(lr.cuda()).cpu()
    assert L.launch_count() > before
    ref = torch.from_numpy(np.load(os.path.join(GOLDEN, f"swinir_variant_{name}.npz"))["y"])
    assert y.shape == ref.shape
    assert _rel(y, ref) < 1e-2


def test_pipelined_runner_matches_direct_forward():
    """srk.PipelinedRunner (host-to-host loop with the D2H copy on a side stream) returns exactly the model's outputs, in order."""
    cfg = synth.CONFIGS["swinir_x2_d2"]
    m = srk.SwinIR(**cfg.as_kwargs()).eval()
    m.load_state_dict(synth.make_swinir_state_dict(cfg, seed=5, kind="init"), strict=True)
    m.cuda()
    xs = [synth.make_lr_batch(2, 32, 32, seed=40 + i).pin_memory() for i in range(5)]
    with torch.no_grad():
        refs = [m(x.cuda()).cpu() for x in xs]
    pr = srk.PipelinedRunner(srk.GraphedModel(m), depth=2)
    outs = [torch.empty_like(refs[0]).pin_memory() for _ in xs]
    for x, o in zip(xs, outs):
        ev = pr.submit(x, o)
        assert ev is not None           # host input: the event after which the pinned buffer may be refilled
        ev.synchronize()
    pr.drain()
    torch.cuda.synchronize()
    for o, r in zip(outs, refs):
        assert torch.equal(o, r)
