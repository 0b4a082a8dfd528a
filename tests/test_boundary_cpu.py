"""CPU-side checks of the drop-in boundary: state_dict contract, weight packing, C-ABI exports."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import _lib as L
from tpu_superresolution_b200 import packing
from oracle import synth
from oracle import swinir_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("name", ["swinir_x2", "swinir_x4"])
def test_state_dict_keys_match_reference_manifest(name):
    """Same keys / shapes / dtypes as the reference's state_dict (finetune_swinir.py:283-285 strict=True)."""
    with open(os.path.join(GOLDEN, f"{name}_manifest.json")) as f:
        man = json.load(f)
    model = srk.SwinIR(**synth.CONFIGS[name].as_kwargs())
    sd = model.state_dict()
    assert [m[0] for m in man] == list(sd.keys())
    for k, shape, dtype in man:
        assert list(sd[k].shape) == shape and str(sd[k].dtype).replace("torch.", "") == dtype, k


def test_strict_load_roundtrip_and_buffers():
    cfg = synth.CONFIGS["swinir_x2_d2"]
    sd = synth.make_swinir_state_dict(cfg, seed=5, kind="stress")
    model = srk.SwinIR(**cfg.as_kwargs())
    model.load_state_dict(sd, strict=True)
    out = model.state_dict()
    for k, v in sd.items():
        assert torch.equal(out[k], v), k
    g = np.load(os.path.join(GOLDEN, "kat_buffers.npz"))
    blk = srk.SwinTransformerBlock(180, (64, 64), 6, window_size=8, shift_size=4, mlp_ratio=2.0)
    assert np.array_equal(blk.attn.relative_position_index.numpy(), g["rpi"])
    assert np.array_equal(blk.attn_mask.numpy().astype(np.int8), g["mask_64"])
    assert np.array_equal(blk.calculate_mask((48, 40)).numpy().astype(np.int8), g["mask_48x40"])
    assert srk.SwinTransformerBlock(180, (64, 64), 6, window_size=8, shift_size=0).attn_mask is None


def test_swizzle_is_an_involution_and_matches_device_formula():
    m = torch.arange(128 * 64, dtype=torch.float32).reshape(128, 64)
    s = packing.swizzle_slab(m)
    assert torch.equal(packing.unswizzle_slab(s), m)
    # device side: 16-byte chunk c of row r lives at byte r*128 + ((c ^ (r & 7)) << 4)   (csrc/umma.cuh sw128_off)
    flat = s.reshape(-1)
    for r, c in [(0, 0), (1, 0), (5, 3), (77, 7), (127, 2)]:
        off_elems = (r * 128 + ((c ^ (r & 7)) << 4)) // 2          # bf16 elements
        assert torch.equal(flat[off_elems:off_elems + 8], m[r, 8 * c:8 * c + 8])


def _unpack_slabs(stream_u8, specs):
    """specs: list of row counts; returns list of (rows, 64) float tensors (un-swizzled)."""
    bf = stream_u8.view(torch.bfloat16)
    out, off = [], 0
    for rows in specs:
        out.append(packing.unswizzle_slab(bf[off:off + rows * 64].reshape(rows, 64)).float())
        off += rows * 64
    assert off == bf.numel()
    return out


@pytest.mark.parametrize("with_ln", [False, True])
def test_pack_attention_reconstructs_reference_math(with_ln):
    """Emulate the kernel's dataflow on the CPU from the packed stream alone and compare with the oracle."""
    cfg = synth.CONFIGS["swinir_x2_d2"]
    sd = synth.make_swinir_state_dict(cfg, seed=99, kind="stress")
    blk = "layers.0.residual_group.blocks.1."
    pre = blk + "attn."
    ln = (sd[blk + "norm1.weight"], sd[blk + "norm1.bias"]) if with_ln else (None, None)
    w, vec = packing.pack_attention(sd[pre + "qkv.weight"], sd[pre + "qkv.bias"], sd[pre + "proj.weight"], sd[pre + "proj.bias"],
                                    sd[pre + "relative_position_bias_table"], *ln)
    assert w.numel() == L.ATTN_WSTREAM_BYTES and vec.numel() == L.ATTN_VEC_FLOATS
    slabs = _unpack_slabs(w, [192] * 3 + [128] * 9 + [192] * 3)
    wv = torch.cat(slabs[0:3], 1)                                                        # (192, 192) padded v-dims
    pairs = [torch.cat(slabs[3 + 3 * p:6 + 3 * p], 1) for p in range(3)]                 # 3 x (128, 192): [q_h | k_h | q_h+1 | k_h+1]
    wqk = [pairs[h >> 1][64 * (h & 1):64 * (h & 1) + 64] for h in range(6)]              # 6 x (64, 192): [q_h | k_h]
    wp = torch.cat(slabs[12:15], 1)                                                      # (192, 192)
    xw = synth.make_tokens(4, 8, 8, 180, seed=5)
    xhat = (xw - xw.mean(-1, keepdim=True)) / torch.sqrt(xw.var(-1, unbiased=False, keepdim=True) + 1e-5) if with_ln else xw
    xb = torch.zeros(4, 64, 192)
    xb[..., :180] = xhat.bfloat16().float()
    v = xb @ wv.T                                                                        # (4, 64, 192) padded head layout, no bias
    o = torch.zeros(4, 64, 192)
    rpb = vec[L.AV_RPB:].view(6, L.AV_RPB_STRIDE)
    idx = O.relative_position_index(8)
    for h in range(6):
        qk = xb @ wqk[h].T
        q, k = qk[..., :32] + vec[L.AV_BIAS_Q + 32 * h:L.AV_BIAS_Q + 32 * h + 32], qk[..., 32:]
        s = q @ k.transpose(-1, -2) + rpb[h][idx]                                        # log2 domain
        pm = torch.exp2(s - s.max(-1, keepdim=True).values)
        o[..., 32 * h:32 * h + 32] = (pm / pm.sum(-1, keepdim=True)) @ v[..., 32 * h:32 * h + 32]
    y = o @ wp.T + vec[L.AV_BIAS_PROJ:L.AV_BIAS_PROJ + 192]
    xin = O.layer_norm(xw, *ln) if with_ln else xw
    ref = O.window_attention(xin, sd, pre, 6, 8, None)
    assert (y[..., 180:].abs().max() == 0) and (v[..., 30:32].abs().max() == 0)
    err = (y[..., :180] - ref).abs().max().item()
    assert err < 3e-2 * ref.abs().max().item(), err       # bf16 weights/inputs only


@pytest.mark.parametrize("with_ln", [False, True])
def test_pack_mlp_reconstructs_reference_math(with_ln):
    cfg = synth.CONFIGS["swinir_x2_d2"]
    sd = synth.make_swinir_state_dict(cfg, seed=99, kind="stress")
    blk = "layers.0.residual_group.blocks.0."
    pre = blk + "mlp."
    ln = (sd[blk + "norm2.weight"], sd[blk + "norm2.bias"]) if with_ln else (None, None)
    w, vec = packing.pack_mlp(sd[pre + "fc1.weight"], sd[pre + "fc1.bias"], sd[pre + "fc2.weight"], sd[pre + "fc2.bias"], *ln)
    slabs = _unpack_slabs(w, [128] * 6 + [192] * 2 + [128] * 3 + [192] * 4)
    w1 = torch.cat([torch.cat(slabs[0:3], 1), torch.cat(slabs[3:6], 1), torch.cat(slabs[8:11], 1)], 0)   # (384, 192)
    w2 = torch.cat(slabs[6:8] + slabs[11:15], 1)                                         # (192, 384)
    x = synth.make_tokens(1, 8, 8, 180, seed=7)[0]
    xhat = (x - x.mean(-1, keepdim=True)) / torch.sqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-5) if with_ln else x
    xb = torch.zeros(64, 192)
    xb[:, :180] = xhat
    y = O.gelu(xb @ w1.T + vec[L.MV_B1:L.MV_B1 + 384]) @ w2.T + vec[L.MV_B2:L.MV_B2 + 192]
    ref = O.mlp(O.layer_norm(x, *ln) if with_ln else x, sd, pre)
    assert (y[:, :180] - ref).abs().max() < 2e-2 * ref.abs().max()
    with pytest.raises(RuntimeError):
        packing.pack_mlp(torch.zeros(720, 180), None, torch.zeros(180, 720), None)


def test_gelu_tanh_polynomial_matches_exact_erf_gelu():
    """csrc/swin_kernels.cu gelu_fast: 0.5 x (1 + tanh(x (c1 + c3 u + c5 u^2))), u = min(x^2, 64)."""
    x = torch.linspace(-30, 30, 240001, dtype=torch.float64)
    u = torch.clamp(x * x, max=64.0)
    g = 0.5 * x * (1 + torch.tanh(x * (7.97507881e-01 + u * (3.70056486e-02 + u * -3.51517176e-04))))
    assert (g - O.gelu(x)).abs().max() < 2.6e-5


def test_header_constants_match_python_mirror():
    hdr = open(os.path.join(ROOT, "include", "srk.h")).read()
    defs = dict(re.findall(r"#define\s+(SRK_\w+)\s+\(?([0-9+* ]+)\)?\s", hdr))
    ev = lambda k: eval(defs[k])
    assert ev("SRK_ATTN_WSTREAM_BYTES") == L.ATTN_WSTREAM_BYTES and ev("SRK_MLP_WSTREAM_BYTES") == L.MLP_WSTREAM_BYTES
    assert ev("SRK_ATTN_VEC_FLOATS") == L.ATTN_VEC_FLOATS and ev("SRK_MLP_VEC_FLOATS") == L.MLP_VEC_FLOATS
    for k, v in [("SRK_AV_BIAS_Q", L.AV_BIAS_Q), ("SRK_AV_BIAS_PROJ", L.AV_BIAS_PROJ),
                 ("SRK_AV_RPB", L.AV_RPB), ("SRK_AV_RPB_STRIDE", L.AV_RPB_STRIDE), ("SRK_MV_B1", L.MV_B1), ("SRK_MV_B2", L.MV_B2),
                 ("SRK_ABI_VERSION", L.ABI_VERSION)]:
        assert ev(k) == v, k


def test_library_loads_and_exports_every_declared_symbol():
    """The C-ABI library must exist in-tree and export everything include/srk.h declares (no compute here)."""
    assert os.path.isfile(L.LIB_PATH), "build with __graft_entry__.build()"
    lib = ctypes.CDLL(L.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "srk.h")).read()
    declared = set(re.findall(r"\b(srk_\w+)\s*\(", hdr))
    assert declared == set(L.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert L.load().srk_abi_version() == L.ABI_VERSION


def test_no_cpu_fallback():
    blk = srk.SwinTransformerBlock(180, (16, 16), 6, window_size=8, shift_size=0, mlp_ratio=2.0).eval()
    with torch.no_grad(), pytest.raises(RuntimeError):
        blk(torch.zeros(1, 256, 180), (16, 16))
    with torch.no_grad(), pytest.raises(RuntimeError):
        srk.WindowAttention(96, (7, 7), 3)._packed()


@pytest.mark.parametrize("name", sorted(__import__("tpu_superresolution_b200").synth.SWINIR_VARIANTS))
def test_constructor_variants_keep_the_reference_state_dict(name):
    """Upsampler / resi_connection / ape variants of network_swinir.py (646-764): same keys and shapes as the unmodified reference
    (recorded by oracle/make_golden_variants.py), strict load of name-keyed synthetic weights."""
    import numpy as np
    from tpu_superresolution_b200 import synth
    m = srk.SwinIR(**synth.SWINIR_VARIANTS[name]).eval()
    keys = str(np.load(os.path.join(ROOT, "tests", "golden", f"swinir_variant_{name}.npz"))["keys"]).split("\n")
    assert list(m.state_dict().keys()) == keys
    m.load_state_dict(synth.generic_state_dict(m.state_dict(), seed=7), strict=True)


def test_split_conv_packing():
    """Tight-mode convolution weights (packing.pack_conv3x3(split=True)): the stream is the k-atoms [hi(w) | lo(w) | hi(w)], with
    hi(w) + lo(w) reproducing w to 2^-22, for one launch of 3 C/64 k-steps over the [lo(x) | hi(x)] image (2 C/64 atoms)."""
    g = torch.Generator().manual_seed(5)
    w, b = torch.randn(180, 180, 3, 3, generator=g) * 0.03, torch.randn(180, generator=g)
    plain, bias, meta = packing.pack_conv3x3(w, b)
    split, bias2, meta2 = packing.pack_conv3x3(w, b, split=True)
    assert meta == {"k_atoms": 3, "np": 192, "cout": 180} and meta2 == {"k_atoms": 9, "a_atoms": 6, "np": 192, "cout": 180}
    n = plain.numel()
    assert torch.equal(bias, bias2) and split.numel() == 3 * n
    hi, lo, hi2 = split[:n], split[n:2 * n], split[2 * n:]
    assert torch.equal(hi, plain) and torch.equal(hi2, plain)
    h, l = hi.view(torch.float16).float(), lo.view(torch.float16).float()
    assert l.abs().max() <= 2.0 ** -11 * h.abs().max() and l.abs().max() > 0
    # first slab = output rows x input channels 0..63 at (dy, dx) = (0, 0): hi + lo against the fp32 weights
    first = lambda st: packing.unswizzle_slab(st.view(torch.float16)[:192 * 64].reshape(192, 64)).float()
    err = (first(hi)[:180] + first(lo)[:180] - w[:, :64, 0, 0]).abs().max()
    assert err <= 2.0 ** -21 * w.abs().max()
    w64 = torch.randn(256, 64, 3, 3, generator=g) * 0.05
    s64, _, m3 = packing.pack_conv3x3(w64, None, split=True, pixel_shuffle=True)
    one, _, m1 = packing.pack_conv3x3(w64, None, pixel_shuffle=True)
    assert m3 == {"k_atoms": 3, "a_atoms": 2, "np": 256, "cout": 256} and m1["k_atoms"] == 1 and s64.numel() == 3 * one.numel()
    with pytest.raises(RuntimeError):
        packing.pack_conv3x3(torch.randn(180, 3, 3, 3), None, split_first=True, split=True)


def _emulate_conv_from_stream(x_atoms, wstream, bias, meta):
    """Host model of conv3x3_kernel's data path (csrc/conv_kernel.cu): k-step ka multiplies input atom
    ia = ka if ka < a_atoms else ka - (k_atoms - a_atoms) with the slabs (ka, dx, dy) of the stream; float64 accumulation.
    x_atoms: (B, a_atoms * 64, H, W) fp16 values as float64.  Returns (B, np, H, W)."""
    k_atoms, np_ = meta["k_atoms"], meta["np"]
    a_atoms = meta.get("a_atoms", k_atoms)
    slabs = wstream.view(torch.float16).reshape(k_atoms, 3, 3, np_, 64)                       # [ka][dx][dy][row][k]
    out = bias.double().view(1, np_, 1, 1).repeat(x_atoms.shape[0], 1, x_atoms.shape[2], x_atoms.shape[3])
    for ka in range(k_atoms):
        ia = ka if ka < a_atoms else ka - (k_atoms - a_atoms)
        w = torch.zeros(np_, 64, 3, 3, dtype=torch.float64)
        for dx in range(3):
            for dy in range(3):
                w[:, :, dy, dx] = packing.unswizzle_slab(slabs[ka, dx, dy].view(torch.int16)).view(torch.float16).double()
        out += torch.nn.functional.conv2d(x_atoms[:, 64 * ia:64 * ia + 64], w, padding=1)
    return out


@pytest.mark.parametrize("cin,cout,kw", [(180, 180, {}), (64, 256, {"pixel_shuffle": True}),
                                          (64, 3, {"out_scale": 0.5, "out_shift": [0.4488, 0.4371, 0.4040]})])
def test_split_conv_stream_reproduces_the_fp32_convolution(cin, cout, kw):
    """The tight mode's algebra end to end on the host: activations as the [lo | hi] fp16 pair image (what srk_rows_to_f16_split
    writes), weights as the packed [hi(w) | lo(w) | hi(w)] stream, the kernel's k-step -> input-atom rule -- against the float64
    convolution of the original weights (incl. the pixel-shuffle row permutation and the folded output scale / shift).  Exact
    accumulation here, so what remains is the dropped lo * lo term and the fp16 rounding of lo: <= 1e-6 of the output's magnitude;
    the same stream without the lo parts (plain fp16 operands) is ~1e-4."""
    g = torch.Generator().manual_seed(cin + cout)
    B, H, W = 1, 9, 11
    x = torch.randn(B, cin, H, W, generator=g)
    w, b = torch.randn(cout, cin, 3, 3, generator=g) * 0.05, torch.randn(cout, generator=g) * 0.1
    ws, bias, meta = packing.pack_conv3x3(w, b, split=True, **kw)
    cp = 64 * ((cin + 63) // 64)
    assert meta["k_atoms"] == 3 * cp // 64 and meta["a_atoms"] == 2 * cp // 64
    hi = x.half().float()
    lo = (x - hi).half().float()
    img = torch.zeros(B, 2 * cp, H, W, dtype=torch.float64)
    img[:, :cin], img[:, cp:cp + cin] = lo.double(), hi.double()                  # [lo | hi], zero padded
    y = _emulate_conv_from_stream(img, ws, bias, meta)[:, :cout]
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1)
    if kw.get("pixel_shuffle"):                                                   # packed row s * 64 + c <- original row 4 c + s
        ref = ref.reshape(B, cout // 4, 4, H, W).transpose(1, 2).reshape(B, cout, H, W)
    if "out_scale" in kw:
        ref = ref * kw["out_scale"] + torch.tensor(kw["out_shift"], dtype=torch.float64).view(1, -1, 1, 1)
    scale = ref.abs().max()
    assert (y - ref).abs().max() <= 1e-6 * scale
    plain_ws, plain_bias, plain_meta = packing.pack_conv3x3(w, b, **kw)
    img_hi = torch.zeros(B, cp, H, W, dtype=torch.float64)
    img_hi[:, :cin] = hi.double()
    y16 = _emulate_conv_from_stream(img_hi, plain_ws, plain_bias, plain_meta)[:, :cout]
    assert (y16 - ref).abs().max() > 20 * (y - ref).abs().max()                   # what the split buys


def test_conv_first_split_stream_reproduces_the_fp32_convolution():
    """conv_first (network_swinir.py:720): the image as [hi(v), v - hi(v), hi(v)] inside one 64-channel atom (srk_image_to_f16_split)
    against weights packed [hi(w), hi(w), w - hi(w)] (pack_conv3x3(split_first=True)), through the same host model."""
    g = torch.Generator().manual_seed(3)
    v = torch.rand(2, 3, 10, 12, generator=g) - 0.45
    w, b = torch.randn(180, 3, 3, 3, generator=g) * 0.2, torch.randn(180, generator=g) * 0.1
    ws, bias, meta = packing.pack_conv3x3(w, b, split_first=True)
    assert meta == {"k_atoms": 1, "np": 192, "cout": 180}
    hi = v.half().float()
    img = torch.zeros(2, 64, 10, 12, dtype=torch.float64)
    img[:, 0:3], img[:, 3:6], img[:, 6:9] = hi.double(), (v - hi).half().double(), hi.double()
    y = _emulate_conv_from_stream(img, ws, bias, meta)[:, :180]
    ref = torch.nn.functional.conv2d(v.double(), w.double(), b.double(), padding=1)
    assert (y - ref).abs().max() <= 2e-6 * ref.abs().max()


def test_fp16_operand_packing_and_precision_table():
    """Tight mode / default MLP (include/srk.h: SRK_OPERANDS_*): the same slab stream with fp16 elements, closer to the fp32 weights
    than the bf16 one; weights outside the fp16 range are an error; set_precision() maps the three modes onto the modules."""
    cfg = synth.CONFIGS["swinir_x2_d2"]
    sd = synth.make_swinir_state_dict(cfg, seed=99, kind="stress")
    pre = "layers.0.residual_group.blocks.0.mlp."
    args = (sd[pre + "fc1.weight"], sd[pre + "fc1.bias"], sd[pre + "fc2.weight"], sd[pre + "fc2.bias"])
    w16, v16 = packing.pack_mlp(*args, operands="fp16")
    wbf, vbf = packing.pack_mlp(*args, operands="bf16")
    wfast, _ = packing.pack_mlp(*args, operands="fp16_fast")
    assert w16.numel() == wbf.numel() == L.MLP_WSTREAM_BYTES and torch.equal(v16, vbf) and torch.equal(w16, wfast)
    first16 = packing.unswizzle_slab(w16.view(torch.float16)[:128 * 64].reshape(128, 64)).float()       # fc1 rows 0..127, k 0..63
    firstbf = packing.unswizzle_slab(wbf.view(torch.bfloat16)[:128 * 64].reshape(128, 64)).float()
    ref = sd[pre + "fc1.weight"][:128, :64].float()
    assert (first16 - ref).abs().max() < (firstbf - ref).abs().max() and (first16 - ref).abs().max() <= 2.0 ** -11 * ref.abs().max()
    big = [a.clone() for a in args]
    big[0][0, 0] = 1.0e5
    with pytest.raises(RuntimeError):
        packing.pack_mlp(*big, operands="fp16")
    packing.pack_mlp(*big, operands="bf16")                                                              # bf16 has the range

    m = srk.SwinIR(**cfg.as_kwargs()).eval()
    blk = m.layers[0].residual_group.blocks[0]
    assert (blk.attn.operands, blk.mlp.operands) == ("bf16", "fp16_fast")                               # the default mode
    for mode, want in srk.swinir.PRECISIONS.items():
        m.set_precision(mode)
        assert all((b.attn.operands, b.mlp.operands) == want for layer in m.layers for b in layer.residual_group.blocks)
        # tight mode only: the 3x3 convolutions as hi / lo fp16 pairs (a flag on the model and on every RSTB, no global state)
        assert m.split_conv == (mode == "fp16") and all(layer.split_conv == (mode == "fp16") for layer in m.layers)
    with pytest.raises(ValueError):
        m.set_precision("fp8")
    assert set(L.OPERANDS) == {"bf16", "fp16", "fp16_fast"}
    hdr = open(os.path.join(ROOT, "include", "srk.h")).read()
    for name, val in (("SRK_OPERANDS_BF16", 0), ("SRK_OPERANDS_F16", 1), ("SRK_OPERANDS_F16_HALF_GELU", 2)):
        assert re.search(name + r"\s*=\s*" + str(val), hdr), name
