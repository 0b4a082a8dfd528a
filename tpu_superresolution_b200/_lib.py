"""ctypes binding of libsrk.so (include/srk.h).  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int32, c_int64, c_void_p, c_char_p, POINTER, Structure

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRK_LIB") or os.path.join(_HERE, "lib", "libsrk.so")      # SRK_LIB: the debug build (__graft_entry__.build_debug)

# mirrors of the #defines in include/srk.h
ABI_VERSION = 3
DIM, DIM_PAD, HEADS, HEAD_DIM, HEAD_PAD, WINDOW, HIDDEN, HIDDEN_PAD = 180, 192, 6, 30, 32, 8, 360, 384
ATTN_WSTREAM_BYTES = 3 * 24576 + 9 * 16384 + 3 * 24576
MLP_WSTREAM_BYTES = 9 * 16384 + 6 * 24576
AV_BIAS_Q, AV_BIAS_PROJ, AV_RPB, AV_RPB_STRIDE = 384, 1024, 1216, 232
ATTN_VEC_FLOATS = 1216 + 6 * 232
MV_B1, MV_B2, MLP_VEC_FLOATS = 384, 768, 960
MODE_IMAGE, MODE_WINDOWS = 0, 1
MASK_NONE, MASK_SHIFT, MASK_EXPLICIT = 0, 1, 2
OPERANDS = {"bf16": 0, "fp16": 1, "fp16_fast": 2}      # include/srk.h: SRK_OPERANDS_* ("fp16" = the tight mode; "fp16_fast": MLP only)


class SwinAttnDesc(Structure):
    _fields_ = [(n, c_int32) for n in ("mode", "batch", "height", "width", "num_windows", "ld_in", "ld_out", "shift",
                                       "apply_ln", "add_residual", "mask_mode", "mask_nw", "operands")]


class BlockSync(Structure):
    """include/srk.h: SrkBlockSync (image progress counters between the halves of consecutive Swin blocks)."""
    _fields_ = [("progress", c_void_p), ("batch", c_int32), ("tokens_per_image", c_int32), ("wait_target", c_int32)]


class LayerBlock(Structure):
    _fields_ = [("attn_wstream", c_void_p), ("attn_vec", c_void_p), ("mlp_wstream", c_void_p), ("mlp_vec", c_void_p),
                ("shift", c_int32), ("reserved", c_int32)]


class LayerDesc(Structure):
    _fields_ = [(n, c_int32) for n in ("batch", "height", "width", "ld", "n_blocks")]


LAYER_MAX_BLOCKS = 8


class MlpDesc(Structure):
    _fields_ = [("num_tokens", c_int64), ("ld_in", c_int32), ("ld_out", c_int32), ("apply_ln", c_int32),
                ("add_residual", c_int32), ("operands", c_int32)]


class LinearDesc(Structure):
    _fields_ = [("num_tokens", c_int64)] + [(n, c_int32) for n in ("a_mode", "k_atoms", "ld_in", "apply_ln", "n_chunks", "act",
                                                                    "out_mode", "ld_out", "add_residual")] + \
               [("plane_phase_mask", ctypes.c_uint32)]


class ConvDesc(Structure):
    _fields_ = [(n, c_int32) for n in ("batch", "height", "width", "k_atoms", "np", "cout", "out_mode", "ld_out", "act")] + \
               [("slope", ctypes.c_float), ("a_atoms", c_int32)]


CONV_OUT_ROWS_F32, CONV_OUT_NHWC_F16, CONV_OUT_SHUFFLE2_F16, CONV_OUT_IMAGE = 0, 1, 2, 3


class WinAttnDesc(Structure):
    _fields_ = [(n, c_int32) for n in ("kind", "batch", "height", "width", "shift_y", "shift_x", "mask_shift", "n_heads",
                                       "emask_nw", "out_mode", "out_ld", "out_col0")]


LIN_A_ROWS, LIN_A_PLANES = 0, 1
LIN_OUT_PLANES, LIN_OUT_ROWS = 0, 1
LIN_ACT_NONE, LIN_ACT_GELU = 0, 1
LIN_SLAB_BYTES = 24576
WA_HAT_WMSA, WA_HAT_OCAB, WA_DAT_8x32, WA_DAT_32x8 = 0, 1, 2, 3

_lib = None
# bench.py sets this to a dict to get CUDA-event timings of every launch: {"swin_attn": [(ev0, ev1), ...], ...}
PROFILE = None
PROFILE_DETAIL = False          # True: token_linear launches are keyed by shape in PROFILE


class _launch:
    """One library call: checks that all tensors live on ONE CUDA device, makes that device current for the duration of the
    call (the kernels' per-device state -- opt-in shared memory size, SM count -- and the launch itself follow the current
    device, not the tensors) and hands out that device's current stream.  With PROFILE set it also brackets the call with
    CUDA events on that stream."""

    def __init__(self, name, *tensors):
        self.name = name
        dev = None
        for t in tensors:
            if t is None:
                continue
            if not t.is_cuda:
                raise RuntimeError("tpu_superresolution_b200 kernels need CUDA tensors (no CPU fallback)")
            if dev is None:
                dev = t.device
            elif t.device != dev:
                raise RuntimeError(f"{name}: tensors on different devices ({dev} and {t.device})")
        if dev is None:
            raise RuntimeError(f"{name}: no CUDA tensor")
        self.device = dev
        self._guard = None

    def __enter__(self):
        if torch.cuda.current_device() != self.device.index:
            self._guard = torch.cuda.device(self.device)
            self._guard.__enter__()
        self.stream = torch.cuda.current_stream(self.device)
        if PROFILE is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record(self.stream)
        return self.stream.cuda_stream

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record(self.stream)
            PROFILE.setdefault(self.name, []).append((self.e0, self.e1))
        if self._guard is not None:
            self._guard.__exit__(*exc)
            self._guard = None
        return False


def load():
    """Load libsrk.so (built in-tree by __graft_entry__.build()); raise if it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "There is no CPU / eager fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.srk_abi_version.restype = c_int32
    lib.srk_last_error_string.restype = c_char_p
    lib.srk_launch_count.restype = c_int64
    lib.srk_swin_attn_fwd.argtypes = [POINTER(SwinAttnDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.srk_swin_mlp_fwd.argtypes = [POINTER(MlpDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.srk_swin_attn_fwd_sync.argtypes = [POINTER(SwinAttnDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(BlockSync), c_void_p]
    lib.srk_swin_mlp_fwd_sync.argtypes = [POINTER(MlpDesc), c_void_p, c_void_p, c_void_p, c_void_p, POINTER(BlockSync), c_void_p]
    lib.srk_swin_layer_fwd.argtypes = [POINTER(LayerDesc), c_void_p, POINTER(LayerBlock), c_void_p, c_void_p]
    lib.srk_layernorm_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p]
    lib.srk_layernorm_f16_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p]
    lib.srk_pixelshuffle_nhwc_fwd.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]
    lib.srk_pixelshuffle_nhwc_bias_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]
    lib.srk_bias_act_add_nhwc.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, ctypes.c_float, c_void_p]
    lib.srk_stitch_accumulate.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                          c_int32, c_int32, c_void_p]
    lib.srk_stitch_normalize.argtypes = [c_void_p, c_void_p, c_int32, c_int64, c_void_p]
    lib.srk_conv3x3_fwd.argtypes = [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.srk_rows_to_f16.argtypes = [c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int64, c_void_p]
    lib.srk_rows_to_f16_split.argtypes = [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int32,
                                          ctypes.c_float, c_int32, c_int32, c_void_p]
    lib.srk_image_to_f16_split.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32,
                                           POINTER(ctypes.c_float), ctypes.c_float, c_void_p, c_void_p]
    lib.srk_gather_tiles.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    lib.srk_stitch_accumulate_strided.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int32, c_int32,
                                                  c_int32, c_int32, c_int32, c_int32, c_void_p]
    lib.srk_stitch_finalize.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p]
    lib.srk_linear_fwd.argtypes = [POINTER(LinearDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.srk_window_attention_fwd.argtypes = [POINTER(WinAttnDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p]
    lib.srk_window_attention_table_floats.argtypes = [c_int32]
    f32 = ctypes.c_float
    lib.srk_dwconv3x3_rows_fwd.argtypes = [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]
    lib.srk_row_stats_fwd.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int64, f32, c_void_p, c_void_p]
    lib.srk_dat_mix_fwd.argtypes = [c_void_p] * 6 + [f32, c_int32, c_int32, c_void_p, c_int64, c_int32, c_void_p]
    lib.srk_dat_channel_gram_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]
    lib.srk_dat_channel_gram_ws_floats.argtypes = [c_int32, c_int32]
    lib.srk_cab_ws_floats.argtypes = [c_int32, c_int32]
    lib.srk_token_mean_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]
    lib.srk_dat_channel_softmax_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    lib.srk_dwconv3x3_rows_planes_fwd.argtypes = [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                  c_int32, c_int32, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]
    lib.srk_token_mean_mlp_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]
    lib.srk_dat_channel_apply_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]
    lib.srk_cab_gate_add.argtypes = [c_void_p] * 8 + [c_int32, ctypes.c_float, c_int32, c_int32, c_void_p]
    lib.srk_debug_set_timeline.argtypes = [c_void_p]
    lib.srk_debug_set_timeline.restype = None
    lib.srk_debug_set_stagger.argtypes = [c_int32, c_int32]
    lib.srk_debug_set_stagger.restype = None
    lib.srk_debug_set_winattn_stagger.argtypes = [c_int32]
    lib.srk_debug_set_winattn_stagger.restype = None
    lib.srk_debug_set_pdl.argtypes = [c_int32]
    lib.srk_debug_set_pdl.restype = None
    for f in ("srk_swin_attn_fwd", "srk_swin_mlp_fwd", "srk_swin_attn_fwd_sync", "srk_swin_mlp_fwd_sync", "srk_swin_layer_fwd", "srk_layernorm_fwd", "srk_layernorm_f16_fwd", "srk_pixelshuffle_nhwc_fwd", "srk_pixelshuffle_nhwc_bias_fwd", "srk_bias_act_add_nhwc",
              "srk_stitch_accumulate", "srk_stitch_normalize", "srk_gather_tiles", "srk_stitch_accumulate_strided", "srk_stitch_finalize",
              "srk_conv3x3_fwd", "srk_rows_to_f16", "srk_rows_to_f16_split", "srk_image_to_f16_split", "srk_linear_fwd", "srk_window_attention_fwd",
              "srk_window_attention_table_floats", "srk_cab_gate_add", "srk_dwconv3x3_rows_fwd", "srk_row_stats_fwd", "srk_dat_mix_fwd",
              "srk_dat_channel_gram_fwd", "srk_dat_channel_apply_fwd", "srk_dat_channel_gram_ws_floats", "srk_cab_ws_floats", "srk_token_mean_fwd", "srk_token_mean_mlp_fwd", "srk_dat_channel_softmax_fwd", "srk_dwconv3x3_rows_planes_fwd"):
        getattr(lib, f).restype = c_int32
    if lib.srk_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libsrk.so ABI {lib.srk_abi_version()} != expected {ABI_VERSION}; rebuild")
    if os.environ.get("SRK_WINATTN_STAGGER"):       # experiment switch: start skew of the second softmax group of winattn_kernel (cycles)
        lib.srk_debug_set_winattn_stagger(int(os.environ["SRK_WINATTN_STAGGER"]))
    if os.environ.get("SRK_PDL") == "0":            # experiment switch: no programmatic dependent launch (include/srk.h: srk_debug_set_pdl)
        lib.srk_debug_set_pdl(0)
    _lib = lib
    return lib


EXPORTS = ("srk_abi_version", "srk_last_error_string", "srk_launch_count", "srk_swin_attn_fwd", "srk_swin_mlp_fwd",
           "srk_swin_attn_fwd_sync", "srk_swin_mlp_fwd_sync", "srk_swin_layer_fwd",
           "srk_layernorm_fwd", "srk_layernorm_f16_fwd", "srk_pixelshuffle_nhwc_fwd", "srk_pixelshuffle_nhwc_bias_fwd", "srk_bias_act_add_nhwc", "srk_stitch_accumulate", "srk_stitch_normalize", "srk_gather_tiles", "srk_stitch_accumulate_strided", "srk_stitch_finalize",
           "srk_conv3x3_fwd", "srk_rows_to_f16", "srk_rows_to_f16_split", "srk_image_to_f16_split", "srk_debug_set_timeline", "srk_debug_set_stagger",
           "srk_linear_fwd", "srk_window_attention_fwd", "srk_window_attention_table_floats", "srk_debug_set_winattn_stagger", "srk_debug_set_pdl",
           "srk_cab_gate_add", "srk_dwconv3x3_rows_fwd", "srk_row_stats_fwd", "srk_dat_mix_fwd", "srk_dat_channel_gram_fwd",
           "srk_dat_channel_apply_fwd", "srk_dat_channel_gram_ws_floats", "srk_cab_ws_floats", "srk_token_mean_fwd", "srk_token_mean_mlp_fwd", "srk_dat_channel_softmax_fwd", "srk_dwconv3x3_rows_planes_fwd")


def _check(rc: int, lib) -> None:
    if rc != 0:
        raise RuntimeError("libsrk: " + lib.srk_last_error_string().decode())


def _require_cuda_f32(*tensors) -> None:
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("tpu_superresolution_b200 kernels need CUDA tensors (no CPU fallback)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"expected float32 activations, got {t.dtype}")


def launch_count() -> int:
    return int(load().srk_launch_count())


def _block_sync(progress, batch, tokens_per_image, wait_target):
    if progress is None:
        return None
    if not (progress.is_cuda and progress.dtype == torch.int32 and progress.is_contiguous() and progress.numel() >= 2 * batch):
        raise RuntimeError("progress must be a contiguous CUDA int32 tensor of at least 2 * batch elements")
    return BlockSync(progress.data_ptr(), batch, tokens_per_image, wait_target)


def swin_attn(x, y, wstream, vec, *, mode, batch=0, height=0, width=0, num_windows=0, ld_in, ld_out, shift=0,
              apply_ln=True, add_residual=True, mask_mode=MASK_NONE, mask=None, progress=None, wait_target=0, operands="bf16") -> None:
    """progress / wait_target: image progress counters (include/srk.h: SrkBlockSync); None = whole-grid ordering."""
    lib = load()
    _require_cuda_f32(x, y, vec, mask)
    d = SwinAttnDesc(mode, batch, height, width, num_windows, ld_in, ld_out, shift, int(apply_ln), int(add_residual),
                     mask_mode, 0 if mask is None else mask.shape[0], OPERANDS[operands])
    sync = _block_sync(progress, batch, height * width, wait_target)
    with _launch("swin_attn", x, y, wstream, vec, mask, progress) as st:
        _check(lib.srk_swin_attn_fwd_sync(ctypes.byref(d), x.data_ptr(), y.data_ptr(), wstream.data_ptr(), vec.data_ptr(),
                                          0 if mask is None else mask.data_ptr(), None if sync is None else ctypes.byref(sync), st), lib)


def swin_mlp(x, y, wstream, vec, *, num_tokens, ld_in, ld_out, apply_ln=True, add_residual=True, progress=None, batch=0,
             tokens_per_image=0, wait_target=0, operands="bf16") -> None:
    lib = load()
    _require_cuda_f32(x, y, vec)
    d = MlpDesc(num_tokens, ld_in, ld_out, int(apply_ln), int(add_residual), OPERANDS[operands])
    sync = _block_sync(progress, batch, tokens_per_image, wait_target)
    with _launch("swin_mlp", x, y, wstream, vec, progress) as st:
        _check(lib.srk_swin_mlp_fwd_sync(ctypes.byref(d), x.data_ptr(), y.data_ptr(), wstream.data_ptr(), vec.data_ptr(),
                                         None if sync is None else ctypes.byref(sync), st), lib)


def swin_layer(y, blocks, progress, *, batch, height, width, ld) -> None:
    """srk_swin_layer_fwd: y = blocks(y) in place for all blocks of a BasicLayer in one persistent launch.
    blocks: list of (attn_wstream, attn_vec, mlp_wstream, mlp_vec, shift); progress: int32 CUDA tensor of >= 2 * batch elements."""
    lib = load()
    _require_cuda_f32(y)
    if len(blocks) > LAYER_MAX_BLOCKS:
        raise RuntimeError(f"swin_layer: at most {LAYER_MAX_BLOCKS} blocks per launch")
    if not (progress.is_cuda and progress.dtype == torch.int32 and progress.is_contiguous() and progress.numel() >= 2 * batch):
        raise RuntimeError("progress must be a contiguous CUDA int32 tensor of at least 2 * batch elements")
    arr = (LayerBlock * len(blocks))()
    tensors = [y, progress]
    for k, (aw, av, mw, mv, shift) in enumerate(blocks):
        _require_cuda_f32(av, mv)
        arr[k] = LayerBlock(aw.data_ptr(), av.data_ptr(), mw.data_ptr(), mv.data_ptr(), int(shift), 0)
        tensors += [aw, av, mw, mv]
    d = LayerDesc(batch, height, width, ld, len(blocks))
    with _launch("swin_layer", *tensors) as st:
        _check(lib.srk_swin_layer_fwd(ctypes.byref(d), y.data_ptr(), arr, progress.data_ptr(), st), lib)


def layernorm(x, y, w, b, *, num_tokens, ld_in, ld_out) -> None:
    """srk_layernorm_fwd; y may be x itself (in place)."""
    lib = load()
    _require_cuda_f32(x, y, w, b)
    with _launch("layernorm", x, y, w, b) as st:
        _check(lib.srk_layernorm_fwd(x.data_ptr(), y.data_ptr(), w.data_ptr(), b.data_ptr(), num_tokens, ld_in, ld_out, st), lib)


def layernorm_f16(x, y16, w, b, *, num_tokens, ld_in, y=None, ld_out=0) -> None:
    """srk_layernorm_f16_fwd: LayerNorm(x) as fp16 NHWC rows (num_tokens, 192) in y16 (and, optionally, fp32 rows in y)."""
    lib = load()
    _require_cuda_f32(x, w, b, y)
    if y16.dtype != torch.float16 or not y16.is_contiguous() or y16.numel() != num_tokens * DIM_PAD:
        raise RuntimeError("layernorm_f16: y16 must be a contiguous fp16 (num_tokens, 192) tensor")
    with _launch("layernorm", x, y16, w, b, y) as st:
        _check(lib.srk_layernorm_f16_fwd(x.data_ptr(), _ptr(y), y16.data_ptr(), w.data_ptr(), b.data_ptr(), num_tokens, ld_in, ld_out, st), lib)


def pixelshuffle_nhwc(x, y, *, batch, height, width, out_channels, r, bias=None) -> None:
    lib = load()
    _require_cuda_f32(x, y, bias)
    with _launch("pixelshuffle", x, y, bias) as st:
        _check(lib.srk_pixelshuffle_nhwc_bias_fwd(x.data_ptr(), _ptr(bias), y.data_ptr(), batch, height, width, out_channels, r, st), lib)


ACT_NONE, ACT_LEAKY_RELU, ACT_GELU = 0, 1, 2


def bias_act_add_nhwc(x, y, *, pixels, channels, bias=None, residual=None, act=ACT_NONE, slope=0.0) -> None:
    """y[p, c] = act(x[p, c] + bias[c]) + residual[p, c] (include/srk.h: srk_bias_act_add_nhwc); y may alias x / residual."""
    lib = load()
    _require_cuda_f32(x, y, bias, residual)
    with _launch("bias_act_add", x, y, bias, residual) as st:
        _check(lib.srk_bias_act_add_nhwc(x.data_ptr(), _ptr(bias), _ptr(residual), y.data_ptr(), pixels, channels, act, float(slope), st), lib)


def stitch_accumulate(tiles, E, Wt, tile_yx, *, channels, tile_h, tile_w, out_h, out_w) -> None:
    lib = load()
    _require_cuda_f32(tiles, E, Wt)
    with _launch("stitch_accumulate", tiles, E, Wt, tile_yx) as st:
        _check(lib.srk_stitch_accumulate(tiles.data_ptr(), E.data_ptr(), Wt.data_ptr(), tile_yx.data_ptr(), tile_yx.shape[0],
                                         channels, tile_h, tile_w, out_h, out_w, st), lib)


def stitch_normalize(E, Wt, *, channels, pixels) -> None:
    lib = load()
    _require_cuda_f32(E, Wt)
    with _launch("stitch_normalize", E, Wt) as st:
        _check(lib.srk_stitch_normalize(E.data_ptr(), Wt.data_ptr(), channels, pixels, st), lib)


OUT_F32, OUT_BF16, OUT_U8 = 0, 1, 2
OUT_DTYPES = {torch.float32: OUT_F32, torch.bfloat16: OUT_BF16, torch.uint8: OUT_U8}


def gather_tiles(slab, src_yx, out) -> None:
    """srk_gather_tiles: out (n, C, th, tw) <- windows of the contiguous LR band `slab` (C, rows, W) at src_yx (n, 2) int32."""
    lib = load()
    _require_cuda_f32(slab, out)
    C, rows, W = slab.shape
    n, c2, th, tw = out.shape
    if not (slab.is_contiguous() and out.is_contiguous() and src_yx.dtype == torch.int32 and src_yx.is_contiguous()
            and c2 == C and src_yx.shape[0] >= n):
        raise RuntimeError("gather_tiles: contiguous (C, rows, W) slab, (n, C, th, tw) output and an int32 (>= n, 2) table expected")
    with _launch("gather_tiles", slab, src_yx, out) as st:
        _check(lib.srk_gather_tiles(slab.data_ptr(), rows * W, rows, W, src_yx.data_ptr(), n, C, th, tw, out.data_ptr(), st), lib)


def stitch_accumulate_strided(tiles, E, dst_yx, *, num_tiles=None) -> None:
    """srk_stitch_accumulate_strided: E[c, y0+ty, x0+tx] += tiles[k, c, ty, tx]; `tiles` (n, C, th, tw) with any strides, `E`
    (C, out_h, out_w) with contiguous planes (any channel stride), `dst_yx` int32 (>= n, 2).  The tiles must be disjoint."""
    lib = load()
    _require_cuda_f32(tiles, E)
    n, C, th, tw = tiles.shape
    n = n if num_tiles is None else num_tiles
    if E.stride(2) != 1 or E.stride(1) != E.shape[2] or dst_yx.dtype != torch.int32 or not dst_yx.is_contiguous() or dst_yx.shape[0] < n:
        raise RuntimeError("stitch_accumulate_strided: E planes must be contiguous, dst_yx int32 (>= n, 2)")
    sn, sc, sy, sx = tiles.stride()
    with _launch("stitch_accumulate", tiles, E, dst_yx) as st:
        _check(lib.srk_stitch_accumulate_strided(tiles.data_ptr(), sn, sc, sy, sx, E.data_ptr(), E.stride(0), dst_yx.data_ptr(), n, C, th, tw,
                                                 E.shape[1], E.shape[2], st), lib)


def stitch_finalize(E, cnt_y, cnt_x, out) -> None:
    """srk_stitch_finalize: out = E / (cnt_y[:, None] * cnt_x[None, :]) converted to out.dtype (fp32 / bf16 / uint8)."""
    lib = load()
    _require_cuda_f32(E, cnt_y, cnt_x)
    if out.dtype not in OUT_DTYPES or out.shape != E.shape:
        raise RuntimeError("stitch_finalize: out must be fp32 / bf16 / uint8 with E's shape")
    for t in (E, out):
        if t.stride(2) != 1 or t.stride(1) != t.shape[2]:
            raise RuntimeError("stitch_finalize: planes must be contiguous")
    if cnt_y.numel() != E.shape[1] or cnt_x.numel() != E.shape[2]:
        raise RuntimeError("stitch_finalize: count tables do not match E")
    with _launch("stitch_finalize", E, cnt_y, cnt_x, out) as st:
        _check(lib.srk_stitch_finalize(E.data_ptr(), E.stride(0), cnt_y.data_ptr(), cnt_x.data_ptr(), out.data_ptr(), out.stride(0),
                                       OUT_DTYPES[out.dtype], E.shape[0], E.shape[1], E.shape[2], st), lib)


def conv3x3(x16, wstream, bias, out, *, batch, height, width, k_atoms, np_, cout, out_mode, ld_out, act=ACT_NONE, slope=0.0,
            residual=None, a_atoms=0) -> None:
    """srk_conv3x3_fwd (include/srk.h): x16 = fp16 NHWC (batch, height, width, 64 a_atoms), a_atoms = k_atoms unless given (the
    tight mode's [lo | hi] image: a_atoms = 2/3 k_atoms); out per out_mode."""
    lib = load()
    in_atoms = a_atoms or k_atoms
    if x16.dtype != torch.float16 or x16.numel() != batch * height * width * 64 * in_atoms or not x16.is_contiguous():
        raise RuntimeError("conv3x3: input must be a contiguous fp16 NHWC tensor of batch * height * width * 64 * a_atoms elements")
    want = torch.float32 if out_mode in (CONV_OUT_ROWS_F32, CONV_OUT_IMAGE) else torch.float16
    if out.dtype != want or not out.is_contiguous():
        raise RuntimeError(f"conv3x3: output must be contiguous {want}")
    if wstream.numel() != 9 * k_atoms * np_ * 128 or bias.numel() != np_:
        raise RuntimeError("conv3x3: weight stream / bias size does not match k_atoms, np")
    _require_cuda_f32(bias, residual)
    d = ConvDesc(batch, height, width, k_atoms, np_, cout, out_mode, ld_out, act, float(slope), a_atoms)
    label = "conv3x3" if PROFILE is None else f"conv3x3_c{64 * k_atoms}_n{np_}"
    with _launch(label, x16, wstream, bias, out, residual) as st:
        _check(lib.srk_conv3x3_fwd(ctypes.byref(d), x16.data_ptr(), wstream.data_ptr(), bias.data_ptr(), _ptr(residual), out.data_ptr(), st), lib)


def rows_to_f16(x, out16, *, channels, ld_in, pixels) -> None:
    """srk_rows_to_f16: fp32 rows (pixels, ld_in) -> fp16 NHWC (pixels, cp), cp = out16.shape[-1] (multiple of 64), zero padded."""
    lib = load()
    _require_cuda_f32(x)
    cp = out16.shape[-1]
    if out16.dtype != torch.float16 or not out16.is_contiguous() or out16.numel() != pixels * cp:
        raise RuntimeError("rows_to_f16: output must be a contiguous fp16 (pixels, cp) tensor")
    with _launch("rows_to_f16", x, out16) as st:
        _check(lib.srk_rows_to_f16(x.data_ptr(), ld_in, channels, out16.data_ptr(), cp, pixels, st), lib)


def rows_to_f16_split(x, out16, *, channels, ld_in, pixels, act=ACT_NONE, slope=0.0, shuffle=None) -> None:
    """srk_rows_to_f16_split: fp32 rows (pixels, ld_in) -> the tight mode's fp16 pair image out16 (P, 2 cp) = [lo | hi],
    hi = fp16(act(x)), lo = fp16(act(x) - hi), cp = 64 * ceil(channels / 64) zero-padded channels each.
    shuffle = (H, W): x holds the 256 channels of a conv + PixelShuffle(2) stage at H x W pixels per image; P = 4 * pixels, cp = 64."""
    lib = load()
    _require_cuda_f32(x)
    opix = pixels * (4 if shuffle else 1)
    if out16.dtype != torch.float16 or not out16.is_contiguous() or out16.dim() != 2 or out16.shape[0] != opix or out16.shape[1] % 128:
        raise RuntimeError(f"rows_to_f16_split: output must be a contiguous fp16 ({opix}, 2 cp) tensor, got {tuple(out16.shape)}")
    cp = out16.shape[1] // 2
    sh, sw = shuffle if shuffle else (0, 0)
    with _launch("rows_to_f16_split", x, out16) as st:
        _check(lib.srk_rows_to_f16_split(x.data_ptr(), ld_in, channels, out16.data_ptr() + 2 * cp, out16.data_ptr(), 2 * cp, cp, pixels, act, slope,
                                         sh, sw, st), lib)


def image_to_f16_split(x, out16, mean, img_range) -> None:
    """srk_image_to_f16_split: (B, C <= 3, H, W) fp32 (any strides) -> fp16 NHWC (B * H * W, 64) of (x - mean) * range, hi / lo split."""
    lib = load()
    _require_cuda_f32(x)
    B, C, H, W = x.shape
    if out16.dtype != torch.float16 or not out16.is_contiguous() or out16.numel() != B * H * W * 64:
        raise RuntimeError("image_to_f16_split: output must be a contiguous fp16 (B * H * W, 64) tensor")
    m = (ctypes.c_float * 3)(*([float(v) for v in mean] + [0.0] * 3)[:3])
    sb, sc, sy, sx = x.stride()
    with _launch("image_to_f16", x, out16) as st:
        _check(lib.srk_image_to_f16_split(x.data_ptr(), sb, sc, sy, sx, C, B, H, W, m, float(img_range), out16.data_ptr(), st), lib)


_ZERO_PAGES = {}


def zero_page(device) -> torch.Tensor:
    """8 KB padding pages (packing.make_pad_pages): source of the bulk copies that fill OCAB's out-of-image key / value rows."""
    key = str(device)
    if key not in _ZERO_PAGES:
        from . import packing
        _ZERO_PAGES[key] = packing.make_pad_pages(device)
    return _ZERO_PAGES[key]


def linear(a, wstream, bias, out, *, num_tokens, a_mode, k_atoms=3, ld_in=0, apply_ln=False, n_chunks, act=LIN_ACT_NONE,
           out_mode, ld_out=0, add_residual=False, plane_phase_mask=0) -> None:
    """srk_linear_fwd: a = fp32 rows (A_ROWS) or uint8/bf16 plane buffer (A_PLANES); out = plane buffer or fp32 rows."""
    lib = load()
    if wstream.numel() * wstream.element_size() != n_chunks * k_atoms * LIN_SLAB_BYTES or bias.numel() != n_chunks * 192:
        raise RuntimeError("linear: weight stream / bias size does not match n_chunks, k_atoms")
    d = LinearDesc(num_tokens, a_mode, k_atoms, ld_in, int(apply_ln), n_chunks, act, out_mode, ld_out, int(add_residual),
                   plane_phase_mask)
    label = "linear" if PROFILE is None or not PROFILE_DETAIL else \
        f"linear {'rows' if a_mode == LIN_A_ROWS else 'planes'}(k{k_atoms}){'+ln' if apply_ln else ''} -> {'planes' if out_mode == LIN_OUT_PLANES else 'rows'} x{n_chunks}{' gelu' if act else ''}{' +res' if add_residual else ''}"
    with _launch(label, a, wstream, bias, out) as st:
        _check(lib.srk_linear_fwd(ctypes.byref(d), a.data_ptr(), wstream.data_ptr(), bias.data_ptr(), out.data_ptr(), st), lib)


def window_attention(q_planes, k_planes, v_planes, table, out, *, kind, batch, height, width, shift=(0, 0), mask_shift=False,
                     n_heads=6, emask=None, out_mode=0, out_ld=0, out_col0=0) -> None:
    """srk_window_attention_fwd on plane buffers written by linear()."""
    lib = load()
    if table.numel() != 2 * ((n_heads + 1) // 2) * lib.srk_window_attention_table_floats(kind):
        raise RuntimeError("window_attention: bias table size does not match kind / n_heads")
    d = WinAttnDesc(kind, batch, height, width, shift[0], shift[1], int(mask_shift), n_heads,
                    0 if emask is None else emask.shape[0], out_mode, out_ld, out_col0)
    zp = zero_page(out.device)
    with _launch("window_attention", q_planes, k_planes, v_planes, table, emask, out, zp) as st:
        _check(lib.srk_window_attention_fwd(ctypes.byref(d), q_planes.data_ptr(), k_planes.data_ptr(), v_planes.data_ptr(),
                                            table.data_ptr(), _ptr(emask), zp.data_ptr(), out.data_ptr(), st), lib)


def cab_gate_add(y, out, w1, b1, w2, b2, *, scale, batch, tokens_per_image, y_bias=None) -> None:
    """srk_cab_gate_add: out += scale * (y + y_bias) * squeeze_excite_gate(y + y_bias) on channels-last (batch, tokens, 180) fp32."""
    lib = load()
    _require_cuda_f32(y, out, w1, b1, w2, b2, y_bias)
    ws = torch.empty(lib.srk_cab_ws_floats(batch, tokens_per_image), dtype=torch.float32, device=y.device)
    with _launch("cab_gate_add", y, out, w1, b1, w2, b2, y_bias) as st:
        _check(lib.srk_cab_gate_add(y.data_ptr(), _ptr(y_bias), out.data_ptr(), ws.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                    b2.data_ptr(), w1.shape[0], float(scale), batch, tokens_per_image, st), lib)


def token_mean(x, *, batch, tokens_per_image):
    """srk_token_mean_fwd: (batch, tokens, 180) fp32 rows -> (batch, 180) mean over the tokens (deterministic)."""
    lib = load()
    _require_cuda_f32(x)
    ws = torch.empty(lib.srk_cab_ws_floats(batch, tokens_per_image), dtype=torch.float32, device=x.device)
    mean = torch.empty(batch, DIM, dtype=torch.float32, device=x.device)
    with _launch("token_mean", x) as st:
        _check(lib.srk_token_mean_fwd(x.data_ptr(), mean.data_ptr(), ws.data_ptr(), batch, tokens_per_image, st), lib)
    return mean


def token_mean_mlp(x, w1, b1, w2, b2, *, batch, tokens_per_image):
    """srk_token_mean_mlp_fwd: token mean of (batch, tokens, 180) rows, then w2 gelu(w1 mean + b1) + b2 -> (batch, 180)."""
    lib = load()
    _require_cuda_f32(x, w1, b1, w2, b2)
    ws = torch.empty(lib.srk_cab_ws_floats(batch, tokens_per_image), dtype=torch.float32, device=x.device)
    out = torch.empty(batch, DIM, dtype=torch.float32, device=x.device)
    with _launch("token_mean", x, w1, b1, w2, b2) as st:
        _check(lib.srk_token_mean_mlp_fwd(x.data_ptr(), out.data_ptr(), ws.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                          w1.shape[0], batch, tokens_per_image, st), lib)
    return out


def dat_channel_softmax(gram, temperature, *, batch):
    """srk_dat_channel_softmax_fwd: (batch, 6, 960) gram -> (batch, 6, 30, 30) attention (normalisation, temperature, softmax)."""
    lib = load()
    _require_cuda_f32(gram, temperature)
    attn = torch.empty((batch, HEADS, HEAD_DIM, HEAD_DIM), dtype=torch.float32, device=gram.device)
    with _launch("dat_channel_softmax", gram, temperature) as st:
        _check(lib.srk_dat_channel_softmax_fwd(gram.data_ptr(), temperature.data_ptr(), attn.data_ptr(), batch, st), lib)
    return attn


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def dwconv3x3_rows(inp, w9c, scale, shift, out, *, ld_in, c_in, ld_out, channels, batch, height, width, act_gelu=False, ln_stats=None,
                   ln_gamma=None, ln_beta=None, gate=None, ld_gate=0, c_gate=0) -> None:
    """srk_dwconv3x3_rows_fwd on fp32 token rows (see include/srk.h)."""
    lib = load()
    _require_cuda_f32(inp, w9c, scale, shift, out, ln_stats, ln_gamma, ln_beta, gate)
    with _launch("dwconv3x3_rows", inp, w9c, scale, shift, out, ln_stats, ln_gamma, ln_beta, gate) as st:
        _check(lib.srk_dwconv3x3_rows_fwd(inp.data_ptr(), ld_in, c_in, w9c.data_ptr(), scale.data_ptr(), shift.data_ptr(), _ptr(ln_stats),
                                          _ptr(ln_gamma), _ptr(ln_beta), _ptr(gate), ld_gate, c_gate, out.data_ptr(), ld_out, channels, batch,
                                          height, width, int(act_gelu), st), lib)


def dwconv3x3_rows_planes(inp, w9c, scale, shift, planes, *, ld_in, c_in, channels, batch, height, width, act_gelu=False, ln_stats=None,
                          ln_gamma=None, ln_beta=None, gate=None, ld_gate=0, c_gate=0) -> None:
    """srk_dwconv3x3_rows_planes_fwd: like dwconv3x3_rows, output as bf16 planes (P, tokens, 64) for linear(a_mode=LIN_A_PLANES)."""
    lib = load()
    _require_cuda_f32(inp, w9c, scale, shift, ln_stats, ln_gamma, ln_beta, gate)
    if planes.dtype != torch.bfloat16 or planes.dim() != 3 or planes.shape[2] != 64 or not planes.is_contiguous() or \
            planes.shape[0] * 64 < channels or planes.shape[1] != batch * height * width:
        raise RuntimeError("dwconv3x3_rows_planes: planes must be a contiguous bf16 (P, tokens, 64) buffer covering the channels")
    with _launch("dwconv3x3_rows", inp, w9c, scale, shift, planes, ln_stats, ln_gamma, ln_beta, gate) as st:
        _check(lib.srk_dwconv3x3_rows_planes_fwd(inp.data_ptr(), ld_in, c_in, w9c.data_ptr(), scale.data_ptr(), shift.data_ptr(), _ptr(ln_stats),
                                                 _ptr(ln_gamma), _ptr(ln_beta), _ptr(gate), ld_gate, c_gate, planes.data_ptr(),
                                                 planes.shape[1] * 128, channels, batch, height, width, int(act_gelu), st), lib)


def row_stats(inp, stats, *, ld_in, c_in, channels, tokens, eps) -> None:
    lib = load()
    _require_cuda_f32(inp, stats)
    with _launch("row_stats", inp, stats) as st:
        _check(lib.srk_row_stats_fwd(inp.data_ptr(), ld_in, c_in, channels, tokens, float(eps), stats.data_ptr(), st), lib)


def dat_mix(att, conv, cmap, w1, b1, w2, b2, mix, *, mode, tokens, tokens_per_image) -> None:
    lib = load()
    _require_cuda_f32(att, conv, cmap, w1, b1, w2, mix)
    with _launch("dat_mix", att, conv, cmap, w1, b1, w2, mix) as st:
        _check(lib.srk_dat_mix_fwd(att.data_ptr(), conv.data_ptr(), cmap.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), float(b2),
                                   w1.shape[0], mode, mix.data_ptr(), tokens, tokens_per_image, st), lib)


def dat_channel_gram(qkv, gram, *, batch, tokens_per_image) -> None:
    lib = load()
    _require_cuda_f32(qkv, gram)
    ws = torch.empty(lib.srk_dat_channel_gram_ws_floats(batch, tokens_per_image), dtype=torch.float32, device=qkv.device)
    with _launch("dat_channel_gram", qkv, gram) as st:
        _check(lib.srk_dat_channel_gram_fwd(qkv.data_ptr(), gram.data_ptr(), ws.data_ptr(), batch, tokens_per_image, st), lib)


def dat_channel_apply(qkv, attn, out, *, batch, tokens_per_image) -> None:
    lib = load()
    _require_cuda_f32(qkv, attn, out)
    with _launch("dat_channel_apply", qkv, attn, out) as st:
        _check(lib.srk_dat_channel_apply_fwd(qkv.data_ptr(), attn.data_ptr(), out.data_ptr(), batch, tokens_per_image, st), lib)
