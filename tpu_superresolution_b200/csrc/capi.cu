// C ABI of libsrk.so (include/srk.h): argument validation, error strings, launch counting.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "kernels.h"

namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
int check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) {
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return 0;
    }
    return fail("%s: %s", what, cudaGetErrorString(e));
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
}  // namespace

extern "C" {

int srk_abi_version(void) { return SRK_ABI_VERSION; }
const char* srk_last_error_string(void) { return g_err; }
void srk_debug_set_stagger(int attn_cycles, int mlp_cycles) { srk::g_stagger_attn = attn_cycles; srk::g_stagger_mlp = mlp_cycles; }
void srk_debug_set_winattn_stagger(int cycles) { srk::g_stagger_winattn = cycles; }
void srk_debug_set_pdl(int enabled) { srk::g_pdl = enabled; }
void srk_debug_set_timeline(void* buf) { srk::g_timeline = static_cast<unsigned long long*>(buf); }
int64_t srk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int srk_swin_attn_fwd(const SrkSwinAttnDesc* d, const float* x, float* y, const void* wstream, const float* vec,
                      const float* mask, void* stream) {
    return srk_swin_attn_fwd_sync(d, x, y, wstream, vec, mask, nullptr, stream);
}

int srk_swin_attn_fwd_sync(const SrkSwinAttnDesc* d, const float* x, float* y, const void* wstream, const float* vec,
                           const float* mask, const SrkBlockSync* sync, void* stream) {
    if (!d || !x || !y || !wstream || !vec) return fail("srk_swin_attn_fwd: null argument");
    if (d->ld_in < SRK_DIM || d->ld_out < SRK_DIM || (d->ld_in & 3) || (d->ld_out & 3))
        return fail("srk_swin_attn_fwd: ld_in/ld_out must be >= %d and multiples of 4 (got %d, %d)", SRK_DIM, d->ld_in, d->ld_out);
    if (!aligned16(x) || !aligned16(y) || !aligned16(wstream) || !aligned16(vec))
        return fail("srk_swin_attn_fwd: pointers must be 16-byte aligned");
    if (d->shift != 0 && d->shift != SRK_WINDOW / 2) return fail("srk_swin_attn_fwd: shift must be 0 or %d (got %d)", SRK_WINDOW / 2, d->shift);
    srk::AttnParams p{};
    p.x = x; p.y = y; p.wstream = static_cast<const uint8_t*>(wstream); p.vec = vec; p.mask = mask;
    p.mode = d->mode; p.shift = d->shift; p.ld_in = d->ld_in; p.ld_out = d->ld_out;
    p.apply_ln = d->apply_ln; p.add_residual = d->add_residual; p.mask_mode = d->mask_mode; p.mask_nw = d->mask_nw;
    p.H = d->height; p.W = d->width;
    const bool need_geom = d->mode == SRK_MODE_IMAGE || d->mask_mode == SRK_MASK_SHIFT;
    if (need_geom) {
        if (d->height <= 0 || d->width <= 0 || d->height % SRK_WINDOW || d->width % SRK_WINDOW)
            return fail("srk_swin_attn_fwd: height/width must be positive multiples of %d (got %d x %d)", SRK_WINDOW, d->height, d->width);
        p.nwx = d->width / SRK_WINDOW;
        p.nw_img = (d->height / SRK_WINDOW) * p.nwx;
    } else {
        p.nwx = 1; p.nw_img = 1;
    }
    if (d->mode == SRK_MODE_IMAGE) {
        if (d->batch <= 0) return fail("srk_swin_attn_fwd: batch must be positive");
        const int64_t tw = static_cast<int64_t>(d->batch) * p.nw_img;
        if (tw > (1ll << 30)) return fail("srk_swin_attn_fwd: too many windows");
        p.total_windows = static_cast<int>(tw);
    } else if (d->mode == SRK_MODE_WINDOWS) {
        if (d->num_windows <= 0) return fail("srk_swin_attn_fwd: num_windows must be positive");
        p.total_windows = d->num_windows;
        if (p.shift != 0) return fail("srk_swin_attn_fwd: shift is only meaningful in SRK_MODE_IMAGE");
    } else {
        return fail("srk_swin_attn_fwd: unknown mode %d", d->mode);
    }
    if (d->mask_mode == SRK_MASK_EXPLICIT) {
        if (!mask || d->mask_nw <= 0 || !aligned16(mask)) return fail("srk_swin_attn_fwd: explicit mask needs a 16-byte aligned pointer and mask_nw > 0");
    } else if (d->mask_mode != SRK_MASK_NONE && d->mask_mode != SRK_MASK_SHIFT) {
        return fail("srk_swin_attn_fwd: unknown mask_mode %d", d->mask_mode);
    }
    if (d->mask_mode == SRK_MASK_SHIFT && d->shift == 0 && d->mode == SRK_MODE_IMAGE) p.mask_mode = SRK_MASK_NONE;
    if (d->mask_mode == SRK_MASK_SHIFT && d->mode == SRK_MODE_WINDOWS) p.shift = SRK_WINDOW / 2;   // regions of the shifted grid
    p.n_tiles = (p.total_windows + 1) / 2;
    if (d->add_residual && x != y) {
        // the kernel adds into y (bulk reduce-add): start from y = x
        if (d->ld_in != d->ld_out) return fail("srk_swin_attn_fwd: out-of-place residual needs ld_in == ld_out");
        const size_t bytes = static_cast<size_t>(p.total_windows) * 64 * d->ld_in * sizeof(float);
        cudaError_t e = cudaMemcpyAsync(y, x, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) return fail("srk_swin_attn_fwd: %s", cudaGetErrorString(e));
    }
    if (sync && sync->progress) {
        if (d->mode != SRK_MODE_IMAGE || (p.nw_img & 1) || sync->batch != d->batch || sync->wait_target < 0)
            return fail("srk_swin_attn_fwd_sync: progress counters need SRK_MODE_IMAGE, an even number of windows per image and batch == desc->batch");
        if (sync->wait_target > 0 && x != y) return fail("srk_swin_attn_fwd_sync: wait_target > 0 needs the in-place form (x == y)");
        p.prog_sig = sync->progress;
        p.prog_wait = sync->progress + sync->batch;
        p.wait_target = sync->wait_target;
    }
    p.dbg = srk::g_timeline;
    p.stagger = p.n_tiles >= 2 * 148 ? srk::g_stagger_attn : 0;
    if (d->operands == SRK_OPERANDS_F16) return check(srk::launch_swin_attn_f16(p, static_cast<cudaStream_t>(stream)), "srk_swin_attn_fwd");
    if (d->operands != SRK_OPERANDS_BF16) return fail("srk_swin_attn_fwd: unknown operands %d", d->operands);
    return check(srk::launch_swin_attn(p, static_cast<cudaStream_t>(stream)), "srk_swin_attn_fwd");
}

int srk_swin_mlp_fwd(const SrkMlpDesc* d, const float* x, float* y, const void* wstream, const float* vec, void* stream) {
    return srk_swin_mlp_fwd_sync(d, x, y, wstream, vec, nullptr, stream);
}

int srk_swin_mlp_fwd_sync(const SrkMlpDesc* d, const float* x, float* y, const void* wstream, const float* vec, const SrkBlockSync* sync,
                          void* stream) {
    if (!d || !x || !y || !wstream || !vec) return fail("srk_swin_mlp_fwd: null argument");
    if (d->ld_in < SRK_DIM || d->ld_out < SRK_DIM || (d->ld_in & 3) || (d->ld_out & 3))
        return fail("srk_swin_mlp_fwd: ld_in/ld_out must be >= %d and multiples of 4 (got %d, %d)", SRK_DIM, d->ld_in, d->ld_out);
    if (!aligned16(x) || !aligned16(y) || !aligned16(wstream) || !aligned16(vec))
        return fail("srk_swin_mlp_fwd: pointers must be 16-byte aligned");
    if (d->num_tokens <= 0 || d->num_tokens > (1ll << 36)) return fail("srk_swin_mlp_fwd: bad num_tokens");
    srk::MlpParams p{};
    p.x = x; p.y = y; p.wstream = static_cast<const uint8_t*>(wstream); p.vec = vec;
    p.num_tokens = d->num_tokens; p.n_tiles = static_cast<int>((d->num_tokens + 127) / 128);
    p.ld_in = d->ld_in; p.ld_out = d->ld_out; p.apply_ln = d->apply_ln; p.add_residual = d->add_residual;
    if (d->add_residual && x != y) {
        if (d->ld_in != d->ld_out) return fail("srk_swin_mlp_fwd: out-of-place residual needs ld_in == ld_out");
        cudaError_t e = cudaMemcpyAsync(y, x, static_cast<size_t>(d->num_tokens) * d->ld_in * sizeof(float), cudaMemcpyDeviceToDevice,
                                        static_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) return fail("srk_swin_mlp_fwd: %s", cudaGetErrorString(e));
    }
    if (sync && sync->progress) {
        if (sync->tokens_per_image <= 0 || (sync->tokens_per_image % 128) || sync->wait_target < 0 ||
            static_cast<int64_t>(sync->batch) * sync->tokens_per_image != d->num_tokens)
            return fail("srk_swin_mlp_fwd_sync: progress counters need tokens_per_image %% 128 == 0 and batch * tokens_per_image == num_tokens");
        if (sync->wait_target > 0 && x != y) return fail("srk_swin_mlp_fwd_sync: wait_target > 0 needs the in-place form (x == y)");
        p.prog_sig = sync->progress + sync->batch;
        p.prog_wait = sync->progress;
        p.wait_target = sync->wait_target;
        p.tokens_per_image = sync->tokens_per_image;
    } else {
        p.tokens_per_image = 128;
    }
    p.dbg = srk::g_timeline;
    p.stagger = p.n_tiles >= 2 * 148 ? srk::g_stagger_mlp : 0;
    if (d->operands == SRK_OPERANDS_F16) return check(srk::launch_swin_mlp_f16(p, static_cast<cudaStream_t>(stream)), "srk_swin_mlp_fwd");
    if (d->operands == SRK_OPERANDS_F16_HALF_GELU) return check(srk::launch_swin_mlp_f16h(p, static_cast<cudaStream_t>(stream)), "srk_swin_mlp_fwd");
    if (d->operands != SRK_OPERANDS_BF16) return fail("srk_swin_mlp_fwd: unknown operands %d", d->operands);
    return check(srk::launch_swin_mlp(p, static_cast<cudaStream_t>(stream)), "srk_swin_mlp_fwd");
}

int srk_swin_layer_fwd(const SrkLayerDesc* d, float* y, const SrkLayerBlock* blocks, int32_t* progress, void* stream) {
    if (!d || !y || !blocks || !progress) return fail("srk_swin_layer_fwd: null argument");
    if (d->batch <= 0 || d->height <= 0 || d->width <= 0 || d->height % SRK_WINDOW || d->width % SRK_WINDOW)
        return fail("srk_swin_layer_fwd: height/width must be positive multiples of %d (got %d x %d)", SRK_WINDOW, d->height, d->width);
    const int64_t tokens_img = static_cast<int64_t>(d->height) * d->width;
    if (tokens_img % 128) return fail("srk_swin_layer_fwd: height * width must be a multiple of 128 (got %lld)", static_cast<long long>(tokens_img));
    if (d->n_blocks < 1 || d->n_blocks > SRK_LAYER_MAX_BLOCKS) return fail("srk_swin_layer_fwd: n_blocks must be 1..%d", SRK_LAYER_MAX_BLOCKS);
    if (d->ld < SRK_DIM || (d->ld & 3) || !aligned16(y)) return fail("srk_swin_layer_fwd: ld must be >= %d and a multiple of 4, y 16-byte aligned", SRK_DIM);
    const int64_t T = d->batch * tokens_img / 128;
    if (T * 2 * d->n_blocks >= (int64_t(1) << 30)) return fail("srk_swin_layer_fwd: too many tiles");
    srk::LayerParams p{};
    p.y = y; p.ld = d->ld; p.B = d->batch; p.H = d->height; p.W = d->width;
    p.nwx = d->width / SRK_WINDOW; p.nw_img = (d->height / SRK_WINDOW) * p.nwx;
    p.T = static_cast<int>(T); p.tiles_per_image = static_cast<int>(tokens_img / 128);
    p.n_blocks = d->n_blocks; p.n_items = static_cast<int>(T * 2 * d->n_blocks);
    p.progress = progress;
    for (int b = 0; b < d->n_blocks; ++b) {
        const SrkLayerBlock& k = blocks[b];
        if (!k.attn_wstream || !k.attn_vec || !k.mlp_wstream || !k.mlp_vec) return fail("srk_swin_layer_fwd: block %d has a null pointer", b);
        if (!aligned16(k.attn_wstream) || !aligned16(k.attn_vec) || !aligned16(k.mlp_wstream) || !aligned16(k.mlp_vec))
            return fail("srk_swin_layer_fwd: block %d: pointers must be 16-byte aligned", b);
        if (k.shift != 0 && k.shift != SRK_WINDOW / 2) return fail("srk_swin_layer_fwd: block %d: shift must be 0 or %d", b, SRK_WINDOW / 2);
        p.blk[b].attn_w = static_cast<const uint8_t*>(k.attn_wstream); p.blk[b].attn_vec = k.attn_vec;
        p.blk[b].mlp_w = static_cast<const uint8_t*>(k.mlp_wstream); p.blk[b].mlp_vec = k.mlp_vec; p.blk[b].shift = k.shift;
    }
    p.dbg = srk::g_timeline;
    cudaError_t e = cudaMemsetAsync(progress, 0, sizeof(int32_t) * 2 * d->batch, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail("srk_swin_layer_fwd: %s", cudaGetErrorString(e));
    return check(srk::launch_swin_layer(p, static_cast<cudaStream_t>(stream)), "srk_swin_layer_fwd");
}

int srk_linear_fwd(const SrkLinearDesc* d, const void* a, const void* wstream, const float* bias, void* out, void* stream) {
    if (!d || !a || !wstream || !bias || !out) return fail("srk_linear_fwd: null argument");
    if (!aligned16(a) || !aligned16(wstream) || !aligned16(bias) || !aligned16(out)) return fail("srk_linear_fwd: pointers must be 16-byte aligned");
    if (d->num_tokens <= 0 || d->num_tokens > (1ll << 31)) return fail("srk_linear_fwd: bad num_tokens");
    if (d->k_atoms != 3 && d->k_atoms != 6) return fail("srk_linear_fwd: k_atoms must be 3 or 6 (got %d)", d->k_atoms);
    if (d->n_chunks < 1 || d->n_chunks > SRK_LIN_MAX_CHUNKS) return fail("srk_linear_fwd: n_chunks must be 1..%d", SRK_LIN_MAX_CHUNKS);
    srk::LinearParams p{};
    p.a_mode = d->a_mode; p.k_atoms = d->k_atoms; p.num_tokens = d->num_tokens;
    p.n_tiles = static_cast<int>((d->num_tokens + 127) / 128);
    p.wstream = static_cast<const uint8_t*>(wstream); p.bias = bias; p.n_chunks = d->n_chunks; p.act = d->act;
    p.out_mode = d->out_mode; p.plane_phase_mask = d->plane_phase_mask;
    if (d->a_mode == SRK_LIN_A_ROWS) {
        if (d->k_atoms != 3) return fail("srk_linear_fwd: fp32 row input needs k_atoms = 3 (K = 180)");
        if (d->ld_in < SRK_DIM || (d->ld_in & 3)) return fail("srk_linear_fwd: bad ld_in %d", d->ld_in);
        p.x = static_cast<const float*>(a); p.ld_in = d->ld_in; p.apply_ln = d->apply_ln;
    } else if (d->a_mode == SRK_LIN_A_PLANES) {
        p.a_planes = static_cast<const uint8_t*>(a); p.a_plane_stride = d->num_tokens * 128;
    } else {
        return fail("srk_linear_fwd: unknown a_mode %d", d->a_mode);
    }
    if (d->out_mode == SRK_LIN_OUT_PLANES) {
        p.out_planes = static_cast<uint8_t*>(out); p.out_plane_stride = d->num_tokens * 128;
    } else if (d->out_mode == SRK_LIN_OUT_ROWS) {
        if (d->k_atoms == 6 && d->n_chunks != 1) return fail("srk_linear_fwd: fp32 row output with k_atoms = 6 needs n_chunks = 1");
        if (d->ld_out < SRK_DIM * d->n_chunks || (d->ld_out & 3)) return fail("srk_linear_fwd: bad ld_out %d", d->ld_out);
        p.y = static_cast<float*>(out); p.ld_out = d->ld_out; p.add_residual = d->add_residual;
    } else {
        return fail("srk_linear_fwd: unknown out_mode %d", d->out_mode);
    }
    p.dbg = srk::g_timeline;
    return check(srk::launch_token_linear(p, static_cast<cudaStream_t>(stream)), "srk_linear_fwd");
}

int srk_window_attention_table_floats(int32_t kind) { return srk::winattn_table_floats(kind); }

int srk_window_attention_fwd(const SrkWinAttnDesc* d, const void* q_planes, const void* k_planes, const void* v_planes,
                             const float* bias_table, const float* emask, const void* pad_pages, void* out, void* stream) {
    const void* zero_page = pad_pages;
    if (!d || !q_planes || !k_planes || !v_planes || !bias_table || !zero_page || !out) return fail("srk_window_attention_fwd: null argument");
    if (!aligned16(q_planes) || !aligned16(k_planes) || !aligned16(v_planes) || !aligned16(bias_table) || !aligned16(zero_page) || !aligned16(out))
        return fail("srk_window_attention_fwd: pointers must be 16-byte aligned");
    int qh, qw, sy, c0, koff = 0, wrap = 1;
    switch (d->kind) {
        case SRK_WA_HAT_WMSA: qh = 16; qw = 16; sy = 48; c0 = 15 * sy + 15; break;
        case SRK_WA_HAT_OCAB: qh = 16; qw = 16; sy = 48; c0 = 23 * sy + 23; koff = -4; wrap = 0; break;
        case SRK_WA_DAT_8x32: qh = 8; qw = 32; sy = 64; c0 = 7 * sy + 31; break;
        case SRK_WA_DAT_32x8: qh = 32; qw = 8; sy = 24; c0 = 31 * sy + 7; break;
        default: return fail("srk_window_attention_fwd: unknown kind %d", d->kind);
    }
    if (d->batch <= 0 || d->height <= 0 || d->width <= 0 || d->height % qh || d->width % qw || d->width % 8)
        return fail("srk_window_attention_fwd: height/width (%d x %d) must be positive multiples of the %d x %d window", d->height, d->width, qh, qw);
    if (d->n_heads < 1 || d->n_heads > SRK_HEADS) return fail("srk_window_attention_fwd: n_heads must be 1..%d", SRK_HEADS);
    if (d->shift_y < 0 || d->shift_y >= qh || d->shift_x < 0 || d->shift_x >= qw) return fail("srk_window_attention_fwd: shift out of range");
    if (!wrap && (d->shift_y || d->shift_x || d->mask_shift || emask)) return fail("srk_window_attention_fwd: OCAB takes no shift / mask");
    if (d->mask_shift && (d->shift_y == 0 || d->shift_x == 0)) return fail("srk_window_attention_fwd: mask_shift needs a non-zero shift");
    // operand-row phase: (x + phase) & 7 must equal the image row & 7 of every window row -> shift_x must be a multiple of 4
    if (d->shift_x & 3) return fail("srk_window_attention_fwd: shift_x must be a multiple of 4");
    if (emask && (d->emask_nw <= 0 || !aligned16(emask))) return fail("srk_window_attention_fwd: explicit mask needs emask_nw > 0");
    const int64_t tokens = static_cast<int64_t>(d->batch) * d->height * d->width;
    if (tokens > (1ll << 31)) return fail("srk_window_attention_fwd: too many tokens");
    srk::WinAttnParams p{};
    p.q_planes = static_cast<const uint8_t*>(q_planes); p.k_planes = static_cast<const uint8_t*>(k_planes);
    p.v_planes = static_cast<const uint8_t*>(v_planes); p.plane_stride = tokens * 128;
    p.tab = bias_table; p.c0 = c0; p.zero_page = static_cast<const uint8_t*>(zero_page);
    p.B = d->batch; p.H = d->height; p.W = d->width; p.koff = koff; p.shift_y = d->shift_y; p.shift_x = d->shift_x;
    p.wrap = wrap; p.mask_shift = d->mask_shift; p.emask = emask; p.emask_nw = emask ? d->emask_nw : 1;
    p.n_heads = d->n_heads; p.nwy = d->height / qh; p.nwx = d->width / qw;
    const int64_t items = static_cast<int64_t>(d->batch) * p.nwy * p.nwx * ((d->n_heads + 1) / 2);
    if (items > (1ll << 30)) return fail("srk_window_attention_fwd: too many windows");
    p.n_items = static_cast<int>(items);
    p.out_mode = d->out_mode;
    if (d->out_mode == 0) {
        p.o_planes = static_cast<uint8_t*>(out); p.o_plane_stride = tokens * 128;
    } else if (d->out_mode == 1) {
        if (d->out_ld < SRK_HEAD_DIM * d->n_heads + d->out_col0 || (d->out_ld & 1) || (d->out_col0 & 1)) return fail("srk_window_attention_fwd: bad out_ld / out_col0");
        p.o_rows = static_cast<float*>(out); p.o_ld = d->out_ld; p.o_col0 = d->out_col0;
    } else {
        return fail("srk_window_attention_fwd: unknown out_mode %d", d->out_mode);
    }
    p.dbg = srk::g_timeline;
    p.stagger = srk::g_stagger_winattn;
    return check(srk::launch_winattn(d->kind, p, static_cast<cudaStream_t>(stream)), "srk_window_attention_fwd");
}

int srk_layernorm_fwd(const float* x, float* y, const float* w, const float* b, int64_t num_tokens, int32_t ld_in,
                      int32_t ld_out, void* stream) {
    if (!x || !y || !w || !b) return fail("srk_layernorm_fwd: null argument");
    if (ld_in < SRK_DIM || ld_out < SRK_DIM || (ld_in & 3) || (ld_out & 3)) return fail("srk_layernorm_fwd: bad ld");
    if (!aligned16(x) || !aligned16(y) || !aligned16(w) || !aligned16(b)) return fail("srk_layernorm_fwd: pointers must be 16-byte aligned");
    if (num_tokens < 0) return fail("srk_layernorm_fwd: bad num_tokens");
    return check(srk::launch_layernorm(x, y, nullptr, w, b, num_tokens, ld_in, ld_out, static_cast<cudaStream_t>(stream)), "srk_layernorm_fwd");
}

int srk_layernorm_f16_fwd(const float* x, float* y, void* y16, const float* w, const float* b, int64_t num_tokens, int32_t ld_in,
                          int32_t ld_out, void* stream) {
    if (!x || !y16 || !w || !b) return fail("srk_layernorm_f16_fwd: null argument");
    if (ld_in < SRK_DIM || (ld_in & 3) || (y && (ld_out < SRK_DIM || (ld_out & 3)))) return fail("srk_layernorm_f16_fwd: bad ld");
    if (!aligned16(x) || !aligned16(y16) || !aligned16(w) || !aligned16(b) || (y && !aligned16(y))) return fail("srk_layernorm_f16_fwd: pointers must be 16-byte aligned");
    if (num_tokens < 0) return fail("srk_layernorm_f16_fwd: bad num_tokens");
    return check(srk::launch_layernorm(x, y, static_cast<__half*>(y16), w, b, num_tokens, ld_in, ld_out, static_cast<cudaStream_t>(stream)),
                 "srk_layernorm_f16_fwd");
}

int srk_cab_gate_add(const float* y, const float* y_bias, float* out, float* sums_ws, const float* w1, const float* b1, const float* w2, const float* b2,
                     int32_t hidden, float scale, int32_t batch, int32_t tokens_per_image, void* stream) {
    if (!y || !out || !sums_ws || !w1 || !b1 || !w2 || !b2) return fail("srk_cab_gate_add: null argument");
    if (!aligned16(y) || !aligned16(out)) return fail("srk_cab_gate_add: y / out must be 16-byte aligned");
    if (hidden < 1 || hidden > 32) return fail("srk_cab_gate_add: hidden must be 1..32 (got %d)", hidden);
    if (batch < 0 || batch > 65535 || tokens_per_image < 0) return fail("srk_cab_gate_add: bad shape");
    cudaError_t e = srk::launch_cab_gate_add(y, y_bias, out, sums_ws, w1, b1, w2, b2, hidden, scale, batch, tokens_per_image,
                                             static_cast<cudaStream_t>(stream));
    if (e == cudaSuccess) g_launches.fetch_add(1, std::memory_order_relaxed);      // two kernels: check() below counts the second
    return check(e, "srk_cab_gate_add");
}

int srk_dwconv3x3_rows_fwd(const float* in, int32_t ld_in, int32_t c_in, const float* w9c, const float* scale, const float* shift,
                           const float* ln_stats, const float* ln_gamma, const float* ln_beta, const float* gate, int32_t ld_gate,
                           int32_t c_gate, float* out, int32_t ld_out, int32_t channels, int32_t batch, int32_t height, int32_t width,
                           int32_t act_gelu, void* stream) {
    if (!in || !w9c || !scale || !shift || !out) return fail("srk_dwconv3x3_rows_fwd: null argument");
    if (ln_stats && (!ln_gamma || !ln_beta)) return fail("srk_dwconv3x3_rows_fwd: LayerNorm input needs gamma and beta");
    if (channels <= 0 || (channels & 3) || (ld_in & 3) || (c_in & 3) || (ld_out & 3) || c_in + channels > ld_in || channels > ld_out)
        return fail("srk_dwconv3x3_rows_fwd: channels / ld / offsets must be multiples of 4 and consistent");
    if (gate && ((ld_gate & 3) || (c_gate & 3) || c_gate + channels > ld_gate)) return fail("srk_dwconv3x3_rows_fwd: bad gate slice");
    if (!aligned16(in) || !aligned16(w9c) || !aligned16(scale) || !aligned16(shift) || !aligned16(out) || (gate && !aligned16(gate)) ||
        (ln_stats && (!aligned16(ln_gamma) || !aligned16(ln_beta))))
        return fail("srk_dwconv3x3_rows_fwd: pointers must be 16-byte aligned");
    if (batch < 0 || height <= 0 || width <= 0) return fail("srk_dwconv3x3_rows_fwd: bad shape");
    return check(srk::launch_dwconv3x3_rows(in, ld_in, c_in, w9c, scale, shift, ln_stats, ln_gamma, ln_beta, gate, ld_gate, c_gate, out,
                                            ld_out, channels, batch, height, width, act_gelu, static_cast<cudaStream_t>(stream)),
                 "srk_dwconv3x3_rows_fwd");
}

int srk_dwconv3x3_rows_planes_fwd(const float* in, int32_t ld_in, int32_t c_in, const float* w9c, const float* scale, const float* shift,
                                  const float* ln_stats, const float* ln_gamma, const float* ln_beta, const float* gate, int32_t ld_gate,
                                  int32_t c_gate, void* out_planes, int64_t plane_stride, int32_t channels, int32_t batch, int32_t height,
                                  int32_t width, int32_t act_gelu, void* stream) {
    if (!in || !w9c || !scale || !shift || !out_planes) return fail("srk_dwconv3x3_rows_planes_fwd: null argument");
    if (ln_stats && (!ln_gamma || !ln_beta)) return fail("srk_dwconv3x3_rows_planes_fwd: LayerNorm input needs gamma and beta");
    if (channels <= 0 || (channels & 3) || (ld_in & 3) || (c_in & 3) || c_in + channels > ld_in)
        return fail("srk_dwconv3x3_rows_planes_fwd: channels / ld / offsets must be multiples of 4 and consistent");
    if (gate && ((ld_gate & 3) || (c_gate & 3) || c_gate + channels > ld_gate)) return fail("srk_dwconv3x3_rows_planes_fwd: bad gate slice");
    if (!aligned16(in) || !aligned16(w9c) || !aligned16(scale) || !aligned16(shift) || !aligned16(out_planes) || (gate && !aligned16(gate)) ||
        (ln_stats && (!aligned16(ln_gamma) || !aligned16(ln_beta))) || (plane_stride & 127))
        return fail("srk_dwconv3x3_rows_planes_fwd: pointers must be 16-byte aligned, plane_stride a multiple of 128");
    if (batch < 0 || height <= 0 || width <= 0 || plane_stride < static_cast<int64_t>(batch) * height * width * 128)
        return fail("srk_dwconv3x3_rows_planes_fwd: bad shape");
    return check(srk::launch_dwconv3x3_rows(in, ld_in, c_in, w9c, scale, shift, ln_stats, ln_gamma, ln_beta, gate, ld_gate, c_gate, nullptr, 0,
                                            channels, batch, height, width, act_gelu, static_cast<cudaStream_t>(stream),
                                            static_cast<uint8_t*>(out_planes), plane_stride),
                 "srk_dwconv3x3_rows_planes_fwd");
}

int srk_row_stats_fwd(const float* in, int32_t ld_in, int32_t c_in, int32_t channels, int64_t tokens, float eps, float* stats, void* stream) {
    if (!in || !stats) return fail("srk_row_stats_fwd: null argument");
    if (channels <= 0 || (channels & 3) || (ld_in & 3) || (c_in & 3) || c_in + channels > ld_in || !aligned16(in))
        return fail("srk_row_stats_fwd: channels / ld / offset must be multiples of 4, 16-byte aligned rows");
    if (tokens < 0) return fail("srk_row_stats_fwd: bad tokens");
    return check(srk::launch_row_stats(in, ld_in, c_in, channels, tokens, eps, stats, static_cast<cudaStream_t>(stream)), "srk_row_stats_fwd");
}

int srk_dat_mix_fwd(const float* att, const float* conv, const float* cmap, const float* w1, const float* b1, const float* w2, float b2,
                    int32_t hidden, int32_t mode, float* mix, int64_t tokens, int32_t tokens_per_image, void* stream) {
    if (!att || !conv || !cmap || !w1 || !b1 || !w2 || !mix) return fail("srk_dat_mix_fwd: null argument");
    if (!aligned16(att) || !aligned16(conv) || !aligned16(cmap) || !aligned16(mix)) return fail("srk_dat_mix_fwd: pointers must be 16-byte aligned");
    if (hidden < 1 || hidden > 16 || (mode != 0 && mode != 1) || tokens < 0 || tokens_per_image <= 0) return fail("srk_dat_mix_fwd: bad arguments");
    return check(srk::launch_dat_mix(att, conv, cmap, w1, b1, w2, b2, hidden, mode, mix, tokens, tokens_per_image,
                                     static_cast<cudaStream_t>(stream)), "srk_dat_mix_fwd");
}

int srk_token_mean_mlp_fwd(const float* x, float* out, float* sums_ws, const float* w1, const float* b1, const float* w2, const float* b2,
                           int32_t hidden, int32_t batch, int32_t tokens_per_image, void* stream) {
    if (!x || !out || !sums_ws || !w1 || !b1 || !w2 || !b2) return fail("srk_token_mean_mlp_fwd: null argument");
    if (batch < 0 || batch > 65535 || tokens_per_image < 0 || hidden < 1 || hidden > 64) return fail("srk_token_mean_mlp_fwd: bad shape");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return check(srk::launch_token_mean_mlp(x, out, sums_ws, w1, b1, w2, b2, hidden, batch, tokens_per_image, static_cast<cudaStream_t>(stream)),
                 "srk_token_mean_mlp_fwd");
}

int srk_token_mean_fwd(const float* x, float* mean, float* sums_ws, int32_t batch, int32_t tokens_per_image, void* stream) {
    if (!x || !mean || !sums_ws) return fail("srk_token_mean_fwd: null argument");
    if (batch < 0 || batch > 65535 || tokens_per_image < 0) return fail("srk_token_mean_fwd: bad shape");
    return check(srk::launch_token_mean(x, mean, sums_ws, batch, tokens_per_image, static_cast<cudaStream_t>(stream)), "srk_token_mean_fwd");
}

int srk_dat_channel_gram_ws_floats(int32_t batch, int32_t tokens_per_image) { return srk::channel_gram_ws_floats(batch, tokens_per_image); }
int srk_cab_ws_floats(int32_t batch, int32_t tokens_per_image) { return srk::cab_ws_floats(batch, tokens_per_image); }

int srk_dat_channel_gram_fwd(const float* qkv, float* gram, float* ws, int32_t batch, int32_t tokens_per_image, void* stream) {
    if (!qkv || !gram || !ws) return fail("srk_dat_channel_gram_fwd: null argument");
    if (!aligned16(qkv)) return fail("srk_dat_channel_gram_fwd: qkv must be 16-byte aligned");
    if (batch < 0 || batch > 65535 || tokens_per_image <= 0) return fail("srk_dat_channel_gram_fwd: bad shape");
    return check(srk::launch_channel_gram(qkv, gram, ws, batch, tokens_per_image, static_cast<cudaStream_t>(stream)), "srk_dat_channel_gram_fwd");
}

int srk_dat_channel_softmax_fwd(const float* gram, const float* temperature, float* attn, int32_t batch, void* stream) {
    if (!gram || !temperature || !attn) return fail("srk_dat_channel_softmax_fwd: null argument");
    if (batch < 0 || batch > 65535) return fail("srk_dat_channel_softmax_fwd: bad shape");
    return check(srk::launch_channel_softmax(gram, temperature, attn, batch, static_cast<cudaStream_t>(stream)), "srk_dat_channel_softmax_fwd");
}

int srk_dat_channel_apply_fwd(const float* qkv, const float* attn, float* out, int32_t batch, int32_t tokens_per_image, void* stream) {
    if (!qkv || !attn || !out) return fail("srk_dat_channel_apply_fwd: null argument");
    if (batch < 0 || batch > 65535 || tokens_per_image <= 0) return fail("srk_dat_channel_apply_fwd: bad shape");
    return check(srk::launch_channel_apply(qkv, attn, out, batch, tokens_per_image, static_cast<cudaStream_t>(stream)), "srk_dat_channel_apply_fwd");
}

int srk_pixelshuffle_nhwc_bias_fwd(const float* x, const float* bias, float* y, int32_t batch, int32_t height, int32_t width,
                                   int32_t out_channels, int32_t r, void* stream) {
    if (!x || !y) return fail("srk_pixelshuffle_nhwc_fwd: null argument");
    if (batch < 0 || height < 0 || width < 0 || out_channels <= 0 || r <= 0) return fail("srk_pixelshuffle_nhwc_fwd: bad shape");
    if (r == 2 && (!aligned16(x) || (bias && !aligned16(bias)))) return fail("srk_pixelshuffle_nhwc_fwd: x / bias must be 16-byte aligned");
    return check(srk::launch_pixelshuffle_nhwc(x, bias, y, batch, height, width, out_channels, r, static_cast<cudaStream_t>(stream)),
                 "srk_pixelshuffle_nhwc_fwd");
}
int srk_pixelshuffle_nhwc_fwd(const float* x, float* y, int32_t batch, int32_t height, int32_t width, int32_t out_channels,
                              int32_t r, void* stream) {
    return srk_pixelshuffle_nhwc_bias_fwd(x, nullptr, y, batch, height, width, out_channels, r, stream);
}

int srk_bias_act_add_nhwc(const float* x, const float* bias, const float* residual, float* y, int64_t pixels, int32_t channels,
                          int32_t act, float slope, void* stream) {
    if (!x || !y) return fail("srk_bias_act_add_nhwc: null argument");
    if (pixels < 0 || channels <= 0) return fail("srk_bias_act_add_nhwc: bad shape");
    if (act != SRK_ACT_NONE && act != SRK_ACT_LEAKY_RELU && act != SRK_ACT_GELU) return fail("srk_bias_act_add_nhwc: unknown activation");
    return check(srk::launch_bias_act_add(x, bias, residual, y, pixels, channels, act, slope, static_cast<cudaStream_t>(stream)),
                 "srk_bias_act_add_nhwc");
}

int srk_stitch_accumulate(const float* tiles, float* E, float* Wt, const int32_t* tile_yx, int32_t num_tiles, int32_t channels,
                          int32_t tile_h, int32_t tile_w, int32_t out_h, int32_t out_w, void* stream) {
    if (!tiles || !E || !Wt || !tile_yx) return fail("srk_stitch_accumulate: null argument");
    if (num_tiles < 0 || num_tiles > 65535 || channels <= 0 || tile_h <= 0 || tile_w <= 0) return fail("srk_stitch_accumulate: bad shape");
    return check(srk::launch_stitch_accumulate(tiles, E, Wt, tile_yx, num_tiles, channels, tile_h, tile_w, out_h, out_w,
                                               static_cast<cudaStream_t>(stream)), "srk_stitch_accumulate");
}

int srk_gather_tiles(const float* slab, int64_t slab_cstride, int32_t slab_rows, int32_t slab_w, const int32_t* src_yx, int32_t num_tiles,
                     int32_t channels, int32_t tile_h, int32_t tile_w, float* out, void* stream) {
    if (!slab || !src_yx || !out) return fail("srk_gather_tiles: null argument");
    if (num_tiles < 0 || num_tiles > 65535 || channels <= 0 || channels > 65535 || tile_h <= 0 || tile_w <= 0 || tile_h > slab_rows || tile_w > slab_w)
        return fail("srk_gather_tiles: bad shape");
    return check(srk::launch_gather_tiles(slab, slab_cstride, slab_w, src_yx, num_tiles, channels, tile_h, tile_w, out,
                                          static_cast<cudaStream_t>(stream)), "srk_gather_tiles");
}

int srk_stitch_accumulate_strided(const float* tiles, int64_t stride_n, int64_t stride_c, int64_t stride_y, int64_t stride_x, float* E,
                                  int64_t e_channel_stride, const int32_t* dst_yx, int32_t num_tiles, int32_t channels, int32_t tile_h,
                                  int32_t tile_w, int32_t out_h, int32_t out_w, void* stream) {
    if (!tiles || !E || !dst_yx) return fail("srk_stitch_accumulate_strided: null argument");
    if (num_tiles < 0 || num_tiles > 65535 || channels <= 0 || tile_h <= 0 || tile_w <= 0 || out_h <= 0 || out_w <= 0)
        return fail("srk_stitch_accumulate_strided: bad shape");
    return check(srk::launch_stitch_accumulate2(tiles, stride_n, stride_c, stride_y, stride_x, E, e_channel_stride, dst_yx, num_tiles, channels,
                                                tile_h, tile_w, out_h, out_w, static_cast<cudaStream_t>(stream)), "srk_stitch_accumulate_strided");
}

int srk_stitch_finalize(const float* E, int64_t e_channel_stride, const float* cnt_y, const float* cnt_x, void* out, int64_t out_channel_stride,
                        int32_t out_dtype, int32_t channels, int32_t out_h, int32_t out_w, void* stream) {
    if (!E || !cnt_y || !cnt_x || !out) return fail("srk_stitch_finalize: null argument");
    if (out_dtype < SRK_OUT_F32 || out_dtype > SRK_OUT_U8 || channels <= 0 || out_h <= 0 || out_w <= 0) return fail("srk_stitch_finalize: bad arguments");
    return check(srk::launch_stitch_finalize(E, e_channel_stride, cnt_y, cnt_x, out, out_channel_stride, out_dtype, channels, out_h, out_w,
                                             static_cast<cudaStream_t>(stream)), "srk_stitch_finalize");
}

int srk_conv3x3_fwd(const SrkConvDesc* d, const void* in_f16, const void* wstream, const float* bias, const float* residual, void* out,
                    void* stream) {
    if (!d || !in_f16 || !wstream || !bias || !out) return fail("srk_conv3x3_fwd: null argument");
    if (d->batch <= 0 || d->height <= 0 || d->width <= 0) return fail("srk_conv3x3_fwd: bad image shape");
    const int a_atoms = d->a_atoms > 0 ? d->a_atoms : d->k_atoms;
    if (d->k_atoms < 1 || d->k_atoms > 12 || a_atoms > 8 || a_atoms > d->k_atoms || d->k_atoms > 2 * a_atoms)
        return fail("srk_conv3x3_fwd: need 1 <= a_atoms <= 8 and a_atoms <= k_atoms <= min(12, 2 a_atoms) (got k_atoms %d, a_atoms %d)", d->k_atoms, d->a_atoms);
    if (!aligned16(in_f16) || !aligned16(wstream) || !aligned16(out) || (residual && !aligned16(residual)))
        return fail("srk_conv3x3_fwd: pointers must be 16-byte aligned");
    if (static_cast<int64_t>(d->batch) * d->height * d->width * 4 >= (int64_t(1) << 31)) return fail("srk_conv3x3_fwd: too many pixels");
    srk::ConvArgs a{};
    a.in = static_cast<const __half*>(in_f16); a.wstream = static_cast<const uint8_t*>(wstream); a.bias = bias; a.residual = residual;
    a.B = d->batch; a.H = d->height; a.W = d->width; a.k_atoms = d->k_atoms; a.np = d->np; a.cout = d->cout;
    a.out_mode = d->out_mode; a.ld_out = d->ld_out; a.act = d->act; a.slope = d->slope; a.a_atoms = a_atoms;
    switch (d->out_mode) {
        case SRK_CONV_OUT_ROWS_F32:
            if (d->np % 32 || d->np < 32 || d->np > 256 || d->cout > d->np || d->cout % 4 || d->cout <= 0 || d->ld_out < d->cout || d->ld_out % 4)
                return fail("srk_conv3x3_fwd: rows output needs np %% 32 == 0 <= 256, cout %% 4 == 0 <= np, ld_out %% 4 == 0 >= cout");
            a.out_f32 = static_cast<float*>(out);
            break;
        case SRK_CONV_OUT_NHWC_F16:
            if (d->np % 32 || d->np < 32 || d->np > 256 || d->ld_out < d->np || d->ld_out % 8 || residual)
                return fail("srk_conv3x3_fwd: fp16 NHWC output needs np %% 32 == 0 <= 256, ld_out %% 8 == 0 >= np, no residual");
            a.out_f16 = static_cast<__half*>(out);
            break;
        case SRK_CONV_OUT_SHUFFLE2_F16:
            if (d->np != 256 || residual) return fail("srk_conv3x3_fwd: pixel-shuffle output needs np == 256 (4 x 64 channels), no residual");
            a.out_f16 = static_cast<__half*>(out);
            break;
        case SRK_CONV_OUT_IMAGE:
            if (d->np != 16 || d->cout < 1 || d->cout > 4 || d->ld_out < d->cout) return fail("srk_conv3x3_fwd: image output needs np == 16, cout <= 4");
            a.out_f32 = static_cast<float*>(out);
            break;
        default:
            return fail("srk_conv3x3_fwd: unknown out_mode %d", d->out_mode);
    }
    if (d->act < SRK_ACT_NONE || d->act > SRK_ACT_GELU) return fail("srk_conv3x3_fwd: unknown activation %d", d->act);
    cudaError_t e = srk::launch_conv3x3(a, static_cast<cudaStream_t>(stream));
    if (e == cudaErrorNotSupported) return fail("srk_conv3x3_fwd: cuTensorMapEncodeTiled is not available from this driver");
    return check(e, "srk_conv3x3_fwd");
}

int srk_rows_to_f16(const float* x, int32_t ld_in, int32_t channels, void* out_f16, int32_t cp, int64_t pixels, void* stream) {
    if (!x || !out_f16) return fail("srk_rows_to_f16: null argument");
    if (channels <= 0 || ld_in < channels || (ld_in & 3) || cp < channels || cp % 64 || !aligned16(x) || !aligned16(out_f16))
        return fail("srk_rows_to_f16: need ld_in %% 4 == 0 >= channels, cp %% 64 == 0 >= channels, 16-byte aligned pointers");
    return check(srk::launch_rows_to_f16(x, ld_in, channels, static_cast<__half*>(out_f16), cp, pixels, static_cast<cudaStream_t>(stream)), "srk_rows_to_f16");
}

int srk_rows_to_f16_split(const float* x, int32_t ld_in, int32_t channels, void* hi_f16, void* lo_f16, int32_t ld_out, int32_t cp,
                          int64_t pixels, int32_t act, float slope, int32_t shuffle_h, int32_t shuffle_w, void* stream) {
    if (!x || !hi_f16 || !lo_f16) return fail("srk_rows_to_f16_split: null argument");
    if (channels <= 0 || ld_in < channels || (ld_in & 3) || cp % 64 || cp <= 0 || ld_out < cp || ld_out % 8 || !aligned16(x) || !aligned16(hi_f16) ||
        !aligned16(lo_f16))
        return fail("srk_rows_to_f16_split: need ld_in %% 4 == 0 >= channels, cp %% 64 == 0, ld_out %% 8 == 0 >= cp, 16-byte aligned pointers");
    if (shuffle_h > 0 || shuffle_w > 0) {
        if (shuffle_h <= 0 || shuffle_w <= 0 || channels != 256 || cp != 64 || pixels % (static_cast<int64_t>(shuffle_h) * shuffle_w))
            return fail("srk_rows_to_f16_split: pixel-shuffle input needs channels == 256, cp == 64, pixels a multiple of shuffle_h * shuffle_w");
    } else if (cp < channels) {
        return fail("srk_rows_to_f16_split: cp must cover the channels");
    }
    if (act < SRK_ACT_NONE || act > SRK_ACT_GELU) return fail("srk_rows_to_f16_split: unknown activation %d", act);
    return check(srk::launch_rows_to_f16_split(x, ld_in, channels, static_cast<__half*>(hi_f16), static_cast<__half*>(lo_f16), ld_out, cp, pixels,
                                               act, slope, shuffle_h, shuffle_w, static_cast<cudaStream_t>(stream)), "srk_rows_to_f16_split");
}

int srk_image_to_f16_split(const float* x, int64_t sb, int64_t sc, int64_t sy, int64_t sx, int32_t channels, int32_t batch, int32_t height,
                           int32_t width, const float* mean3, float range, void* out_f16, void* stream) {
    if (!x || !mean3 || !out_f16) return fail("srk_image_to_f16_split: null argument");
    if (channels < 1 || channels > 3 || batch <= 0 || height <= 0 || width <= 0 || !aligned16(out_f16)) return fail("srk_image_to_f16_split: bad arguments");
    return check(srk::launch_image_to_f16_split(x, sb, sc, sy, sx, channels, batch, height, width, mean3, range, static_cast<__half*>(out_f16),
                                                static_cast<cudaStream_t>(stream)), "srk_image_to_f16_split");
}

int srk_stitch_normalize(float* E, const float* Wt, int32_t channels, int64_t pixels, void* stream) {
    if (!E || !Wt) return fail("srk_stitch_normalize: null argument");
    return check(srk::launch_stitch_normalize(E, Wt, channels, pixels, static_cast<cudaStream_t>(stream)), "srk_stitch_normalize");
}

}  // extern "C"
