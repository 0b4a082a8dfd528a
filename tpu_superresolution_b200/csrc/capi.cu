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
void srk_debug_set_timeline(void* buf) { srk::g_timeline = static_cast<unsigned long long*>(buf); }
int64_t srk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int srk_swin_attn_fwd(const SrkSwinAttnDesc* d, const float* x, float* y, const void* wstream, const float* vec,
                      const float* mask, void* stream) {
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
    p.dbg = srk::g_timeline;
    p.stagger = p.n_tiles >= 2 * 148 ? srk::g_stagger_attn : 0;
    return check(srk::launch_swin_attn(p, static_cast<cudaStream_t>(stream)), "srk_swin_attn_fwd");
}

int srk_swin_mlp_fwd(const SrkMlpDesc* d, const float* x, float* y, const void* wstream, const float* vec, void* stream) {
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
    p.dbg = srk::g_timeline;
    p.stagger = p.n_tiles >= 2 * 148 ? srk::g_stagger_mlp : 0;
    return check(srk::launch_swin_mlp(p, static_cast<cudaStream_t>(stream)), "srk_swin_mlp_fwd");
}

int srk_layernorm_fwd(const float* x, float* y, const float* w, const float* b, int64_t num_tokens, int32_t ld_in,
                      int32_t ld_out, void* stream) {
    if (!x || !y || !w || !b) return fail("srk_layernorm_fwd: null argument");
    if (ld_in < SRK_DIM || ld_out < SRK_DIM || (ld_in & 3) || (ld_out & 3)) return fail("srk_layernorm_fwd: bad ld");
    if (!aligned16(x) || !aligned16(y) || !aligned16(w) || !aligned16(b)) return fail("srk_layernorm_fwd: pointers must be 16-byte aligned");
    if (num_tokens < 0) return fail("srk_layernorm_fwd: bad num_tokens");
    return check(srk::launch_layernorm(x, y, w, b, num_tokens, ld_in, ld_out, static_cast<cudaStream_t>(stream)), "srk_layernorm_fwd");
}

int srk_pixelshuffle_nhwc_fwd(const float* x, float* y, int32_t batch, int32_t height, int32_t width, int32_t out_channels,
                              int32_t r, void* stream) {
    if (!x || !y) return fail("srk_pixelshuffle_nhwc_fwd: null argument");
    if (batch < 0 || height < 0 || width < 0 || out_channels <= 0 || r <= 0) return fail("srk_pixelshuffle_nhwc_fwd: bad shape");
    if (r == 2 && !aligned16(x)) return fail("srk_pixelshuffle_nhwc_fwd: x must be 16-byte aligned");
    return check(srk::launch_pixelshuffle_nhwc(x, y, batch, height, width, out_channels, r, static_cast<cudaStream_t>(stream)),
                 "srk_pixelshuffle_nhwc_fwd");
}

int srk_stitch_accumulate(const float* tiles, float* E, float* Wt, const int32_t* tile_yx, int32_t num_tiles, int32_t channels,
                          int32_t tile_h, int32_t tile_w, int32_t out_h, int32_t out_w, void* stream) {
    if (!tiles || !E || !Wt || !tile_yx) return fail("srk_stitch_accumulate: null argument");
    if (num_tiles < 0 || num_tiles > 65535 || channels <= 0 || tile_h <= 0 || tile_w <= 0) return fail("srk_stitch_accumulate: bad shape");
    return check(srk::launch_stitch_accumulate(tiles, E, Wt, tile_yx, num_tiles, channels, tile_h, tile_w, out_h, out_w,
                                               static_cast<cudaStream_t>(stream)), "srk_stitch_accumulate");
}

int srk_stitch_normalize(float* E, const float* Wt, int32_t channels, int64_t pixels, void* stream) {
    if (!E || !Wt) return fail("srk_stitch_normalize: null argument");
    return check(srk::launch_stitch_normalize(E, Wt, channels, pixels, static_cast<cudaStream_t>(stream)), "srk_stitch_normalize");
}

}  // extern "C"
