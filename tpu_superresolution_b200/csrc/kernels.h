// Internal launch interface between capi.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "srk.h"

namespace srk {

struct AttnParams {
    const float* x;
    float* y;
    const uint8_t* wstream;
    const float* vec;
    const float* mask;
    int mode, H, W, shift, nwx, nw_img;
    int total_windows, n_tiles;
    int ld_in, ld_out;
    int apply_ln, add_residual, mask_mode, mask_nw;
    unsigned long long* dbg;   // optional timeline buffer (srk_debug_set_timeline), else nullptr
    int stagger;               // start skew in cycles per (CTA index mod 4)
};

struct MlpParams {
    const float* x;
    float* y;
    const uint8_t* wstream;
    const float* vec;
    int64_t num_tokens;
    int n_tiles;
    int ld_in, ld_out;
    int apply_ln, add_residual;
    unsigned long long* dbg;
    int stagger;
};

extern unsigned long long* g_timeline;
extern int g_stagger_attn, g_stagger_mlp;
cudaError_t launch_swin_attn(const AttnParams& p, cudaStream_t stream);
cudaError_t launch_swin_mlp(const MlpParams& p, cudaStream_t stream);
cudaError_t launch_layernorm(const float* x, float* y, const float* w, const float* b, int64_t num_tokens, int ld_in,
                             int ld_out, cudaStream_t stream);
cudaError_t launch_pixelshuffle_nhwc(const float* x, float* y, int batch, int height, int width, int out_channels, int r,
                                     cudaStream_t stream);
cudaError_t launch_stitch_accumulate(const float* tiles, float* E, float* Wt, const int32_t* tile_yx, int num_tiles,
                                     int channels, int tile_h, int tile_w, int out_h, int out_w, cudaStream_t stream);
cudaError_t launch_stitch_normalize(float* E, const float* Wt, int channels, int64_t pixels, cudaStream_t stream);

}  // namespace srk
