// HBM-bound helper kernels: row LayerNorm, channels-last PixelShuffle, tile stitcher.
// All are pure streaming kernels: 128-bit coalesced loads/stores, grid-stride, no shared memory.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace srk {

// ---- LayerNorm over 180 channels (patch_embed.norm / final norm, network_swinir.py:526-527, :800)
// half-warp per token, 3 float4 per lane (45 float4 = 180 floats)
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        int64_t num_tokens, int ld_in, int ld_out) {
    const int l16 = threadIdx.x & 15;
    const int64_t hw = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 4;
    const int64_t nhw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 4;
    for (int64_t tok = hw; tok < ((num_tokens + 1) & ~int64_t(1)); tok += nhw) {   // keep both half-warps in the shuffles
        const bool live = tok < num_tokens;
        float4 v[3];
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
            const int f = l16 + 16 * jj;
            v[jj] = (live && f < SRK_DIM / 4) ? __ldg(reinterpret_cast<const float4*>(x + tok * ld_in) + f)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float s = 0.f;
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) s += (v[jj].x + v[jj].y) + (v[jj].z + v[jj].w);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / SRK_DIM);
        float q = 0.f;
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
            if (l16 + 16 * jj < SRK_DIM / 4) {
                v[jj].x -= mean; v[jj].y -= mean; v[jj].z -= mean; v[jj].w -= mean;
                q += (v[jj].x * v[jj].x + v[jj].y * v[jj].y) + (v[jj].z * v[jj].z + v[jj].w * v[jj].w);
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.0f / SRK_DIM) + 1e-5f);
        if (live) {
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
                const int f = l16 + 16 * jj;
                if (f < SRK_DIM / 4) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(w) + f);
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + f);
                    float4 o;
                    o.x = v[jj].x * rstd * g.x + bb.x; o.y = v[jj].y * rstd * g.y + bb.y;
                    o.z = v[jj].z * rstd * g.z + bb.z; o.w = v[jj].w * rstd * g.w + bb.w;
                    reinterpret_cast<float4*>(y + tok * ld_out)[f] = o;
                }
            }
        }
    }
}

cudaError_t launch_layernorm(const float* x, float* y, const float* w, const float* b, int64_t num_tokens, int ld_in,
                             int ld_out, cudaStream_t stream) {
    if (num_tokens <= 0) return cudaSuccess;
    const int64_t blocks = (num_tokens * 16 + 255) / 256;
    const int grid = static_cast<int>(blocks < 148 * 16 ? blocks : 148 * 16);
    layernorm_kernel<<<grid, 256, 0, stream>>>(x, y, w, b, num_tokens, ld_in, ld_out);
    return cudaGetLastError();
}

// ---- PixelShuffle(2) on NHWC: out[b, 2h+i, 2w+j, c] = in[b, h, w, 4c + 2i + j]   (network_swinir.py:585)
// One float4 load = the 4 sub-pixels of channel c; a group of `oc` lanes reads 16*oc contiguous bytes and writes
// four oc*4-byte contiguous runs.  Generic r via the scalar path.
__global__ void __launch_bounds__(256) pixelshuffle2_nhwc_kernel(const float4* __restrict__ x, float* __restrict__ y,
                                                                 int64_t in_pixels, int height, int width, int oc) {
    const int64_t total = in_pixels * oc;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t pix = i / oc;
        const int c = static_cast<int>(i - pix * oc);
        const int w = static_cast<int>(pix % width);
        const int64_t bh = pix / width;                      // b * height + h
        const float4 v = __ldg(x + i);
        float* o = y + ((bh * 2) * (2 * static_cast<int64_t>(width)) + 2 * w) * oc + c;
        o[0] = v.x;
        o[oc] = v.y;
        o[2 * static_cast<int64_t>(width) * oc] = v.z;
        o[2 * static_cast<int64_t>(width) * oc + oc] = v.w;
    }
}

__global__ void __launch_bounds__(256) pixelshuffle_nhwc_generic_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                        int64_t out_elems, int height, int width, int oc,
                                                                        int r) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < out_elems;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % oc);
        int64_t t = i / oc;
        const int ow = static_cast<int>(t % (static_cast<int64_t>(width) * r));
        t /= static_cast<int64_t>(width) * r;
        const int oh = static_cast<int>(t % (static_cast<int64_t>(height) * r));
        const int64_t b = t / (static_cast<int64_t>(height) * r);
        const int h = oh / r, ii = oh % r, w = ow / r, jj = ow % r;
        y[i] = __ldg(x + ((b * height + h) * width + w) * (static_cast<int64_t>(oc) * r * r) + c * r * r + ii * r + jj);
    }
}

cudaError_t launch_pixelshuffle_nhwc(const float* x, float* y, int batch, int height, int width, int out_channels, int r,
                                     cudaStream_t stream) {
    const int64_t in_pixels = static_cast<int64_t>(batch) * height * width;
    if (in_pixels == 0) return cudaSuccess;
    if (r == 2) {
        const int64_t total = in_pixels * out_channels;
        const int64_t blocks = (total + 255) / 256;
        const int grid = static_cast<int>(blocks < 148 * 32 ? blocks : 148 * 32);
        pixelshuffle2_nhwc_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), y, in_pixels, height, width,
                                                            out_channels);
    } else {
        const int64_t total = in_pixels * out_channels * r * r;
        const int64_t blocks = (total + 255) / 256;
        const int grid = static_cast<int>(blocks < 148 * 32 ? blocks : 148 * 32);
        pixelshuffle_nhwc_generic_kernel<<<grid, 256, 0, stream>>>(x, y, total, height, width, out_channels, r);
    }
    return cudaGetLastError();
}

// ---- overlapping-tile stitcher (BASELINE.json configs[4]): E += tile, Wt += 1 over the tile footprint.
// tiles: (num_tiles, channels, tile_h, tile_w) NCHW fp32;  E: (channels, out_h, out_w);  Wt: (out_h, out_w).
// Tiles of one launch must be pairwise disjoint (the tiler launches the 4 parity classes separately).
__global__ void __launch_bounds__(256) stitch_accumulate_kernel(const float* __restrict__ tiles, float* __restrict__ E,
                                                                float* __restrict__ Wt, const int32_t* __restrict__ tile_yx,
                                                                int channels, int tile_h, int tile_w, int out_h, int out_w) {
    const int tile = blockIdx.y;
    const int y0 = tile_yx[2 * tile], x0 = tile_yx[2 * tile + 1];
    const int per = tile_h * tile_w;
    const float* src = tiles + static_cast<int64_t>(tile) * channels * per;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
        const int ty = i / tile_w, tx = i - ty * tile_w;
        const int oy = y0 + ty, ox = x0 + tx;
        if (oy < 0 || ox < 0 || oy >= out_h || ox >= out_w) continue;   // rows of a seam tile outside this rank's band
        const int64_t o = static_cast<int64_t>(oy) * out_w + ox;
        for (int c = 0; c < channels; ++c) E[static_cast<int64_t>(c) * out_h * out_w + o] += src[c * per + i];
        Wt[o] += 1.0f;
    }
}

__global__ void __launch_bounds__(256) stitch_normalize_kernel(float* __restrict__ E, const float* __restrict__ Wt,
                                                               int channels, int64_t pixels) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < pixels;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float inv = 1.0f / Wt[i];
        for (int c = 0; c < channels; ++c) E[c * pixels + i] *= inv;
    }
}

cudaError_t launch_stitch_accumulate(const float* tiles, float* E, float* Wt, const int32_t* tile_yx, int num_tiles,
                                     int channels, int tile_h, int tile_w, int out_h, int out_w, cudaStream_t stream) {
    if (num_tiles <= 0) return cudaSuccess;
    const int per = tile_h * tile_w;
    dim3 grid((per + 255) / 256, num_tiles);
    stitch_accumulate_kernel<<<grid, 256, 0, stream>>>(tiles, E, Wt, tile_yx, channels, tile_h, tile_w, out_h, out_w);
    return cudaGetLastError();
}

cudaError_t launch_stitch_normalize(float* E, const float* Wt, int channels, int64_t pixels, cudaStream_t stream) {
    if (pixels <= 0) return cudaSuccess;
    const int64_t blocks = (pixels + 255) / 256;
    const int grid = static_cast<int>(blocks < 148 * 32 ? blocks : 148 * 32);
    stitch_normalize_kernel<<<grid, 256, 0, stream>>>(E, Wt, channels, pixels);
    return cudaGetLastError();
}

}  // namespace srk
