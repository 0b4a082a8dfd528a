// HBM-bound helper kernels: row LayerNorm, channels-last PixelShuffle, tile stitcher.
// All are pure streaming kernels: 128-bit coalesced loads/stores, grid-stride, no shared memory.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace srk {

// ---- LayerNorm over 180 channels (patch_embed.norm / final norm, network_swinir.py:526-527, :800)
// half-warp per token, 3 float4 per lane (45 float4 = 180 floats)
// x and y may be the same buffer (the final norm runs in place): no __restrict__ / read-only path on them.  y16 (optional): the
// result as fp16 NHWC rows of SRK_DIM_PAD channels, zero padded -- the input layout of the 3x3 convolution (conv_kernel.cu).
__global__ void __launch_bounds__(256) layernorm_kernel(const float* x, float* y, __half* __restrict__ y16,
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
            v[jj] = (live && f < SRK_DIM / 4) ? reinterpret_cast<const float4*>(x + tok * ld_in)[f] : make_float4(0.f, 0.f, 0.f, 0.f);
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
                    if (y) reinterpret_cast<float4*>(y + tok * ld_out)[f] = o;
                    if (y16) {
                        const __half2 lo = __floats2half2_rn(o.x, o.y), hi = __floats2half2_rn(o.z, o.w);
                        reinterpret_cast<uint2*>(y16 + tok * SRK_DIM_PAD)[f] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
                    }
                } else if (y16) {
                    reinterpret_cast<uint2*>(y16 + tok * SRK_DIM_PAD)[f] = make_uint2(0u, 0u);      // channels 180..191
                }
            }
        }
    }
}

cudaError_t launch_layernorm(const float* x, float* y, __half* y16, const float* w, const float* b, int64_t num_tokens, int ld_in,
                             int ld_out, cudaStream_t stream) {
    if (num_tokens <= 0) return cudaSuccess;
    const int64_t blocks = (num_tokens * 16 + 255) / 256;
    const int grid = static_cast<int>(blocks < 148 * 16 ? blocks : 148 * 16);
    layernorm_kernel<<<grid, 256, 0, stream>>>(x, y, y16, w, b, num_tokens, ld_in, ld_out);
    return cudaGetLastError();
}

// ---- PixelShuffle(2) on NHWC: out[b, 2h+i, 2w+j, c] = in[b, h, w, 4c + 2i + j]   (network_swinir.py:585)
// One float4 load = the 4 sub-pixels of channel c; a group of `oc` lanes reads 16*oc contiguous bytes and writes
// four oc*4-byte contiguous runs.  Generic r via the scalar path.
__global__ void __launch_bounds__(256) pixelshuffle2_nhwc_kernel(const float4* __restrict__ x, const float4* __restrict__ bias,
                                                                 float* __restrict__ y, int64_t in_pixels, int height, int width, int oc) {
    const int64_t total = in_pixels * oc;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t pix = i / oc;
        const int c = static_cast<int>(i - pix * oc);
        const int w = static_cast<int>(pix % width);
        const int64_t bh = pix / width;                      // b * height + h
        float4 v = __ldg(x + i);
        if (bias) { const float4 b = __ldg(bias + c); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }   // conv bias of channels 4c..4c+3
        float* o = y + ((bh * 2) * (2 * static_cast<int64_t>(width)) + 2 * w) * oc + c;
        o[0] = v.x;
        o[oc] = v.y;
        o[2 * static_cast<int64_t>(width) * oc] = v.z;
        o[2 * static_cast<int64_t>(width) * oc + oc] = v.w;
    }
}

__global__ void __launch_bounds__(256) pixelshuffle_nhwc_generic_kernel(const float* __restrict__ x, const float* __restrict__ bias,
                                                                        float* __restrict__ y, int64_t out_elems, int height, int width,
                                                                        int oc, int r) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < out_elems;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % oc);
        int64_t t = i / oc;
        const int ow = static_cast<int>(t % (static_cast<int64_t>(width) * r));
        t /= static_cast<int64_t>(width) * r;
        const int oh = static_cast<int>(t % (static_cast<int64_t>(height) * r));
        const int64_t b = t / (static_cast<int64_t>(height) * r);
        const int h = oh / r, ii = oh % r, w = ow / r, jj = ow % r;
        const int ic = c * r * r + ii * r + jj;
        y[i] = __ldg(x + ((b * height + h) * width + w) * (static_cast<int64_t>(oc) * r * r) + ic) + (bias ? __ldg(bias + ic) : 0.f);
    }
}

cudaError_t launch_pixelshuffle_nhwc(const float* x, const float* bias, float* y, int batch, int height, int width, int out_channels,
                                     int r, cudaStream_t stream) {
    const int64_t in_pixels = static_cast<int64_t>(batch) * height * width;
    if (in_pixels == 0) return cudaSuccess;
    if (r == 2) {
        const int64_t total = in_pixels * out_channels;
        const int64_t blocks = (total + 255) / 256;
        const int grid = static_cast<int>(blocks < 148 * 32 ? blocks : 148 * 32);
        pixelshuffle2_nhwc_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(bias), y,
                                                            in_pixels, height, width, out_channels);
    } else {
        const int64_t total = in_pixels * out_channels * r * r;
        const int64_t blocks = (total + 255) / 256;
        const int grid = static_cast<int>(blocks < 148 * 32 ? blocks : 148 * 32);
        pixelshuffle_nhwc_generic_kernel<<<grid, 256, 0, stream>>>(x, bias, y, total, height, width, out_channels, r);
    }
    return cudaGetLastError();
}

// ---- y = act(x + bias[c]) + residual on channels-last activations ([pixels, C] rows): the per-channel bias of a
// library convolution, an optional LeakyReLU and the block's residual add in one pass (torch runs these as a
// non-vectorised broadcast add plus one more elementwise kernel per operation).  y may alias x or residual.
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }   // nn.GELU() (exact)
template <bool VEC>
__global__ void __launch_bounds__(256) bias_act_add_kernel(const float* x, const float* __restrict__ bias, const float* residual,
                                                           float* y, int64_t total, int channels, int act, float slope) {
    const int64_t n = VEC ? total / 4 : total;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (VEC) {
            const int c = static_cast<int>((i * 4) % channels);
            float4 v = reinterpret_cast<const float4*>(x)[i];
            if (bias) { const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c)); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
            if (act == 1) {
                v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
                v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
            } else if (act == 2) {
                v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
            }
            if (residual) { const float4 r = reinterpret_cast<const float4*>(residual)[i]; v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
            reinterpret_cast<float4*>(y)[i] = v;
        } else {
            float v = x[i];
            if (bias) v += __ldg(bias + static_cast<int>(i % channels));
            if (act == 1) v = v > 0.f ? v : v * slope;
            else if (act == 2) v = gelu_erf(v);
            if (residual) v += residual[i];
            y[i] = v;
        }
    }
}

cudaError_t launch_bias_act_add(const float* x, const float* bias, const float* residual, float* y, int64_t pixels, int channels,
                                int act, float slope, cudaStream_t stream) {
    const int64_t total = pixels * channels;
    if (total <= 0) return cudaSuccess;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool vec = (channels % 4 == 0) && al16(x) && al16(y) && (!bias || al16(bias)) && (!residual || al16(residual));
    const int64_t n = vec ? total / 4 : total;
    const int64_t blocks = (n + 255) / 256;
    const int grid = static_cast<int>(blocks < 148 * 32 ? blocks : 148 * 32);
    if (vec) bias_act_add_kernel<true><<<grid, 256, 0, stream>>>(x, bias, residual, y, total, channels, act, slope);
    else     bias_act_add_kernel<false><<<grid, 256, 0, stream>>>(x, bias, residual, y, total, channels, act, slope);
    return cudaGetLastError();
}

// ---- overlapping-tile stitcher (BASELINE.json configs[4]): E += tile, Wt += 1 over the tile footprint.
// tiles: (num_tiles, channels, tile_h, tile_w) NCHW fp32;  E: (channels, out_h, out_w);  Wt: (out_h, out_w).
// Tiles of one launch must be pairwise disjoint (the tiler launches the 9 parity-or-last classes separately).
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

// ---- tiler v2 (tiling.py): tile gather, strided accumulate without the count plane, finalize with count tables.
// gather: out[n][c][ty][tx] = slab[c][y0 + ty][x0 + tx] for the LR band `slab` (channels, rows, width) of this rank.
__global__ void __launch_bounds__(256) gather_tiles_kernel(const float* __restrict__ slab, int64_t slab_cstride, int slab_w,
                                                           const int32_t* __restrict__ src_yx, int channels, int tile_h, int tile_w,
                                                           float* __restrict__ out) {
    const int tile = blockIdx.z, c = blockIdx.y;
    const int y0 = src_yx[2 * tile], x0 = src_yx[2 * tile + 1];
    const int per = tile_h * tile_w;
    float* dst = out + (static_cast<int64_t>(tile) * channels + c) * per;
    const float* src = slab + c * slab_cstride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
        const int ty = i / tile_w, tx = i - ty * tile_w;
        dst[i] = __ldg(src + static_cast<int64_t>(y0 + ty) * slab_w + x0 + tx);
    }
}
// accumulate: E[c][y0 + ty][x0 + tx] += tiles[n][c][ty][tx] for any element strides of `tiles` (the models return channels-last
// memory behind an NCHW shape).  Tiles of one launch must be pairwise disjoint (tiling.py launches per class segment); the
// per-pixel cover count is NOT accumulated -- it is the product of two 1-D tables (stitch_finalize_kernel).
__global__ void __launch_bounds__(256) stitch_accumulate2_kernel(const float* __restrict__ tiles, int64_t sn, int64_t sc, int64_t sy, int64_t sx,
                                                                 float* __restrict__ E, int64_t e_cstride, const int32_t* __restrict__ dst_yx,
                                                                 int channels, int tile_h, int tile_w, int out_h, int out_w) {
    const int tile = blockIdx.y;
    const int y0 = dst_yx[2 * tile], x0 = dst_yx[2 * tile + 1];
    const int per = tile_h * tile_w;
    const float* src = tiles + tile * sn;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
        const int ty = i / tile_w, tx = i - ty * tile_w;
        const int oy = y0 + ty, ox = x0 + tx;
        if (oy < 0 || ox < 0 || oy >= out_h || ox >= out_w) continue;   // rows of a seam tile outside this rank's band
        const int64_t o = static_cast<int64_t>(oy) * out_w + ox;
        const float* s = src + ty * sy + tx * sx;
        for (int c = 0; c < channels; ++c) E[c * e_cstride + o] += __ldg(s + c * sc);
    }
}
// finalize: out[c][y][x] = convert(E[c][y][x] / (cnt_y[y] * cnt_x[x])); dtype 0 = fp32 (out may alias E), 1 = bf16, 2 = uint8
// (round(clamp(v, 0, 1) * 255)).  cnt_* hold small integers as floats, so the divisor is exact: this IS E / W of the upstream rule.
template <int DTYPE>
__global__ void __launch_bounds__(256) stitch_finalize_kernel(const float* __restrict__ E, int64_t e_cstride, const float* __restrict__ cnt_y,
                                                              const float* __restrict__ cnt_x, void* __restrict__ out, int64_t o_cstride,
                                                              int channels, int out_h, int out_w) {
    const int64_t pixels = static_cast<int64_t>(out_h) * out_w;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < pixels; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int y = static_cast<int>(i / out_w), x = static_cast<int>(i - static_cast<int64_t>(y) * out_w);
        const float w = __ldg(cnt_y + y) * __ldg(cnt_x + x);
        for (int c = 0; c < channels; ++c) {
            const float v = E[c * e_cstride + i] / w;
            if (DTYPE == 0) static_cast<float*>(out)[c * o_cstride + i] = v;
            else if (DTYPE == 1) static_cast<__nv_bfloat16*>(out)[c * o_cstride + i] = __float2bfloat16_rn(v);
            else static_cast<uint8_t*>(out)[c * o_cstride + i] = static_cast<uint8_t>(__float2int_rn(fminf(fmaxf(v, 0.f), 1.f) * 255.f));
        }
    }
}

cudaError_t launch_gather_tiles(const float* slab, int64_t slab_cstride, int slab_w, const int32_t* src_yx, int num_tiles, int channels,
                                int tile_h, int tile_w, float* out, cudaStream_t stream) {
    if (num_tiles <= 0) return cudaSuccess;
    dim3 grid((tile_h * tile_w + 255) / 256, channels, num_tiles);
    gather_tiles_kernel<<<grid, 256, 0, stream>>>(slab, slab_cstride, slab_w, src_yx, channels, tile_h, tile_w, out);
    return cudaGetLastError();
}
cudaError_t launch_stitch_accumulate2(const float* tiles, int64_t sn, int64_t sc, int64_t sy, int64_t sx, float* E, int64_t e_cstride,
                                      const int32_t* dst_yx, int num_tiles, int channels, int tile_h, int tile_w, int out_h, int out_w,
                                      cudaStream_t stream) {
    if (num_tiles <= 0) return cudaSuccess;
    dim3 grid((tile_h * tile_w + 255) / 256, num_tiles);
    stitch_accumulate2_kernel<<<grid, 256, 0, stream>>>(tiles, sn, sc, sy, sx, E, e_cstride, dst_yx, channels, tile_h, tile_w, out_h, out_w);
    return cudaGetLastError();
}
cudaError_t launch_stitch_finalize(const float* E, int64_t e_cstride, const float* cnt_y, const float* cnt_x, void* out, int64_t o_cstride,
                                   int dtype, int channels, int out_h, int out_w, cudaStream_t stream) {
    const int64_t pixels = static_cast<int64_t>(out_h) * out_w;
    if (pixels <= 0) return cudaSuccess;
    const int64_t blocks = (pixels + 255) / 256;
    const int grid = static_cast<int>(blocks < 148 * 32 ? blocks : 148 * 32);
    if (dtype == 0) stitch_finalize_kernel<0><<<grid, 256, 0, stream>>>(E, e_cstride, cnt_y, cnt_x, out, o_cstride, channels, out_h, out_w);
    else if (dtype == 1) stitch_finalize_kernel<1><<<grid, 256, 0, stream>>>(E, e_cstride, cnt_y, cnt_x, out, o_cstride, channels, out_h, out_w);
    else stitch_finalize_kernel<2><<<grid, 256, 0, stream>>>(E, e_cstride, cnt_y, cnt_x, out, o_cstride, channels, out_h, out_w);
    return cudaGetLastError();
}

// ---- HAT CAB tail (hat_arch.py:41-59 ChannelAttention + :307 `+ conv_x * conv_scale`):
//      out[b, t, c] += scale * y[b, t, c] * sigmoid(W2 relu(W1 mean_t(y[b, :, c]) + b1) + b2)[c]
// y is the CAB's second convolution output in channels-last memory (B, H*W, 180).  Two streaming passes over y (L2 resident):
// (1) per-(image, channel) sums; (2) the 180 -> hidden -> 180 squeeze-excite gate (recomputed by every block: 2 x 180 x hidden
// MACs) fused with the scaled residual add.  Replaces five torch launches (mean, two 1x1 convs, ReLU / sigmoid, addcmul).
constexpr int CAB_TOK_PER_BLOCK = 64;
// Sum of the per-chunk partial sums of one image -> s_out[180], by 45 float4 channel groups x KG chunk groups of threads (a serial
// loop over the 64 chunks per channel cost every block ~5 us of dependent L2 round trips).  Fixed order: run-to-run identical.
template <int KG>
__device__ __forceinline__ void sum_partials(const float* __restrict__ sums, int chunks, float* s_out, float4 (*s_part)[SRK_DIM / 4]) {
    const int cg = threadIdx.x % (SRK_DIM / 4), kg = threadIdx.x / (SRK_DIM / 4);
    if (kg < KG) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = kg; k < chunks; k += KG) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(sums + static_cast<int64_t>(k) * SRK_DIM) + cg);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        s_part[kg][cg] = a;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < SRK_DIM; c += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < KG; ++g) s += reinterpret_cast<const float*>(s_part[g])[c];
        s_out[c] = s;
    }
    __syncthreads();
}
__global__ void __launch_bounds__(192) cab_pool_kernel(const float* __restrict__ y, float* __restrict__ sums, int tokens_per_image) {
    const int b = blockIdx.y, c = threadIdx.x;
    const int t0 = blockIdx.x * CAB_TOK_PER_BLOCK;
    const int t1 = min(t0 + CAB_TOK_PER_BLOCK, tokens_per_image);
    if (c >= SRK_DIM) return;
    const float* src = y + (static_cast<int64_t>(b) * tokens_per_image + t0) * SRK_DIM + c;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int t = t0;
    for (; t + 4 <= t1; t += 4, src += 4 * SRK_DIM) {
        s0 += __ldg(src); s1 += __ldg(src + SRK_DIM); s2 += __ldg(src + 2 * SRK_DIM); s3 += __ldg(src + 3 * SRK_DIM);
    }
    for (; t < t1; ++t, src += SRK_DIM) s0 += __ldg(src);
    // one partial per (image, token chunk): summed in a fixed order by cab_gate_add_kernel, so the result is run-to-run identical
    sums[(static_cast<int64_t>(b) * gridDim.x + blockIdx.x) * SRK_DIM + c] = (s0 + s1) + (s2 + s3);
}

__global__ void __launch_bounds__(256) cab_gate_add_kernel(const float* __restrict__ y, const float* __restrict__ y_bias,
                                                           float* __restrict__ out,
                                                           const float* __restrict__ sums, const float* __restrict__ w1,
                                                           const float* __restrict__ b1, const float* __restrict__ w2,
                                                           const float* __restrict__ b2, int hidden, float scale,
                                                           int tokens_per_image) {
    __shared__ float s_mean[SRK_DIM], s_hid[32];
    __shared__ __align__(16) float s_gate[SRK_DIM], s_yb[SRK_DIM];      // s_yb: the bias of the conv that produced y (or 0)
    __shared__ float4 s_part[5][SRK_DIM / 4];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // hidden <= 8 (HAT: 180 / 30 = 6): this thread's squeeze / excite weights are fetched before the partial sums, so the three
    // dependent global round trips of the prologue (partial sums -> W1 -> W2) become one
    const bool small = hidden <= 8;
    float w1r[6], w2r[8], b1r = 0.f, b2r = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) w1r[k] = (small && warp < hidden && lane + 32 * k < SRK_DIM) ? __ldg(w1 + warp * SRK_DIM + lane + 32 * k) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) w2r[j] = (small && threadIdx.x < SRK_DIM && j < hidden) ? __ldg(w2 + threadIdx.x * hidden + j) : 0.f;
    if (small && warp < hidden) b1r = __ldg(b1 + warp);
    if (small && threadIdx.x < SRK_DIM) b2r = __ldg(b2 + threadIdx.x);
    sum_partials<5>(sums + static_cast<int64_t>(b) * gridDim.x * SRK_DIM, gridDim.x, s_mean, s_part);
    for (int c = threadIdx.x; c < SRK_DIM; c += blockDim.x) {
        const float yb = y_bias ? __ldg(y_bias + c) : 0.f;
        s_yb[c] = yb;
        s_mean[c] = s_mean[c] / static_cast<float>(tokens_per_image) + yb;      // mean(y + bias) = mean(y) + bias
    }
    __syncthreads();
    if (small) {
        if (warp < hidden) {                                 // hidden unit `warp`: lanes stride over the 180 inputs (same order as below)
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 6; ++k) if (lane + 32 * k < SRK_DIM) a = fmaf(w1r[k], s_mean[lane + 32 * k], a);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) s_hid[warp] = fmaxf(a + b1r, 0.f);
        }
        __syncthreads();
        if (threadIdx.x < SRK_DIM) {
            float a = b2r;
#pragma unroll
            for (int j = 0; j < 8; ++j) if (j < hidden) a = fmaf(w2r[j], s_hid[j], a);
            s_gate[threadIdx.x] = scale / (1.0f + __expf(-a));
        }
    } else {
        for (int j = warp; j < hidden; j += 8) {             // hidden unit j: one warp, lanes stride over the 180 inputs
            float a = 0.f;
            for (int c = lane; c < SRK_DIM; c += 32) a = fmaf(w1[j * SRK_DIM + c], s_mean[c], a);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) s_hid[j] = fmaxf(a + b1[j], 0.f);
        }
        __syncthreads();
        for (int c = threadIdx.x; c < SRK_DIM; c += blockDim.x) {
            float a = b2[c];
            for (int j = 0; j < hidden; ++j) a = fmaf(w2[c * hidden + j], s_hid[j], a);
            s_gate[c] = scale / (1.0f + __expf(-a));
        }
    }
    __syncthreads();
    // 64 tokens x 45 float4 per block (a plain block-stride loop: a (channel group, row phase) thread mapping with gate / bias in
    // registers and four rows in flight was SLOWER, 7.47 vs 7.24 ms per HAT step)
    const int t0 = blockIdx.x * CAB_TOK_PER_BLOCK;
    const int nt = min(CAB_TOK_PER_BLOCK, tokens_per_image - t0);
    const int64_t base = (static_cast<int64_t>(b) * tokens_per_image + t0) * (SRK_DIM / 4);
    const float4* ys = reinterpret_cast<const float4*>(y) + base;
    float4* os = reinterpret_cast<float4*>(out) + base;
    const float4* gs = reinterpret_cast<const float4*>(s_gate);
    const float4* bs = reinterpret_cast<const float4*>(s_yb);
    for (int i = threadIdx.x; i < nt * (SRK_DIM / 4); i += blockDim.x) {
        const float4 g = gs[i % (SRK_DIM / 4)], yb = bs[i % (SRK_DIM / 4)], v = __ldg(ys + i);
        float4 o = os[i];
        o.x = fmaf(v.x + yb.x, g.x, o.x); o.y = fmaf(v.y + yb.y, g.y, o.y); o.z = fmaf(v.z + yb.z, g.z, o.z); o.w = fmaf(v.w + yb.w, g.w, o.w);
        os[i] = o;
    }
}

// ---- mean over the tokens of channels-last (batch, tokens, 180) rows: the per-chunk partial sums of cab_pool_kernel, summed in a
//      fixed order (run-to-run identical).  DAT's channel-interaction pooling (dat_arch.py:305-310, AdaptiveAvgPool2d(1)).
__global__ void __launch_bounds__(192) token_mean_reduce_kernel(const float* __restrict__ sums, float* __restrict__ mean, int chunks,
                                                                int tokens_per_image) {
    const int b = blockIdx.x, c = threadIdx.x;
    if (c >= SRK_DIM) return;
    float s = 0.f;
    for (int k = 0; k < chunks; ++k) s += __ldg(sums + (static_cast<int64_t>(b) * chunks + k) * SRK_DIM + c);
    mean[b * SRK_DIM + c] = s / static_cast<float>(tokens_per_image);
}

// ---- ... followed by the squeeze MLP of DAT's channel interaction (dat_arch.py:305-310): out[b] = W2 gelu(W1 mean[b] + b1) + b2 with the
//      eval BatchNorm folded into W1 / b1 and the exact (erf) GELU; one block per image, a few thousand MACs.
__global__ void __launch_bounds__(192) token_mean_mlp_kernel(const float* __restrict__ sums, const float* __restrict__ w1,
                                                             const float* __restrict__ b1, const float* __restrict__ w2,
                                                             const float* __restrict__ b2, int hidden, float* __restrict__ out, int chunks,
                                                             int tokens_per_image) {
    __shared__ float s_mean[SRK_DIM];
    __shared__ float s_h[64];
    __shared__ float4 s_part[4][SRK_DIM / 4];
    const int b = blockIdx.x, c = threadIdx.x;
    sum_partials<4>(sums + static_cast<int64_t>(b) * chunks * SRK_DIM, chunks, s_mean, s_part);
    if (c < SRK_DIM) s_mean[c] = s_mean[c] / static_cast<float>(tokens_per_image);
    __syncthreads();
    for (int j = c >> 5; j < hidden; j += 6) {             // warp per hidden unit: lane-strided dot product, butterfly sum
        float a = 0.f;
        for (int k = c & 31; k < SRK_DIM; k += 32) a = fmaf(__ldg(w1 + j * SRK_DIM + k), s_mean[k], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        a += __ldg(b1 + j);
        if ((c & 31) == 0) s_h[j] = 0.5f * a * (1.0f + erff(a * 0.70710678118654752440f));
    }
    __syncthreads();
    if (c < SRK_DIM) {
        float a = __ldg(b2 + c);
#pragma unroll 8
        for (int j = 0; j < hidden; ++j) a = fmaf(__ldg(w2 + c * hidden + j), s_h[j], a);
        out[b * SRK_DIM + c] = a;
    }
}

cudaError_t launch_token_mean_mlp(const float* x, float* out, float* sums, const float* w1, const float* b1, const float* w2, const float* b2,
                                  int hidden, int batch, int tokens_per_image, cudaStream_t stream) {
    if (batch <= 0 || tokens_per_image <= 0) return cudaSuccess;
    const int chunks = (tokens_per_image + CAB_TOK_PER_BLOCK - 1) / CAB_TOK_PER_BLOCK;
    cab_pool_kernel<<<dim3(chunks, batch), 192, 0, stream>>>(x, sums, tokens_per_image);
    token_mean_mlp_kernel<<<batch, 192, 0, stream>>>(sums, w1, b1, w2, b2, hidden, out, chunks, tokens_per_image);
    return cudaGetLastError();
}

cudaError_t launch_token_mean(const float* x, float* mean, float* sums, int batch, int tokens_per_image, cudaStream_t stream) {
    if (batch <= 0 || tokens_per_image <= 0) return cudaSuccess;
    const int chunks = (tokens_per_image + CAB_TOK_PER_BLOCK - 1) / CAB_TOK_PER_BLOCK;
    cab_pool_kernel<<<dim3(chunks, batch), 192, 0, stream>>>(x, sums, tokens_per_image);
    token_mean_reduce_kernel<<<batch, 192, 0, stream>>>(sums, mean, chunks, tokens_per_image);
    return cudaGetLastError();
}

int cab_ws_floats(int batch, int tokens_per_image) {
    return batch * ((tokens_per_image + CAB_TOK_PER_BLOCK - 1) / CAB_TOK_PER_BLOCK) * SRK_DIM;
}

cudaError_t launch_cab_gate_add(const float* y, const float* y_bias, float* out, float* sums, const float* w1, const float* b1, const float* w2,
                                const float* b2, int hidden, float scale, int batch, int tokens_per_image, cudaStream_t stream) {
    if (batch <= 0 || tokens_per_image <= 0) return cudaSuccess;
    dim3 grid((tokens_per_image + CAB_TOK_PER_BLOCK - 1) / CAB_TOK_PER_BLOCK, batch);
    cab_pool_kernel<<<grid, 192, 0, stream>>>(y, sums, tokens_per_image);
    cab_gate_add_kernel<<<grid, 256, 0, stream>>>(y, y_bias, out, sums, w1, b1, w2, b2, hidden, scale, tokens_per_image);
    return cudaGetLastError();
}

}  // namespace srk
