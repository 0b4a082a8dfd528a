// HBM / L2-bound kernels of the DAT blocks (dat_arch.py), all on fp32 channels-last token rows [token][ld]:
//
//   dwconv3x3_rows_kernel   depthwise 3x3 (zero padding) + per-channel affine (eval BatchNorm folded) + GELU, optionally on
//                           LayerNorm-ed input rows and multiplied by a gate operand:
//                             * dwconv + BN + GELU of v     (dat_arch.py:300-304, :418, :508)
//                             * SpatialGate: x1 * dwconv(LayerNorm(x2))   (dat_arch.py:38-54)
//   row_stats_kernel        per-token mean / rstd over a channel slice (the SpatialGate LayerNorm statistics)
//   dat_mix_kernel          adaptive interaction module (dat_arch.py:420-433, :510-523): per-token squeeze MLP 180 -> 11 -> 1
//                           (BatchNorm folded, exact GELU), sigmoid gates, attention / convolution branch mix
//   channel_gram_kernel     per (image, head): q^T k over the tokens, squared norms of q, k (dat_arch.py:497-500)
//   channel_apply_kernel    out[tok, h*30+d1] = sum_d2 A[b, h, d1, d2] v[tok, h*30+d2]      (dat_arch.py:505)
// 128-bit accesses where the layout allows, grid-stride or one block per token chunk, no tensor cores: these are
// streaming kernels whose bound is bytes moved.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kernels.h"

namespace srk {

namespace {
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
}  // namespace

// ---- depthwise 3x3 on token rows.  in: rows of `ld_in` floats, channel slice [c_in, c_in + C); weights w[9][C] (tap major),
//      out = act((sum_taps w * xin) * scale + shift) [* gate], xin = LayerNorm(in) when stats != nullptr (gamma, beta given).
//      Block = (image, band of DW_ROWS output rows, slab of DW_SLAB float4 channel groups), 16 x DW_SLAB threads: thread =
//      (channel group c, column phase xl), so everything per-channel (tap weights, LayerNorm affine, BN scale / shift) sits in
//      registers and the loops have no divisions.  The block walks down its band with a ring of four input rows in shared
//      memory (zero halo column on each side, LayerNorm applied on the way in): while output row y is computed from rows
//      y-1 .. y+1, the global loads of row y+2 are already in flight in registers.  Each input element is read
//      (DW_ROWS + 2) / DW_ROWS times instead of nine.
constexpr int DW_SLAB = 15;            // float4 channel groups per block: 60 channels, 240 contiguous bytes per token
constexpr int DW_XL = 16;              // column phases per block
constexpr int DW_ROWS = 8;             // output rows per block
constexpr int DW_XC = 64;              // output columns per block (blockIdx.z = column chunk)
constexpr int DW_MAXCOL = 5;           // tile columns per thread: ceil((DW_XC + 2) / DW_XL)
__device__ __forceinline__ float gelu_erf_fast(float x) {
    // erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, below fp32 rounding of the GELU output): a dozen instructions
    // instead of the ~70 of erff(), which dominated this kernel's instruction count
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
    const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
    const float e = 1.0f - poly * __expf(-z * z);
    return 0.5f * x * (1.0f + copysignf(e, x));
}
template <bool HAS_LN, bool HAS_GATE>
__global__ void __launch_bounds__(DW_SLAB * DW_XL, 2) dwconv3x3_rows_kernel(const float* __restrict__ in, int ld_in, int c_in,
                                                                         const float* __restrict__ w, const float* __restrict__ scale,
                                                                         const float* __restrict__ shift, const float* __restrict__ stats,
                                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                         const float* __restrict__ gate, int ld_gate, int c_gate,
                                                                         float* __restrict__ out, int ld_out, int C, int H, int W, int act_gelu,
                                                                         uint8_t* __restrict__ out_planes, long long plane_stride) {
    extern __shared__ float4 dw_smem[];                      // [4][WC + 2][DW_SLAB]: ring of input rows
    const int C4 = C >> 2;
    const int c = threadIdx.x % DW_SLAB, xl = threadIdx.x / DW_SLAB;
    const int c4 = blockIdx.y * DW_SLAB + c;                 // this thread's float4 channel group
    const bool live = c4 < C4;
    const int bands = (H + DW_ROWS - 1) / DW_ROWS;
    const int img = blockIdx.x / bands, band = blockIdx.x - img * bands;
    const int y0 = band * DW_ROWS, y1 = min(y0 + DW_ROWS, H);
    const int64_t img0 = static_cast<int64_t>(img) * H * W;  // first token of this image
    const int cx0 = blockIdx.z * DW_XC;                      // first output column of this block
    const int WC = min(DW_XC, W - cx0);                      // output columns of this block
    const int WP = WC + 2;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = zero4;
    if (HAS_LN && live) { g4 = __ldg(reinterpret_cast<const float4*>(gamma) + c4); b4 = __ldg(reinterpret_cast<const float4*>(beta) + c4); }
    float4 wk[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wk[k] = live ? __ldg(reinterpret_cast<const float4*>(w + k * C) + c4) : zero4;
    const float4 sc = live ? __ldg(reinterpret_cast<const float4*>(scale) + c4) : zero4;
    const float4 sh = live ? __ldg(reinterpret_cast<const float4*>(shift) + c4) : zero4;

    float4 rv[DW_MAXCOL];                                    // one input row's columns of this thread, in flight
    float2 rs[DW_MAXCOL];
    auto row_issue = [&](int yy) {                           // start the loads of input row yy (zeros outside the image)
        const bool rowok = live && yy >= 0 && yy < H;
        const int64_t rtok = img0 + static_cast<int64_t>(yy) * W;
#pragma unroll
        for (int k = 0; k < DW_MAXCOL; ++k) {
            const int col = xl + k * DW_XL, xx = cx0 + col - 1;      // tile column col holds input column xx
            rv[k] = zero4; rs[k] = make_float2(0.f, 0.f);
            if (rowok && col < WP && xx >= 0 && xx < W) {
                rv[k] = __ldg(reinterpret_cast<const float4*>(in + (rtok + xx) * ld_in + c_in) + c4);
                if (HAS_LN) rs[k] = __ldg(reinterpret_cast<const float2*>(stats) + rtok + xx);
            }
        }
    };
    auto row_commit = [&](int yy) {                          // normalise and store the row into ring slot yy & 3
        const bool rowok = live && yy >= 0 && yy < H;
        float4* trow = dw_smem + ((yy + 4) & 3) * WP * DW_SLAB + c;
#pragma unroll
        for (int k = 0; k < DW_MAXCOL; ++k) {
            const int col = xl + k * DW_XL, xx = cx0 + col - 1;
            if (col < WP) {
                float4 v = rv[k];
                if (HAS_LN && rowok && xx >= 0 && xx < W) {   // (zero padding applies to the normalised tensor)
                    const float mu = rs[k].x, r_ = rs[k].y;
                    v.x = (v.x - mu) * r_ * g4.x + b4.x; v.y = (v.y - mu) * r_ * g4.y + b4.y;
                    v.z = (v.z - mu) * r_ * g4.z + b4.z; v.w = (v.w - mu) * r_ * g4.w + b4.w;
                }
                trow[col * DW_SLAB] = v;
            }
        }
    };
    row_issue(y0 - 1); row_commit(y0 - 1);
    row_issue(y0);     row_commit(y0);
    row_issue(y0 + 1); row_commit(y0 + 1);
    __syncthreads();
    for (int y = y0; y < y1; ++y) {
        const bool more = y + 1 < y1;
        if (more) row_issue(y + 2);                          // in flight while row y is computed
        float4 gv[DW_XC / DW_XL];                            // this row's gate operands: loaded up front, not one exposed latency per output
        if (HAS_GATE && live) {
#pragma unroll
            for (int k = 0; k < DW_XC / DW_XL; ++k) {
                const int x = xl + k * DW_XL;
                gv[k] = x < WC ? __ldg(reinterpret_cast<const float4*>(gate + (img0 + static_cast<int64_t>(y) * W + cx0 + x) * ld_gate + c_gate) + c4)
                               : zero4;
            }
        }
        if (live) {
            const float4* r0 = dw_smem + ((y + 3) & 3) * WP * DW_SLAB + c;      // rows y - 1, y, y + 1
            const float4* r1 = dw_smem + (y & 3) * WP * DW_SLAB + c;
            const float4* r2 = dw_smem + ((y + 1) & 3) * WP * DW_SLAB + c;
            const int64_t rtok = img0 + static_cast<int64_t>(y) * W + cx0;
#pragma unroll
            for (int k = 0; k < DW_XC / DW_XL; ++k) {
                const int x = xl + k * DW_XL;
                if (x >= WC) break;
                float4 acc = zero4;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float4* rr = r == 0 ? r0 : (r == 1 ? r1 : r2);
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {         // taps in (dy, dx) row-major order, as the 9-load version summed them
                        const float4 v = rr[(x + dx) * DW_SLAB];
                        const float4 ww = wk[r * 3 + dx];
                        acc.x = fmaf(ww.x, v.x, acc.x); acc.y = fmaf(ww.y, v.y, acc.y); acc.z = fmaf(ww.z, v.z, acc.z); acc.w = fmaf(ww.w, v.w, acc.w);
                    }
                }
                const int64_t tok = rtok + x;
                float4 o = make_float4(fmaf(acc.x, sc.x, sh.x), fmaf(acc.y, sc.y, sh.y), fmaf(acc.z, sc.z, sh.z), fmaf(acc.w, sc.w, sh.w));
                if (act_gelu) { o.x = gelu_erf_fast(o.x); o.y = gelu_erf_fast(o.y); o.z = gelu_erf_fast(o.z); o.w = gelu_erf_fast(o.w); }
                if (HAS_GATE) { o.x *= gv[k].x; o.y *= gv[k].y; o.z *= gv[k].z; o.w *= gv[k].w; }
                if (out_planes != nullptr) {
                    // bf16 planes [channel / 64][token][128 B], 16-byte chunks permuted by chunk ^ (token & 7): the A-operand layout
                    // token_linear_kernel reads with one bulk copy per k-atom (SRK_LIN_A_PLANES); this thread's 4 channels = 8 bytes
                    const int ch = 4 * c4, pl = ch >> 6, chunk = (ch & 63) >> 3;
                    uint8_t* dst = out_planes + pl * plane_stride + tok * 128 + (((chunk ^ static_cast<int>(tok & 7)) << 4) | ((ch & 4) << 1));
                    __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                    *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
                } else {
                    reinterpret_cast<float4*>(out + tok * ld_out)[c4] = o;
                }
            }
        }
        if (more) {
            // slot (y + 2) & 3 held row y - 2: nobody reads it in this iteration, so no barrier is needed before the write
            row_commit(y + 2);
            __syncthreads();
        }
    }
}

// ---- LayerNorm statistics of a channel slice: stats[tok] = (mean, rstd); one warp per token
__global__ void __launch_bounds__(256) row_stats_kernel(const float* __restrict__ in, int ld_in, int c_in, int C, int64_t tokens,
                                                        float eps, float* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int C4 = C >> 2;
    for (int64_t tok = warp; tok < tokens; tok += nwarps) {
        const float4* src = reinterpret_cast<const float4*>(in + tok * ld_in + c_in);
        float s = 0.f, q = 0.f;
        for (int c = lane; c < C4; c += 32) {
            const float4 v = __ldg(src + c);
            s += (v.x + v.y) + (v.z + v.w);
            q = fmaf(v.x, v.x, q); q = fmaf(v.y, v.y, q); q = fmaf(v.z, v.z, q); q = fmaf(v.w, v.w, q);
        }
        s = warp_sum(s); q = warp_sum(q);
        if (lane == 0) {
            const float mean = s / C;
            const float var = fmaxf(q / C - mean * mean, 0.f);
            stats[2 * tok] = mean;
            stats[2 * tok + 1] = rsqrtf(var + eps);
        }
    }
}

// ---- adaptive interaction module.  s(tok) = w2 . gelu(W1 src(tok) + b1) + b2 with src = att (mode 0) or conv (mode 1);
//      mode 0 (spatial block, dat_arch.py:420-433): mix = att * sigmoid(cmap[b]) + sigmoid(s) * conv
//      mode 1 (channel block, dat_arch.py:510-523): mix = att * sigmoid(s) + conv * sigmoid(cmap[b])
//      One warp per token; W1 (hidden x 180, BatchNorm folded) lives in shared memory.
//      Lane = token for the squeeze MLP: a warp stages 32 source rows in shared memory (coalesced), each lane then runs the
//      180 -> hidden matvec of ITS token with the weights broadcast from shared memory -- no cross-lane reductions and one GELU
//      per (token, unit).  The first version (warp = token, 11 warp reductions and 11 redundant GELUs per token) spent ~600
//      instructions per token and was issue bound at 82 us; the output pass is coalesced again (lane = float4 of a row).
constexpr int MIX_MAX_HIDDEN = 16;
constexpr int MIX_WARPS = 4;                       // warps per block, 32 tokens each
constexpr int MIX_STRIDE = 188;                    // floats per staged row: 16-byte aligned, conflict-free for lane-per-row LDS.128
__global__ void __launch_bounds__(32 * MIX_WARPS) dat_mix_kernel(const float* __restrict__ att, const float* __restrict__ conv,
                                                                 const float* __restrict__ cmap, const float* __restrict__ w1,
                                                                 const float* __restrict__ b1, const float* __restrict__ w2, float b2, int hidden,
                                                                 int mode, float* __restrict__ mix, int64_t tokens, int tokens_per_image) {
    extern __shared__ __align__(16) float mix_smem[];
    float* s_w1t = mix_smem;                                               // [180][16]: W1 transposed, hidden padded to 16 (zeros)
    float* s_b1 = s_w1t + SRK_DIM * MIX_MAX_HIDDEN;                        // [16]
    float* s_w2 = s_b1 + MIX_MAX_HIDDEN;                                   // [16]
    float* s_rows = s_w2 + MIX_MAX_HIDDEN;                                 // [MIX_WARPS][32][MIX_STRIDE]
    for (int i = threadIdx.x; i < SRK_DIM * MIX_MAX_HIDDEN; i += blockDim.x) {
        const int c = i / MIX_MAX_HIDDEN, j = i - c * MIX_MAX_HIDDEN;
        s_w1t[i] = j < hidden ? w1[j * SRK_DIM + c] : 0.f;
    }
    if (threadIdx.x < MIX_MAX_HIDDEN) {
        s_b1[threadIdx.x] = threadIdx.x < hidden ? b1[threadIdx.x] : 0.f;
        s_w2[threadIdx.x] = threadIdx.x < hidden ? w2[threadIdx.x] : 0.f;      // padded units contribute w2 = 0
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* rows = s_rows + wib * 32 * MIX_STRIDE;
    constexpr int C4 = SRK_DIM / 4;                       // 45 float4 per row
    const bool has2 = lane + 32 < C4;
    const float* src = mode == 0 ? att : conv;            // the operand the squeeze MLP looks at
    const int64_t ngroups = (tokens + 31) >> 5;
    for (int64_t grp = static_cast<int64_t>(blockIdx.x) * MIX_WARPS + wib; grp < ngroups; grp += static_cast<int64_t>(gridDim.x) * MIX_WARPS) {
        const int64_t tok0 = grp << 5;
        const int nt = static_cast<int>(min(static_cast<int64_t>(32), tokens - tok0));
        // ---- stage the 32 source rows (coalesced: lane = float4 of the row)
        //      with cp.async: all 32 rows in flight at once, no register round trip (four batches of eight loads before: DAT x2
        //      21.04 -> 20.65 ms/step).  Also staging the OTHER operand's rows this way (no global loads in the output pass, but
        //      204 KB per CTA = 4 warps per SM instead of 8) was slower: 20.94.
        {
            const uint32_t rows_u = static_cast<uint32_t>(__cvta_generic_to_shared(rows));
            for (int t = 0; t < nt; ++t) {
                const float4* r4 = reinterpret_cast<const float4*>(src + (tok0 + t) * SRK_DIM);
                const uint32_t d = rows_u + static_cast<uint32_t>(t * MIX_STRIDE * 4 + lane * 16);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(r4 + lane) : "memory");
                if (has2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 512u), "l"(r4 + lane + 32) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        // ---- lane = token: hidden pre-activations
        float acc[MIX_MAX_HIDDEN];
#pragma unroll
        for (int j = 0; j < MIX_MAX_HIDDEN; ++j) acc[j] = 0.f;
        if (lane < nt) {
            const float4* x4 = reinterpret_cast<const float4*>(rows + lane * MIX_STRIDE);
#pragma unroll 3
            for (int c4 = 0; c4 < C4; ++c4) {
                const float4 xv = x4[c4];
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4* wv = reinterpret_cast<const float4*>(s_w1t + (4 * c4 + e) * MIX_MAX_HIDDEN);      // broadcast
#pragma unroll
                    for (int q = 0; q < MIX_MAX_HIDDEN / 4; ++q) {
                        const float4 ww = wv[q];
                        acc[4 * q] = fmaf(ww.x, xs[e], acc[4 * q]); acc[4 * q + 1] = fmaf(ww.y, xs[e], acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(ww.z, xs[e], acc[4 * q + 2]); acc[4 * q + 3] = fmaf(ww.w, xs[e], acc[4 * q + 3]);
                    }
                }
            }
        }
        float sacc = b2;
#pragma unroll
        for (int j = 0; j < MIX_MAX_HIDDEN; ++j) sacc = fmaf(s_w2[j], gelu_erf_fast(acc[j] + s_b1[j]), sacc);
        const float sg_mine = sigmoidf_(sacc);            // gate of token tok0 + lane
        // ---- output pass (coalesced): lane = float4 `lane` / `lane + 32` of every row of the group
        const float4* m4 = reinterpret_cast<const float4*>(cmap + (tok0 / tokens_per_image) * SRK_DIM);
        const int64_t img_end = (tok0 / tokens_per_image + 1) * tokens_per_image;     // a group may straddle two images
        float4 cg0, cg1 = make_float4(0.f, 0.f, 0.f, 0.f);
        auto load_cg = [&](const float4* m) {
            const float4 cm = __ldg(m + lane);
            cg0 = make_float4(sigmoidf_(cm.x), sigmoidf_(cm.y), sigmoidf_(cm.z), sigmoidf_(cm.w));
            if (has2) {
                const float4 cn = __ldg(m + lane + 32);
                cg1 = make_float4(sigmoidf_(cn.x), sigmoidf_(cn.y), sigmoidf_(cn.z), sigmoidf_(cn.w));
            }
        };
        load_cg(m4);
        const float* oth = mode == 0 ? conv : att;        // the operand that is not staged
        // mode 0: mix = att * cg + sg * conv (staged = att); mode 1: mix = att * sg + conv * cg (staged = conv)
        auto emit = [&](int t) {
            const int64_t tok = tok0 + t;
            const float sg = __shfl_sync(0xffffffffu, sg_mine, t);
            const float4* s4 = reinterpret_cast<const float4*>(rows + t * MIX_STRIDE);
            const float4* o4 = reinterpret_cast<const float4*>(oth + tok * SRK_DIM);
            float4* y4 = reinterpret_cast<float4*>(mix + tok * SRK_DIM);
            {
                const float4 a = s4[lane], b = __ldg(o4 + lane);
                y4[lane] = make_float4(fmaf(a.x, cg0.x, sg * b.x), fmaf(a.y, cg0.y, sg * b.y), fmaf(a.z, cg0.z, sg * b.z), fmaf(a.w, cg0.w, sg * b.w));
            }
            if (has2) {
                const float4 a = s4[lane + 32], b = __ldg(o4 + lane + 32);
                y4[lane + 32] = make_float4(fmaf(a.x, cg1.x, sg * b.x), fmaf(a.y, cg1.y, sg * b.y), fmaf(a.z, cg1.z, sg * b.z), fmaf(a.w, cg1.w, sg * b.w));
            }
        };
        if (nt == 32 && tok0 + 32 <= img_end) {           // the common case, unrolled: 8 rows of the other operand in flight per lane
#pragma unroll 8
            for (int t = 0; t < 32; ++t) emit(t);
        } else {
            for (int t = 0; t < nt; ++t) {
                if (tok0 + t == img_end) load_cg(m4 + C4); // (warp-uniform) next image's channel map
                emit(t);
            }
        }
        __syncwarp();
    }
}

// ---- channel attention statistics.  qkv rows [tok][540] = q | k | v.  One block per (token chunk, head, image): 64 tokens of
//      q, k staged in shared memory; gram[b][h] = 30x30 products, then 30 |q_d|^2, then 30 |k_d|^2 (per-block partials + ordered reduce).
constexpr int GRAM_TOK = 32;
constexpr int GRAM_STRIDE = SRK_HEAD_DIM * SRK_HEAD_DIM + 2 * SRK_HEAD_DIM;      // 960 floats per (image, head)
// thread (head, 5x5 tile of the 30x30 products): 10 shared-memory reads per 25 FMAs; 6 heads x 36 tiles = 216 threads
constexpr int GRAM_CHUNKS = 4;             // token chunks per block: partial sums stay in registers across them (8: too few blocks, 104 vs 77 us; 2: too many partials)
__global__ void __launch_bounds__(224) channel_gram_kernel(const float* __restrict__ qkv, float* __restrict__ part, int tokens_per_image) {
    __shared__ float s_qk[GRAM_TOK][2 * SRK_DIM + 4];            // q | k rows of the token chunk
    const int b = blockIdx.y;
    const int i = threadIdx.x;
    const bool worker = i < SRK_HEADS * 36;
    const int h = i / 36, tile = i - h * 36, r0 = (tile / 6) * 5, c0 = (tile % 6) * 5;
    const int qo = h * SRK_HEAD_DIM + r0, ko = SRK_DIM + h * SRK_HEAD_DIM + c0;
    float acc[5][5], nq[5], nk[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        nq[a] = 0.f; nk[a] = 0.f;
#pragma unroll
        for (int c = 0; c < 5; ++c) acc[a][c] = 0.f;
    }
    for (int ch = 0; ch < GRAM_CHUNKS; ++ch) {
        const int t0 = (blockIdx.x * GRAM_CHUNKS + ch) * GRAM_TOK;
        const int nt = min(GRAM_TOK, tokens_per_image - t0);
        if (nt <= 0) break;
        const float* base = qkv + (static_cast<int64_t>(b) * tokens_per_image + t0) * (3 * SRK_DIM);
        __syncthreads();
        for (int e = threadIdx.x; e < nt * (2 * SRK_DIM / 4); e += blockDim.x) {
            const int tk = e / (2 * SRK_DIM / 4), r = e - tk * (2 * SRK_DIM / 4);
            const float4 v = __ldg(reinterpret_cast<const float4*>(base + static_cast<int64_t>(tk) * (3 * SRK_DIM)) + r);
            *reinterpret_cast<float4*>(&s_qk[tk][4 * r]) = v;
        }
        __syncthreads();
        if (worker) {
            for (int tk = 0; tk < nt; ++tk) {
                float qv[5], kv[5];
#pragma unroll
                for (int a = 0; a < 5; ++a) { qv[a] = s_qk[tk][qo + a]; kv[a] = s_qk[tk][ko + a]; }
#pragma unroll
                for (int a = 0; a < 5; ++a) {
#pragma unroll
                    for (int c = 0; c < 5; ++c) acc[a][c] = fmaf(qv[a], kv[c], acc[a][c]);
                    nq[a] = fmaf(qv[a], qv[a], nq[a]);
                    nk[a] = fmaf(kv[a], kv[a], nk[a]);
                }
            }
        }
    }
    if (!worker) return;
    // one partial per (image, block): reduced in a fixed order by channel_gram_reduce_kernel (run-to-run identical results)
    float* dst = part + ((static_cast<int64_t>(b) * gridDim.x + blockIdx.x) * SRK_HEADS + h) * GRAM_STRIDE;
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int c = 0; c < 5; ++c) dst[(r0 + a) * SRK_HEAD_DIM + c0 + c] = acc[a][c];
    if (tile % 6 == 0) {            // first tile of a tile row owns the q norms of its 5 rows
#pragma unroll
        for (int a = 0; a < 5; ++a) dst[SRK_HEAD_DIM * SRK_HEAD_DIM + r0 + a] = nq[a];
    }
    if (tile / 6 == 0) {            // first tile of a tile column owns the k norms of its 5 columns
#pragma unroll
        for (int a = 0; a < 5; ++a) dst[SRK_HEAD_DIM * SRK_HEAD_DIM + SRK_HEAD_DIM + c0 + a] = nk[a];
    }
}

__global__ void __launch_bounds__(256) channel_gram_reduce_kernel(const float* __restrict__ part, float* __restrict__ gram, int nblk, int per_image) {
    // one element per thread, four independent partial sums (k mod 4) so that the loads of a thread overlap; fixed order
    const int b = blockIdx.y;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < per_image; e += gridDim.x * blockDim.x) {
        const float* src = part + static_cast<int64_t>(b) * nblk * per_image + e;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int k = 0;
        for (; k + 4 <= nblk; k += 4) {
            s0 += __ldg(src + static_cast<int64_t>(k) * per_image);     s1 += __ldg(src + static_cast<int64_t>(k + 1) * per_image);
            s2 += __ldg(src + static_cast<int64_t>(k + 2) * per_image); s3 += __ldg(src + static_cast<int64_t>(k + 3) * per_image);
        }
        for (; k < nblk; ++k) s0 += __ldg(src + static_cast<int64_t>(k) * per_image);
        gram[static_cast<int64_t>(b) * per_image + e] = (s0 + s1) + (s2 + s3);
    }
}

// ---- out[tok, h*30 + d1] = sum_d2 attn[b, h, d1, d2] * v[tok, h*30 + d2]; one block per (image, 32-token chunk), the image's six
//      30x30 matrices in shared memory
constexpr int APPLY_TOK = 32;
// warp = 32 tokens x one head: every lane keeps its token's 30 v values in registers and the head's 30x30 matrix is read from
// shared memory as warp-wide broadcasts (one wavefront per read); 6 warps per block = the 6 heads of the same 32 tokens
__global__ void __launch_bounds__(192) channel_apply_kernel(const float* __restrict__ qkv, const float* __restrict__ attn, float* __restrict__ out,
                                                            int tokens_per_image) {
    // matrix rows padded 30 -> 32 floats: a row is read as eight 16-byte broadcasts instead of thirty 4-byte ones (the kernel was
    // bound by the shared-memory instruction rate: one LDS per FMA)
    __shared__ __align__(16) float s_a[SRK_HEADS * SRK_HEAD_DIM * 32];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < SRK_HEADS * SRK_HEAD_DIM * 32; i += blockDim.x) {
        const int r = i >> 5, c = i & 31;                 // r = h * 30 + d1
        s_a[i] = c < SRK_HEAD_DIM ? attn[static_cast<int64_t>(b) * SRK_HEADS * SRK_HEAD_DIM * SRK_HEAD_DIM + r * SRK_HEAD_DIM + c] : 0.f;
    }
    __syncthreads();
    const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * APPLY_TOK + lane;
    if (t >= tokens_per_image) return;
    const int64_t tok = static_cast<int64_t>(b) * tokens_per_image + t;
    const float2* v2 = reinterpret_cast<const float2*>(qkv + tok * (3 * SRK_DIM) + 2 * SRK_DIM + h * SRK_HEAD_DIM);
    float v[32];
#pragma unroll
    for (int d = 0; d < SRK_HEAD_DIM / 2; ++d) { const float2 x = __ldg(v2 + d); v[2 * d] = x.x; v[2 * d + 1] = x.y; }
    v[30] = 0.f; v[31] = 0.f;
    const float4* ah = reinterpret_cast<const float4*>(s_a + h * SRK_HEAD_DIM * 32);
    float2* o2 = reinterpret_cast<float2*>(out + tok * SRK_DIM + h * SRK_HEAD_DIM);
#pragma unroll 2
    for (int d1 = 0; d1 < SRK_HEAD_DIM; d1 += 2) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {                     // same summation order as before (d2 ascending)
            const float4 w0 = ah[d1 * 8 + q], w1 = ah[(d1 + 1) * 8 + q];
            a0 = fmaf(w0.x, v[4 * q], a0); a0 = fmaf(w0.y, v[4 * q + 1], a0); a0 = fmaf(w0.z, v[4 * q + 2], a0); a0 = fmaf(w0.w, v[4 * q + 3], a0);
            a1 = fmaf(w1.x, v[4 * q], a1); a1 = fmaf(w1.y, v[4 * q + 1], a1); a1 = fmaf(w1.z, v[4 * q + 2], a1); a1 = fmaf(w1.w, v[4 * q + 3], a1);
        }
        o2[d1 >> 1] = make_float2(a0, a1);
    }
}

// ------------------------------------------------------------------------------------------------ launchers
static int grid_for(int64_t threads_needed) {
    const int64_t blocks = (threads_needed + 255) / 256;
    return static_cast<int>(blocks < 148 * 16 ? (blocks > 0 ? blocks : 1) : 148 * 16);
}

cudaError_t launch_dwconv3x3_rows(const float* in, int ld_in, int c_in, const float* w, const float* scale, const float* shift,
                                  const float* stats, const float* gamma, const float* beta, const float* gate, int ld_gate, int c_gate,
                                  float* out, int ld_out, int C, int batch, int H, int W, int act_gelu, cudaStream_t stream,
                                  uint8_t* out_planes, long long plane_stride) {
    if (batch <= 0) return cudaSuccess;
    const int C4 = C >> 2;
    const size_t smem = static_cast<size_t>(4) * (DW_XC + 2) * DW_SLAB * sizeof(float4);
    const dim3 grid(batch * ((H + DW_ROWS - 1) / DW_ROWS), (C4 + DW_SLAB - 1) / DW_SLAB, (W + DW_XC - 1) / DW_XC);
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        kern<<<grid, DW_SLAB * DW_XL, smem, stream>>>(in, ld_in, c_in, w, scale, shift, stats, gamma, beta, gate, ld_gate, c_gate, out, ld_out, C, H,
                                                    W, act_gelu, out_planes, plane_stride);
        return cudaSuccess;
    };
    cudaError_t e;
    if (stats) e = gate ? go(dwconv3x3_rows_kernel<true, true>) : go(dwconv3x3_rows_kernel<true, false>);
    else       e = gate ? go(dwconv3x3_rows_kernel<false, true>) : go(dwconv3x3_rows_kernel<false, false>);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t launch_row_stats(const float* in, int ld_in, int c_in, int C, int64_t tokens, float eps, float* stats, cudaStream_t stream) {
    if (tokens <= 0) return cudaSuccess;
    row_stats_kernel<<<grid_for(tokens * 32), 256, 0, stream>>>(in, ld_in, c_in, C, tokens, eps, stats);
    return cudaGetLastError();
}

cudaError_t launch_dat_mix(const float* att, const float* conv, const float* cmap, const float* w1, const float* b1, const float* w2,
                           float b2, int hidden, int mode, float* mix, int64_t tokens, int tokens_per_image, cudaStream_t stream) {
    if (tokens <= 0) return cudaSuccess;
    const size_t smem = (static_cast<size_t>(SRK_DIM) * MIX_MAX_HIDDEN + 2 * MIX_MAX_HIDDEN + MIX_WARPS * 32 * MIX_STRIDE) * sizeof(float);
    static bool configured[SRK_MAX_DEVICES] = {};
    if (cudaError_t e = configure_smem_once(configured, dat_mix_kernel, static_cast<int>(smem)); e != cudaSuccess) return e;
    const int64_t groups = (tokens + 31) / 32;
    const int64_t blocks = (groups + MIX_WARPS - 1) / MIX_WARPS;
    const int grid = static_cast<int>(blocks < 148 * 2 ? blocks : 148 * 2);       // 2 CTAs of 108 KB per SM, persistent over token groups
    dat_mix_kernel<<<grid, 32 * MIX_WARPS, smem, stream>>>(att, conv, cmap, w1, b1, w2, b2, hidden, mode, mix, tokens, tokens_per_image);
    return cudaGetLastError();
}

static int gram_blocks(int tokens_per_image) { return (tokens_per_image + GRAM_TOK * GRAM_CHUNKS - 1) / (GRAM_TOK * GRAM_CHUNKS); }
int channel_gram_ws_floats(int batch, int tokens_per_image) { return batch * gram_blocks(tokens_per_image) * SRK_HEADS * GRAM_STRIDE; }

cudaError_t launch_channel_gram(const float* qkv, float* gram, float* ws, int batch, int tokens_per_image, cudaStream_t stream) {
    if (batch <= 0 || tokens_per_image <= 0) return cudaSuccess;
    const int nblk = gram_blocks(tokens_per_image);
    channel_gram_kernel<<<dim3(nblk, batch), 224, 0, stream>>>(qkv, ws, tokens_per_image);
    channel_gram_reduce_kernel<<<dim3((SRK_HEADS * GRAM_STRIDE + 255) / 256, batch), 256, 0, stream>>>(ws, gram, nblk, SRK_HEADS * GRAM_STRIDE);
    return cudaGetLastError();
}

// ---- attn[b][h] = softmax_d2( gram[d1][d2] / (max(|q_d1|, eps) max(|k_d2|, eps)) * temperature[h] ): the F.normalize of q and k over
//      the tokens folded into the 30 x 30 logits (dat_arch.py:497-503).  One warp per (image, head), lane = row d1.
__global__ void __launch_bounds__(32) channel_softmax_kernel(const float* __restrict__ gram, const float* __restrict__ temperature,
                                                             float* __restrict__ attn) {
    constexpr int D = SRK_HEAD_DIM;
    const int h = blockIdx.x, b = blockIdx.y, i = threadIdx.x;
    const float* g = gram + (static_cast<int64_t>(b) * SRK_HEADS + h) * (D * D + 2 * D);
    __shared__ float s_nk[D];
    if (i < D) s_nk[i] = fmaxf(sqrtf(g[D * D + D + i]), 1e-12f);
    __syncwarp();
    if (i >= D) return;
    const float nq = fmaxf(sqrtf(g[D * D + i]), 1e-12f), t = __ldg(temperature + h);
    float l[D], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        l[j] = g[i * D + j] / (nq * s_nk[j]) * t;
        mx = fmaxf(mx, l[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) { l[j] = expf(l[j] - mx); sum += l[j]; }
    float* o = attn + ((static_cast<int64_t>(b) * SRK_HEADS + h) * D + i) * D;
#pragma unroll
    for (int j = 0; j < D; ++j) o[j] = l[j] / sum;
}

cudaError_t launch_channel_softmax(const float* gram, const float* temperature, float* attn, int batch, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    channel_softmax_kernel<<<dim3(SRK_HEADS, batch), 32, 0, stream>>>(gram, temperature, attn);
    return cudaGetLastError();
}

cudaError_t launch_channel_apply(const float* qkv, const float* attn, float* out, int batch, int tokens_per_image, cudaStream_t stream) {
    if (batch <= 0 || tokens_per_image <= 0) return cudaSuccess;
    dim3 grid((tokens_per_image + APPLY_TOK - 1) / APPLY_TOK, batch);
    channel_apply_kernel<<<grid, 192, 0, stream>>>(qkv, attn, out, tokens_per_image);
    return cudaGetLastError();
}

}  // namespace srk
