// 3x3 convolution (stride 1, zero padding 1) as an implicit GEMM on tcgen05 for sm_100a, with the bias / activation /
// residual / pixel-shuffle tail fused into the epilogue.  Replaces the library (cuDNN) convolutions of the three networks:
// network_swinir.py:465 (RSTB conv), :729 (conv_after_body), :720 (conv_first), :742-745 (conv_before_upsample, Upsample,
// conv_last); hat_arch.py:67-72 (CAB) and the same tails in hat_arch.py / dat_arch.py.
//
//   out[b, y, x, n] = act(bias[n] + sum_{dy, dx, c} in[b, y + dy - 1, x + dx - 1, c] * w[n, c, dy, dx])  (+ residual)
//
// Layout.  Activations are fp16 NHWC with the channel count padded to a multiple of 64 (one "k-atom" = 64 channels = 128 B per
// pixel); fp16 (11-bit significand, like the TF32 the library used) rather than bf16 because the tail convolutions produce
// pixels directly.  Weights arrive pre-packed (packing.pack_conv3x3) as k_atoms x 3 (dx) x 3 (dy) slabs of NP rows x 128 B,
// 128-byte swizzled (NP = padded C_out, a multiple of 16).
//
// One persistent CTA per SM; work unit = a PATCH of TH x TW = 256 output pixels of one image (two M = 128 accumulators, so every
// weight slab fetched from L2 is used by 256 pixels).  For each (k-atom, dx) ONE tensor-map TMA load (cp.async.bulk.tensor.4d,
// SASS UTMALDG) brings the (TH + 2) x TW x 64-channel input box at column offset dx - 1, rows y0 - 1 .., into shared memory as
// a 128-byte-swizzled K-major operand image; out-of-image pixels are zero-filled by the TMA unit, which IS the zero padding of
// the convolution.  The three dy taps read the same box at row offsets dy * TW (TW is a multiple of 8, so the offsets keep the
// 1024-byte swizzle phase) -- 3 box loads instead of 9 tile loads per k-atom.  Per tap one 1-D bulk copy streams the weight slab.
//   warp 0: box producer   warp 1: weight producer   warp 2: tcgen05.mma issuer (warp-uniform, umma.cuh)   warp 3: TMEM allocation
//   warps 4..11: epilogue (warp w: accumulator (w - 4) / 4, TMEM lane quadrant w % 4): TMEM -> +bias -> activation ->
//   32 x 32 transposes through a private 4 KB shared-memory tile -> 128-byte coalesced global stores.
// Output modes: fp32 token rows (optionally += residual, which may alias the output: the RSTB / long-skip adds), fp16 NHWC
// planes for a following convolution, fp16 NHWC of the 2x pixel-shuffled image (Upsample: conv + nn.PixelShuffle(2) in one
// pass, the weight rows are permuted at pack time), or the 3-channel image itself.
// Algorithmic work: 2 * 9 * C_in * C_out FLOP per output pixel (583 200 per token for 180 -> 180).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#define SRK_OOL_TIMEOUT 1
#include "umma.cuh"

namespace srk {

constexpr int CONV_THREADS = 384;
constexpr uint32_t CV_ABOX = 49152;                  // largest box: (4 + 2) x 64 pixels x 128 B
constexpr uint32_t CV_A = 0;                         // 2 box stages
constexpr uint32_t CV_W = 2 * CV_ABOX;               // weight ring: up to 4 slabs, 96 KB
constexpr uint32_t CV_WBYTES = 98304;
constexpr uint32_t CV_STAGE = CV_W + CV_WBYTES;      // 8 epilogue warps x 4 KB transpose tiles
constexpr uint32_t CV_BIAS = CV_STAGE + 8 * 4096;    // 256 floats
constexpr uint32_t CV_BAR = CV_BIAS + 1024;
constexpr uint32_t CV_END = CV_BAR + 256;
constexpr uint32_t CONV_SMEM = CV_END + 1024;        // + alignment slack
static_assert(CONV_SMEM <= 232448, "conv kernel shared memory exceeds 227 KB");
enum { CB_AFULL = 0, CB_AEMPTY = 2, CB_WFULL = 4, CB_WEMPTY = 8, CB_ACCFULL = 12, CB_ACCEMPTY = 13, CB_COUNT = 14 };

struct ConvParams {
    alignas(64) CUtensorMap tmap;   // input: fp16 (B, H, W, 64 k_atoms), box {64, TW, TH + 2, 1}, SWIZZLE_128B, zero fill
    const uint8_t* wstream;
    const float* bias;              // NP floats
    float* out_f32;
    __half* out_f16;
    const float* residual;          // fp32 rows like out_f32 (may alias it) or nullptr
    int H, W, B;
    int tw_log2, th;                // patch = th rows x (1 << tw_log2) columns = 256 pixels
    int patches_x, patches_y, n_patches;
    int k_atoms, a_atoms, np, cout;     // k_atoms k-steps of 64 channels over an image of a_atoms atoms (hi/lo split convolutions)
    int out_mode, ld_out, act;
    float slope;
    uint32_t idesc, box_bytes, slab_bytes;
    int n_wstages;
    unsigned long long* dbg;        // optional timeline buffer (srk_debug_set_timeline): CTA 0's clock64() stamps of its first patches
};
#define CV_TL(it, id) do { if (p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && (it) < 8) p.dbg[(it) * 64 + (id)] = clock64(); } while (0)

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float conv_act(float v, int act, float slope) {
    if (act == SRK_ACT_LEAKY_RELU) return v > 0.f ? v : v * slope;
    if (act == SRK_ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));      // nn.GELU() (exact), hat_arch.py:68
    return v;
}

__global__ void __launch_bounds__(CONV_THREADS, 1) conv3x3_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - raw);
    float* s_bias = reinterpret_cast<float*>(sm + CV_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + CV_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + CB_COUNT + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TW = 1 << p.tw_log2;

    pdl_launch_dependents();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_bias[i] = i < p.np ? p.bias[i] : 0.f;       // constants: before the PDL wait
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&bars[CB_AFULL + i], 1); mbar_init(&bars[CB_AEMPTY + i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&bars[CB_WFULL + i], 1); mbar_init(&bars[CB_WEMPTY + i], 1); }
        mbar_init(&bars[CB_ACCFULL], 1);
        mbar_init(&bars[CB_ACCEMPTY], 8);
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmap)) : "memory");
    }
    if (warp == 3) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int n_steps = 3 * p.k_atoms;                  // (k-atom, dx) box loads per patch

    auto patch_geom = [&](int patch, int& b, int& y0, int& x0) {
        const int per_img = p.patches_x * p.patches_y;
        b = patch / per_img;
        const int r = patch - b * per_img;
        const int py = r / p.patches_x;
        y0 = py * p.th;
        x0 = (r - py * p.patches_x) * TW;
    };

    if (warp == 0) {
        // ===================================================== box producer (tensor-map TMA)
        pdl_wait();                                   // the input was written by the previous kernel
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int patch = blockIdx.x; patch < p.n_patches; patch += gridDim.x) {
                int b, y0, x0;
                patch_geom(patch, b, y0, x0);
                for (int s = 0; s < n_steps; ++s) {
                    const int ka = s / 3, dx = s - 3 * ka;
                    // input atom: the last k_atoms - a_atoms steps re-read the image's last atoms ([lo | hi] x [hi(w) | lo(w) | hi(w)])
                    const int ia = ka >= p.a_atoms ? ka - (p.k_atoms - p.a_atoms) : ka;
                    mbar_wait(&bars[CB_AEMPTY + stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars[CB_AFULL + stage], p.box_bytes);
                    tma_load_4d(sbase + CV_A + stage * CV_ABOX, &p.tmap, 64 * ia, x0 + dx - 1, y0 - 1, b, &bars[CB_AFULL + stage]);
                    stage ^= 1;
                    if (stage == 0) phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================== weight producer (constant slabs: no PDL wait)
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const int n_slabs = 3 * n_steps;
            for (int patch = blockIdx.x; patch < p.n_patches; patch += gridDim.x) {
                uint32_t off = 0;
                for (int s = 0; s < n_slabs; ++s) {
                    mbar_wait(&bars[CB_WEMPTY + stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars[CB_WFULL + stage], p.slab_bytes);
                    bulk_g2s(sm + CV_W + stage * p.slab_bytes, p.wstream + off, p.slab_bytes, &bars[CB_WFULL + stage]);
                    off += p.slab_bytes;
                    if (++stage == static_cast<uint32_t>(p.n_wstages)) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ===================================================== MMA issuer (whole warp, uniform operands: umma.cuh)
        uint32_t as = 0, aph = 0, ws = 0, wph = 0, acc_ph = 0;
        int it = 0;
        for (int patch = blockIdx.x; patch < p.n_patches; patch += gridDim.x, ++it) {
            CV_TL(it, 0);
            mbar_wait(&bars[CB_ACCEMPTY], acc_ph ^ 1);            // the epilogue has drained the previous patch's accumulators
            tc_fence_after();
            CV_TL(it, 1);
            long long wwait = 0;
            for (int s = 0; s < n_steps; ++s) {
                if (s < 12) CV_TL(it, 2 + 2 * s);
                mbar_wait(&bars[CB_AFULL + as], aph);
                tc_fence_after();
                if (s < 12) CV_TL(it, 3 + 2 * s);
                const uint32_t box = sbase + CV_A + as * CV_ABOX;
#pragma unroll 1
                for (int dy = 0; dy < 3; ++dy) {
                    const long long w0 = p.dbg ? clock64() : 0;
                    mbar_wait(&bars[CB_WFULL + ws], wph);
                    tc_fence_after();
                    if (p.dbg) wwait += clock64() - w0;
                    const uint64_t bd = umma_desc_sw128(sbase + CV_W + ws * p.slab_bytes);
                    const uint32_t a0 = box + static_cast<uint32_t>(dy * TW) * 128u;
                    const uint32_t accum = (s | dy) != 0;
                    umma_ss_w4(tmem, umma_desc_sw128(a0), bd, p.idesc, accum);
                    umma_ss_w4(tmem + 256, umma_desc_sw128(a0 + 16384u), bd, p.idesc, accum);
                    umma_commit_w(&bars[CB_WEMPTY + ws]);
                    if (++ws == static_cast<uint32_t>(p.n_wstages)) { ws = 0; wph ^= 1; }
                }
                umma_commit_w(&bars[CB_AEMPTY + as]);
                as ^= 1;
                if (as == 0) aph ^= 1;
            }
            umma_commit_w(&bars[CB_ACCFULL]);
            acc_ph ^= 1;
            CV_TL(it, 30);
            if (p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && it < 8) p.dbg[it * 64 + 31] = static_cast<unsigned long long>(wwait);
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================================================== epilogue
        const int mt = (warp - 4) >> 2, q = warp & 3;
        const uint32_t acc = tmem + static_cast<uint32_t>(256 * mt) + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* stg = sm + CV_STAGE + (warp - 4) * 4096;
        const uint32_t stg_u = sbase + CV_STAGE + (warp - 4) * 4096;
        uint32_t ph = 0;
        int it = 0;
        pdl_wait();                                   // residual / output buffers may still be in use by the previous kernel
        for (int patch = blockIdx.x; patch < p.n_patches; patch += gridDim.x, ++it) {
            int b, y0, x0;
            patch_geom(patch, b, y0, x0);
            const int R = 128 * mt + 32 * q + lane;
            const int y = y0 + (R >> p.tw_log2), x = x0 + (R & (TW - 1));
            const bool valid = y < p.H && x < p.W;
            const int pix = valid ? (b * p.H + y) * p.W + x : -1;                          // NHWC pixel index of this lane's row
            const int pix2 = valid ? (b * 2 * p.H + 2 * y) * 2 * p.W + 2 * x : -1;          // (2y, 2x) of the pixel-shuffled image
            mbar_wait(&bars[CB_ACCFULL], ph); ph ^= 1;
            tc_fence_after();
            if (warp == 4) CV_TL(it, 40);
            if (p.out_mode == SRK_CONV_OUT_IMAGE) {
                // ---- C_out <= 4 (conv_last): the row's lane stores its pixel directly (consecutive lanes = consecutive pixels)
                uint32_t v[16];
                tmem_ld16(acc, v);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k < p.cout) {
                            float o = conv_act(__uint_as_float(v[k]) + s_bias[k], p.act, p.slope);
                            const int64_t a = static_cast<int64_t>(pix) * p.ld_out + k;
                            if (p.residual) o += p.residual[a];
                            p.out_f32[a] = o;
                        }
                    }
                }
            } else {
                const int n_chunks = p.np >> 5;
#pragma unroll 1
                for (int c = 0; c < n_chunks; ++c) {
                    uint32_t v[32];
                    tmem_ld32(acc + 32 * c, v);
                    tmem_ld_wait();
                    float f[32];
                    {   // bias as eight 16-byte broadcasts; the activation switch sits OUTSIDE the element loops (a per-element
                        // runtime switch serialised 32 dependent LDS + branches per chunk: 3.8 K cycles per chunk, measured)
                        const float4* b4 = reinterpret_cast<const float4*>(s_bias + 32 * c);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float4 bb = b4[k];
                            f[4 * k] = __uint_as_float(v[4 * k]) + bb.x;         f[4 * k + 1] = __uint_as_float(v[4 * k + 1]) + bb.y;
                            f[4 * k + 2] = __uint_as_float(v[4 * k + 2]) + bb.z; f[4 * k + 3] = __uint_as_float(v[4 * k + 3]) + bb.w;
                        }
                        if (p.act == SRK_ACT_LEAKY_RELU) {
                            const float sl = p.slope;
#pragma unroll
                            for (int i = 0; i < 32; ++i) f[i] = f[i] > 0.f ? f[i] : f[i] * sl;
                        } else if (p.act == SRK_ACT_GELU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) f[i] = 0.5f * f[i] * (1.0f + erff(f[i] * 0.70710678118654752f));      // nn.GELU() (exact), hat_arch.py:68
                        }
                    }
                    if (p.out_mode == SRK_CONV_OUT_ROWS_F32) {
                        // transpose: lane = row -> 8 lanes per row, 128 B (32 floats) contiguous per row
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            *reinterpret_cast<float4*>(stg + lane * 128 + ((k ^ (lane & 7)) << 4)) = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
                        __syncwarp();
                        // all residual loads of the chunk first, then the stores: `residual` may alias `out` (in-place += ), and a
                        // load behind each store would serialise eight global round trips per chunk.  Every lane reads exactly the
                        // addresses it writes, so the order between lanes does not matter.
                        const int k = lane & 7, col = 32 * c + 4 * k;
                        int64_t addr[8];
                        float4 res[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int pr = __shfl_sync(0xffffffffu, pix, 4 * i + (lane >> 3));
                            addr[i] = (pr >= 0 && col < p.cout) ? static_cast<int64_t>(pr) * p.ld_out + col : static_cast<int64_t>(-1);
                            res[i] = (p.residual != nullptr && addr[i] >= 0) ? *reinterpret_cast<const float4*>(p.residual + addr[i])
                                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = 4 * i + (lane >> 3);
                            float4 o = *reinterpret_cast<const float4*>(stg + rr * 128 + ((k ^ (rr & 7)) << 4));
                            o.x += res[i].x; o.y += res[i].y; o.z += res[i].z; o.w += res[i].w;
                            if (addr[i] >= 0) *reinterpret_cast<float4*>(p.out_f32 + addr[i]) = o;
                        }
                        __syncwarp();
                    } else {
                        // fp16 outputs: 64 B (32 halves) per row and chunk; 4 lanes per row on the way out
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            st_shared_v4(stg_u + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4), pack_f16x2(f[8 * k], f[8 * k + 1]), pack_f16x2(f[8 * k + 2], f[8 * k + 3]),
                                         pack_f16x2(f[8 * k + 4], f[8 * k + 5]), pack_f16x2(f[8 * k + 6], f[8 * k + 7]));
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int rr = 8 * i + (lane >> 2), k = lane & 3;
                            const uint4 o = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((k ^ ((rr >> 1) & 3)) << 4));
                            if (p.out_mode == SRK_CONV_OUT_NHWC_F16) {
                                const int pr = __shfl_sync(0xffffffffu, pix, rr);
                                if (pr >= 0) *reinterpret_cast<uint4*>(p.out_f16 + static_cast<int64_t>(pr) * p.ld_out + 32 * c + 8 * k) = o;
                            } else {      // SRK_CONV_OUT_SHUFFLE2_F16: weight row n' = (2 i + j) * 64 + ch  ->  pixel (2y + i, 2x + j), channel ch
                                const int pr = __shfl_sync(0xffffffffu, pix2, rr);
                                const int sub = c >> 1;
                                if (pr >= 0)
                                    *reinterpret_cast<uint4*>(p.out_f16 + (static_cast<int64_t>(pr) + (sub & 1) + (sub >> 1) * 2 * p.W) * 64 + 32 * (c & 1) + 8 * k) = o;
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[CB_ACCEMPTY]);
            if (warp == 4) CV_TL(it, 41);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc(tmem, 512);
}

// ---- fp32 token rows (P, ld_in), C channels -> fp16 NHWC (P, cp) with zero padding (the A operand layout of conv3x3_kernel)
__global__ void __launch_bounds__(256) rows_to_f16_kernel(const float* __restrict__ x, int ld_in, int C, __half* __restrict__ out, int cp,
                                                          int64_t pixels) {
    const int groups = cp >> 3;
    const int64_t total = pixels * groups;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t pix = i / groups;
        const int c0 = static_cast<int>(i - pix * groups) * 8;
        float v[8];
        const float* src = x + pix * ld_in + c0;
        if (c0 + 8 <= C) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = c0 + k < C ? __ldg(src + k) : 0.f;
        }
        *reinterpret_cast<uint4*>(out + pix * cp + c0) = make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
    }
}

// ---- fp32 token rows -> the fp16 PAIR hi = fp16(v), lo = fp16(v - hi) of the tight mode's split convolutions (v = act(x)):
// conv(x, w) ~= hi(x) * hi(w) + lo(x) * hi(w) + hi(x) * lo(w) on the fp16 tensor-core kernel, fp32 accumulation; the dropped
// lo * lo term is 2^-22 relative.  `hi` and `lo` are NHWC images with row pitch ld_out: the callers pass lo = base, hi = base + cp,
// ld_out = 2 cp, and ONE launch of conv3x3_kernel walks k_atoms = 3 cp / 64 k-steps over that image's a_atoms = 2 cp / 64 atoms
// (the last cp / 64 steps re-read the hi atoms) against weights packed [hi(w) | lo(w) | hi(w)]: the two small products are
// accumulated first, the large one on top (the order matters: the tensor core's fp32 accumulation truncates).
// The activation of the PREVIOUS layer (conv_before_upsample's LeakyReLU, network_swinir.py:743) is applied on the way in, and
// with shuffle_h > 0 the rows are the 4 x 64 channels of a conv + nn.PixelShuffle(2) stage (weights packed pixel_shuffle=True,
// network_swinir.py:584-585): group s = 2 i + j of input pixel (y, x) goes to output pixel (2 y + i, 2 x + j).
__global__ void __launch_bounds__(256) rows_to_f16_split_kernel(const float* __restrict__ x, int ld_in, int C, __half* __restrict__ hi,
                                                                __half* __restrict__ lo, int ld_out, int cp, int64_t pixels,
                                                                int act, float slope, int sh, int sw) {
    const int groups = (sh > 0 ? 256 : cp) >> 3;
    const int64_t total = pixels * groups;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t pix = i / groups;
        int c0 = static_cast<int>(i - pix * groups) * 8;
        float v[8];
        const float* src = x + pix * ld_in + c0;
        if (c0 + 8 <= C) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = c0 + k < C ? __ldg(src + k) : 0.f;
        }
        int64_t opix = pix;
        if (sh > 0) {
            const int sub = c0 >> 6;
            c0 &= 63;
            const int64_t b = pix / (static_cast<int64_t>(sh) * sw);
            const int r = static_cast<int>(pix - b * sh * sw), y = r / sw, xx = r - y * sw;
            opix = (b * 2 * sh + 2 * y + (sub >> 1)) * 2 * sw + 2 * xx + (sub & 1);
        }
        float l[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (act == SRK_ACT_LEAKY_RELU) v[k] = v[k] > 0.f ? v[k] : v[k] * slope;
            else if (act == SRK_ACT_GELU) v[k] = 0.5f * v[k] * (1.0f + erff(v[k] * 0.70710678118654752f));
            const float h = __half2float(__float2half_rn(v[k]));
            l[k] = v[k] - h;
        }
        const uint4 H = make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
        const uint4 Lo = make_uint4(pack_f16x2(l[0], l[1]), pack_f16x2(l[2], l[3]), pack_f16x2(l[4], l[5]), pack_f16x2(l[6], l[7]));
        const int64_t o = opix * ld_out + c0;
        *reinterpret_cast<uint4*>(hi + o) = H;
        *reinterpret_cast<uint4*>(lo + o) = Lo;
    }
}

// ---- network input (B, C <= 3, H, W) fp32, any strides -> fp16 NHWC (P, 64) for conv_first (network_swinir.py:720, :803-804):
// v = (x - mean[c]) * range; channels [0, C) = hi(v), [C, 2C) = v - hi(v), [2C, 3C) = hi(v) again, rest 0.  With the weights packed
// as [hi(w), hi(w), w - hi(w)] the fp16 MMA computes hi*hi + lo*hi + hi*lo: the input and weight roundings cancel to second
// order (the 64-channel k-atom is mostly padding for a 3-channel input anyway, so the split is free).
__global__ void __launch_bounds__(256) image_to_f16_split_kernel(const float* __restrict__ x, int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                                                                 int C, int H, int W, int64_t pixels, float m0, float m1, float m2, float range,
                                                                 __half* __restrict__ out) {
    for (int64_t pix = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; pix < pixels; pix += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int xx = static_cast<int>(pix % W);
        const int64_t t = pix / W;
        const int yy = static_cast<int>(t % H);
        const int64_t b = t / H;
        __align__(16) __half h[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) h[k] = __float2half_rn(0.f);
        for (int c = 0; c < C; ++c) {
            const float v = (__ldg(x + b * sb + c * sc + yy * sy + xx * sx) - (c == 0 ? m0 : (c == 1 ? m1 : m2))) * range;
            const __half hi = __float2half_rn(v);
            h[c] = hi;
            h[C + c] = __float2half_rn(v - __half2float(hi));
            h[2 * C + c] = hi;
        }
        uint4* dst = reinterpret_cast<uint4*>(out + pix * 64);
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[k] = reinterpret_cast<const uint4*>(h)[k];
    }
}

// ------------------------------------------------------------------------------------------------ launchers
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    return fn;
}

// patch width: the power of two in {64, 32, 16, 8} that covers the image with the least padding (ties: the widest)
static void pick_patch(int H, int W, int& tw_log2, int& th) {
    long best = -1;
    for (int l = 6; l >= 3; --l) {
        const int tw = 1 << l, t = 256 / tw;
        const long area = static_cast<long>((W + tw - 1) / tw) * tw * ((H + t - 1) / t) * t;
        if (best < 0 || area < best) { best = area; tw_log2 = l; th = t; }
    }
}

cudaError_t launch_conv3x3(const ConvArgs& a, cudaStream_t stream) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return cudaErrorNotSupported;
    ConvParams p{};
    pick_patch(a.H, a.W, p.tw_log2, p.th);
    const int TW = 1 << p.tw_log2;
    const int a_atoms = a.a_atoms > 0 ? a.a_atoms : a.k_atoms;
    const cuuint64_t cp = 64ull * a_atoms;
    const cuuint64_t gdim[4] = {cp, static_cast<cuuint64_t>(a.W), static_cast<cuuint64_t>(a.H), static_cast<cuuint64_t>(a.B)};
    const cuuint64_t gstr[3] = {cp * 2, cp * 2 * a.W, cp * 2 * a.W * a.H};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(TW), static_cast<cuuint32_t>(p.th + 2), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(a.in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    p.wstream = a.wstream; p.bias = a.bias; p.out_f32 = a.out_f32; p.out_f16 = a.out_f16; p.residual = a.residual;
    p.H = a.H; p.W = a.W; p.B = a.B;
    p.patches_x = (a.W + TW - 1) / TW;
    p.patches_y = (a.H + p.th - 1) / p.th;
    p.n_patches = a.B * p.patches_x * p.patches_y;
    p.k_atoms = a.k_atoms; p.a_atoms = a_atoms; p.np = a.np; p.cout = a.cout;
    p.out_mode = a.out_mode; p.ld_out = a.ld_out; p.act = a.act; p.slope = a.slope;
    p.idesc = (1u << 4) | (static_cast<uint32_t>(a.np >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);      // kind::f16, A = B = fp16, D = fp32, K-major
    p.box_bytes = static_cast<uint32_t>((p.th + 2) * TW * 128);
    p.slab_bytes = static_cast<uint32_t>(a.np * 128);
    p.dbg = g_timeline;
    p.n_wstages = static_cast<int>(CV_WBYTES / p.slab_bytes) < 4 ? static_cast<int>(CV_WBYTES / p.slab_bytes) : 4;
    static bool configured[SRK_MAX_DEVICES] = {};
    if (cudaError_t e = configure_smem_once(configured, conv3x3_kernel, CONV_SMEM); e != cudaSuccess) return e;
    const int sms = device_num_sms();
    const int grid = p.n_patches < sms ? p.n_patches : sms;
    return launch_pdl(conv3x3_kernel, grid, CONV_THREADS, CONV_SMEM, stream, p);
}

cudaError_t launch_rows_to_f16(const float* x, int ld_in, int C, __half* out, int cp, int64_t pixels, cudaStream_t stream) {
    if (pixels <= 0) return cudaSuccess;
    const int64_t total = pixels * (cp >> 3);
    const int64_t blocks = (total + 255) / 256;
    const int grid = static_cast<int>(blocks < 148 * 16 ? blocks : 148 * 16);
    rows_to_f16_kernel<<<grid, 256, 0, stream>>>(x, ld_in, C, out, cp, pixels);
    return cudaGetLastError();
}

cudaError_t launch_rows_to_f16_split(const float* x, int ld_in, int C, __half* hi, __half* lo, int ld_out, int cp, int64_t pixels,
                                     int act, float slope, int shuffle_h, int shuffle_w, cudaStream_t stream) {
    if (pixels <= 0) return cudaSuccess;
    const int64_t total = pixels * ((shuffle_h > 0 ? 256 : cp) >> 3);
    const int64_t blocks = (total + 255) / 256;
    const int grid = static_cast<int>(blocks < 148 * 16 ? blocks : 148 * 16);
    rows_to_f16_split_kernel<<<grid, 256, 0, stream>>>(x, ld_in, C, hi, lo, ld_out, cp, pixels, act, slope, shuffle_h, shuffle_w);
    return cudaGetLastError();
}

cudaError_t launch_image_to_f16_split(const float* x, int64_t sb, int64_t sc, int64_t sy, int64_t sx, int C, int B, int H, int W,
                                      const float* mean3, float range, __half* out, cudaStream_t stream) {
    const int64_t pixels = static_cast<int64_t>(B) * H * W;
    if (pixels <= 0) return cudaSuccess;
    const int64_t blocks = (pixels + 255) / 256;
    const int grid = static_cast<int>(blocks < 148 * 16 ? blocks : 148 * 16);
    image_to_f16_split_kernel<<<grid, 256, 0, stream>>>(x, sb, sc, sy, sx, C, H, W, pixels, mean3[0], mean3[1], mean3[2], range, out);
    return cudaGetLastError();
}

}  // namespace srk
