// Fused Swin-block kernels for sm_100a (tcgen05 + TMEM + bulk TMA), see DESIGN.md.
//
//   swin_attn_kernel : LN1 -> cyclic shift + window partition (index math in the loads) -> qkv GEMM
//                      -> per head S = q k^T (+ relative position bias, + shifted-window mask),
//                      softmax, P v -> proj GEMM -> + shortcut -> window reverse / un-shift store.
//                      Replaces network_swinir.py:244-276 (and :114-145 in SRK_MODE_WINDOWS).
//   swin_mlp_kernel  : LN2 -> fc1 -> exact erf GELU -> fc2 -> + shortcut.  Replaces :277, :24-30.
//
// One CTA per SM, persistent over 128-token tiles (= two 8x8 windows).  Warp roles:
//   warp 0 lane 0 : weight producer -- streams the pre-swizzled bf16 weight slabs (packing.py) from
//                   L2 into a 3-stage shared-memory ring with 1-D bulk TMA;
//   warp 1 lane 0 : tcgen05.mma issuer (all GEMMs accumulate in TMEM);
//   warps 2..5    : 128 "row" threads: thread <-> TMEM lane <-> token row.  LayerNorm, operand
//                   images (bf16, 128-byte swizzle), softmax (a full 64-key row per thread, no
//                   shuffles), epilogues.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "umma.cuh"

namespace srk {

// ------------------------------------------------------------------------------------------------
// shared-memory plan (bytes from a 1024-aligned base)
// ------------------------------------------------------------------------------------------------
constexpr uint32_t ATOM_A = 16384;      // 128 rows x 128 B: one k-atom of a 128-row operand image
constexpr uint32_t VT_ATOM = 24576;     // 192 rows x 128 B: one k-atom (64 keys) of the V^T image
constexpr uint32_t RING_STAGE = 24576;  // largest weight slab (192 rows x 64 k)
constexpr int RING_N = 3;

// K1
constexpr uint32_t A_XA = 0;                          // LN1(x) image [128 x 192]; later the O image
constexpr uint32_t A_VT = A_XA + 3 * ATOM_A;          // V^T image [192 x 128 keys]
constexpr uint32_t A_QI = A_VT + 2 * VT_ATOM;         // Q image of the current head pair [128 x 64]
constexpr uint32_t A_KI = A_QI + ATOM_A;              // K image of the current head pair
constexpr uint32_t A_RING = A_KI + ATOM_A;            // weight ring
constexpr uint32_t A_VEC = A_RING + RING_N * RING_STAGE;
constexpr uint32_t A_BAR = A_VEC + ((SRK_ATTN_VEC_FLOATS * 4 + 127) / 128) * 128;
constexpr uint32_t A_END = A_BAR + 256;
constexpr uint32_t K1_SMEM = A_END + 1024;            // + alignment slack
static_assert(K1_SMEM <= 232448, "K1 shared memory exceeds 227 KB");
constexpr int STAGE_LD = 36;                          // fp32 words per row of the store transposer
static_assert(128 * STAGE_LD * 4 <= 2 * ATOM_A, "transposer must fit in the Q/K images");

// TMEM columns (fp32, 128 lanes)
constexpr uint32_t TC_VT0 = 0, TC_VT1 = 128;          // V^T accumulators (two 128-row halves of Wv)
constexpr uint32_t TC_S = 0;                          // S = q k^T, 128 x 128 (block diagonal is used)
constexpr uint32_t TC_P = 128;                        // P (bf16 pairs) as the A operand of P v, 64 cols
constexpr uint32_t TC_QK = 192;                       // [q(2p) q(2p+1) k(2p) k(2p+1)] of a head pair
constexpr uint32_t TC_O = 320;                        // O accumulator, 6 heads x 32
constexpr uint32_t TC_PROJ = 0;                       // proj accumulator, 192 cols

constexpr uint32_t IDESC_128x128 = umma_idesc_bf16(128, 128);
constexpr uint32_t IDESC_128x192 = umma_idesc_bf16(128, 192);
constexpr uint32_t IDESC_128x32 = umma_idesc_bf16(128, 32);

constexpr float LOG2E = 1.4426950408889634f;

enum {  // K1 barrier slots
    B_FULL = 0, B_EMPTY = 3, B_XA = 6, B_VTF0 = 7, B_VTF1 = 8, B_VTD = 9, B_QKF = 10, B_QKR = 11, B_SF = 12,
    B_PR = 13, B_OF = 14, B_OR = 15, B_PJF = 16, B_COUNT = 17
};

__device__ __forceinline__ int64_t window_token(const AttnParams& p, int gw, int t) {
    if (p.mode == SRK_MODE_WINDOWS) return static_cast<int64_t>(gw) * 64 + t;
    const int b = gw / p.nw_img, w = gw - b * p.nw_img;
    const int wy = w / p.nwx, wx = w - wy * p.nwx;
    int yy = wy * 8 + (t >> 3) + p.shift;
    if (yy >= p.H) yy -= p.H;
    int xx = wx * 8 + (t & 7) + p.shift;
    if (xx >= p.W) xx -= p.W;
    return (static_cast<int64_t>(b) * p.H + yy) * p.W + xx;
}

// LayerNorm (optional) of 128 gathered token rows -> bf16 SW128 image at `xa` (3 k-atoms).
// Half-warp per token: 16 lanes x 3 float4 cover the 180 channels (45 float4) fully coalesced.
template <typename TokFn>
__device__ __forceinline__ void ln_rows_to_image(const float* __restrict__ x, int ld, const float* s_w, const float* s_b,
                                                 int apply_ln, uint32_t xa, int cw, int lane, TokFn tok_of_row) {
    const int l16 = lane & 15;
#pragma unroll 4
    for (int pass = 0; pass < 16; ++pass) {
        const int r = cw * 32 + pass * 2 + (lane >> 4);
        const int64_t tok = tok_of_row(r);
        float4 v[3];
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
            const int f = l16 + 16 * jj;
            v[jj] = (tok >= 0 && f < SRK_DIM / 4) ? __ldg(reinterpret_cast<const float4*>(x + tok * ld) + f)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (apply_ln) {
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
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
                const int f = l16 + 16 * jj;
                if (f < SRK_DIM / 4) {
                    const float4 g = reinterpret_cast<const float4*>(s_w)[f];
                    const float4 b = reinterpret_cast<const float4*>(s_b)[f];
                    v[jj].x = v[jj].x * rstd * g.x + b.x; v[jj].y = v[jj].y * rstd * g.y + b.y;
                    v[jj].z = v[jj].z * rstd * g.z + b.z; v[jj].w = v[jj].w * rstd * g.w + b.w;
                }
            }
        }
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {   // channel 4f = 64*jj + 4*l16 -> atom jj, chunk l16>>1, byte (l16&1)*8
            const uint32_t addr = xa + jj * ATOM_A + sw128_off(r, l16 >> 1) + (l16 & 1) * 8;
            st_shared_v2(addr, pack_bf16x2(v[jj].x, v[jj].y), pack_bf16x2(v[jj].z, v[jj].w));
        }
    }
}

// 32 fp32 accumulators (+ per-column bias from smem, * scale) -> 4 x 16-byte bf16 chunks of one image row.
__device__ __forceinline__ void store_row_chunks(uint32_t img_atom, uint32_t row, uint32_t c16base, const uint32_t (&v)[32],
                                                 const float* bias, float scale) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = 8 * k + 2 * e;
            const float a = (__uint_as_float(v[i]) + (bias ? bias[i] : 0.f)) * scale;
            const float b = (__uint_as_float(v[i + 1]) + (bias ? bias[i + 1] : 0.f)) * scale;
            w[e] = pack_bf16x2(a, b);
        }
        st_shared_v4(img_atom + sw128_off(row, c16base + k), w[0], w[1], w[2], w[3]);
    }
}

// Final epilogue shared by K1/K2: 192 accumulator columns (+bias) -> fp32 rows, via a shared-memory
// transposer so that global loads (shortcut) and stores are coalesced (4 rows x 128 B per warp instruction).
template <typename TokFn>
__device__ __forceinline__ void store_rows_coalesced(uint32_t tmem_acc, uint32_t lanebase, float* stage, const float* s_bias,
                                                     const float* __restrict__ x, int ld_in, float* __restrict__ y, int ld_out,
                                                     int add_residual, int row, int cw, int lane, TokFn tok_of_row) {
    int64_t toks[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) toks[it] = tok_of_row(cw * 32 + it * 4 + (lane >> 3));
    const int c4 = lane & 7;
#pragma unroll 1
    for (int c = 0; c < 6; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_acc + lanebase + 32 * c, v);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float4 o;
            o.x = __uint_as_float(v[4 * k + 0]) + s_bias[32 * c + 4 * k + 0];
            o.y = __uint_as_float(v[4 * k + 1]) + s_bias[32 * c + 4 * k + 1];
            o.z = __uint_as_float(v[4 * k + 2]) + s_bias[32 * c + 4 * k + 2];
            o.w = __uint_as_float(v[4 * k + 3]) + s_bias[32 * c + 4 * k + 3];
            *reinterpret_cast<float4*>(stage + row * STAGE_LD + 4 * k) = o;
        }
        named_bar_sync(1, 128);
        const int ch = 32 * c + 4 * c4;
        if (ch < SRK_DIM) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = cw * 32 + it * 4 + (lane >> 3);
                const int64_t tok = toks[it];
                if (tok >= 0) {
                    float4 o = *reinterpret_cast<const float4*>(stage + rr * STAGE_LD + 4 * c4);
                    if (add_residual) {
                        const float4 s = __ldg(reinterpret_cast<const float4*>(x + tok * ld_in + ch));
                        o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
                    }
                    *reinterpret_cast<float4*>(y + tok * ld_out + ch) = o;
                }
            }
        }
        named_bar_sync(1, 128);
    }
}

// ------------------------------------------------------------------------------------------------
// K1: attention half
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192, 1) swin_attn_kernel(const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - raw);
    float* s_vec = reinterpret_cast<float*>(sm + A_VEC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + A_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < SRK_ATTN_VEC_FLOATS; i += blockDim.x) s_vec[i] = p.vec[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING_N; ++i) { mbar_init(&bars[B_FULL + i], 1); mbar_init(&bars[B_EMPTY + i], 1); }
        mbar_init(&bars[B_XA], 128);  mbar_init(&bars[B_VTF0], 1); mbar_init(&bars[B_VTF1], 1);
        mbar_init(&bars[B_VTD], 128); mbar_init(&bars[B_QKF], 1);  mbar_init(&bars[B_QKR], 128);
        mbar_init(&bars[B_SF], 1);    mbar_init(&bars[B_PR], 128); mbar_init(&bars[B_OF], 1);
        mbar_init(&bars[B_OR], 128);  mbar_init(&bars[B_PJF], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================================================== weight producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                uint32_t off = 0;
                for (int s = 0; s < 18; ++s) {
                    const uint32_t bytes = s < 15 ? 16384u : 24576u;
                    mbar_wait(&bars[B_EMPTY + stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars[B_FULL + stage], bytes);
                    bulk_g2s(sm + A_RING + stage * RING_STAGE, p.wstream + off, bytes, &bars[B_FULL + stage]);
                    off += bytes;
                    if (++stage == RING_N) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            uint32_t ph_xa = 0, ph_vtd = 0, ph_qkr = 0, ph_pr = 0, ph_or = 0;
            const uint32_t xa = sbase + A_XA, vt = sbase + A_VT, qi = sbase + A_QI, ki = sbase + A_KI;
            const uint32_t ring = sbase + A_RING;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(&bars[B_XA], ph_xa); ph_xa ^= 1;
                tc_fence_after();
                // ---- V^T = Wv * LN(x)^T : A = Wv slab (128 v-dims), B = x image (128 tokens)
                for (int m = 0; m < 2; ++m) {
                    for (int ka = 0; ka < 3; ++ka) {
                        mbar_wait(&bars[B_FULL + stage], phase);
                        tc_fence_after();
                        const uint32_t a = ring + stage * RING_STAGE, b = xa + ka * ATOM_A;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_ss(tmem + TC_VT0 + m * 128, umma_desc_sw128(a + ks * 32), umma_desc_sw128(b + ks * 32),
                                    IDESC_128x128, (ka | ks) != 0);
                        umma_commit(&bars[B_EMPTY + stage]);
                        if (++stage == RING_N) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&bars[m == 0 ? B_VTF0 : B_VTF1]);
                }
                mbar_wait(&bars[B_VTD], ph_vtd); ph_vtd ^= 1;
                tc_fence_after();
                for (int pr = 0; pr < 3; ++pr) {
                    // ---- [q k] of head pair pr: A = x image, B = weight slab (128 rows)
                    for (int ka = 0; ka < 3; ++ka) {
                        mbar_wait(&bars[B_FULL + stage], phase);
                        tc_fence_after();
                        const uint32_t a = xa + ka * ATOM_A, b = ring + stage * RING_STAGE;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_ss(tmem + TC_QK, umma_desc_sw128(a + ks * 32), umma_desc_sw128(b + ks * 32), IDESC_128x128,
                                    (ka | ks) != 0);
                        umma_commit(&bars[B_EMPTY + stage]);
                        if (++stage == RING_N) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&bars[B_QKF]);
                    mbar_wait(&bars[B_QKR], ph_qkr); ph_qkr ^= 1;
                    tc_fence_after();
                    for (int j = 0; j < 2; ++j) {
                        const int h = 2 * pr + j;
                        // ---- S = q_h k_h^T over the whole 128-token tile (two windows, block diagonal used)
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            umma_ss(tmem + TC_S, umma_desc_sw128(qi + j * 64 + ks * 32), umma_desc_sw128(ki + j * 64 + ks * 32),
                                    IDESC_128x128, ks != 0);
                        umma_commit(&bars[B_SF]);
                        mbar_wait(&bars[B_PR], ph_pr); ph_pr ^= 1;
                        tc_fence_after();
                        // ---- O_h = P v_h : A = P (TMEM), B = V^T rows of head h (32 x 128 keys)
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                            umma_ts(tmem + TC_O + 32 * h, tmem + TC_P + 8 * kk,
                                    umma_desc_sw128(vt + (kk >> 2) * VT_ATOM + h * 4096 + (kk & 3) * 32), IDESC_128x32, kk != 0);
                        if (h == 5) umma_commit(&bars[B_OF]);
                    }
                }
                mbar_wait(&bars[B_OR], ph_or); ph_or ^= 1;
                tc_fence_after();
                // ---- proj: A = O image, B = Wproj slab (192 rows)
                for (int ka = 0; ka < 3; ++ka) {
                    mbar_wait(&bars[B_FULL + stage], phase);
                    tc_fence_after();
                    const uint32_t a = xa + ka * ATOM_A, b = ring + stage * RING_STAGE;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_ss(tmem + TC_PROJ, umma_desc_sw128(a + ks * 32), umma_desc_sw128(b + ks * 32), IDESC_128x192,
                                (ka | ks) != 0);
                    umma_commit(&bars[B_EMPTY + stage]);
                    if (++stage == RING_N) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars[B_PJF]);
            }
        }
        __syncwarp();
    } else {
        // ===================================================== 128 row threads
        const int cw = warp - 2;                    // 0..3: which 32-row slice this warp loads / stores
        const int q = warp & 3;                     // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;              // accumulator row == token row of the tile
        const uint32_t lanebase = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t xa = sbase + A_XA, vt = sbase + A_VT, qi = sbase + A_QI, ki = sbase + A_KI;
        float* stage_buf = reinterpret_cast<float*>(sm + A_QI);
        const int half = row >> 6, t = row & 63;
        const int rpb_base = (t >> 3) * 15 + (t & 7) + 112;
        uint32_t ph_vt0 = 0, ph_vt1 = 0, ph_qkf = 0, ph_sf = 0, ph_of = 0, ph_pjf = 0;

        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            auto tok_of_row = [&](int r) -> int64_t {
                const int gw = tile * 2 + (r >> 6);
                return gw < p.total_windows ? window_token(p, gw, r & 63) : static_cast<int64_t>(-1);
            };
            // ---- phase 0: gather + LN1 -> x image
            ln_rows_to_image(p.x, p.ld_in, s_vec + SRK_AV_LN_W, s_vec + SRK_AV_LN_B, p.apply_ln, xa, cw, lane, tok_of_row);
            fence_proxy_async_smem();
            mbar_arrive(&bars[B_XA]);

            // mask bits of this row (closed form of calculate_mask, network_swinir.py:216-237)
            const int gw_row = tile * 2 + half;
            uint32_t mh = 0xffu, mw = 0xffu;
            if (p.mask_mode == SRK_MASK_SHIFT) {
                const int w = gw_row % p.nw_img;
                const int wy = w / p.nwx, wx = w - wy * p.nwx;
                auto reg = [&](int pos, int L) { return (pos >= L - 8 ? 1 : 0) + (pos >= L - p.shift ? 1 : 0); };
                const int rh = reg(wy * 8 + (t >> 3), p.H), rw = reg(wx * 8 + (t & 7), p.W);
                mh = 0; mw = 0;
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    mh |= (reg(wy * 8 + a, p.H) == rh ? 1u : 0u) << a;
                    mw |= (reg(wx * 8 + a, p.W) == rw ? 1u : 0u) << a;
                }
            }
            const float* emask = nullptr;
            if (p.mask_mode == SRK_MASK_EXPLICIT && gw_row < p.total_windows)
                emask = p.mask + (static_cast<int64_t>(gw_row % p.mask_nw) * 64 + t) * 64;

            // ---- phase 1: V^T accumulators -> V^T image (thread = v-dim row, columns = tokens)
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (m == 0) { mbar_wait(&bars[B_VTF0], ph_vt0); ph_vt0 ^= 1; }
                else        { mbar_wait(&bars[B_VTF1], ph_vt1); ph_vt1 ^= 1; }
                tc_fence_after();
                if (m == 0 || q < 2) {      // v-dims 192..255 are padding
                    const int vrow = m * 128 + row;
                    const float bv = s_vec[SRK_AV_BIAS_V + vrow];
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t v[32];
                        tmem_ld32(tmem + lanebase + TC_VT0 + m * 128 + 32 * c, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + bv);
                        store_row_chunks(vt + (c >> 1) * VT_ATOM, vrow, (c & 1) * 4, v, nullptr, 1.0f);
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(&bars[B_VTD]);

            float inv_sum[6];
#pragma unroll
            for (int pr = 0; pr < 3; ++pr) {
                // ---- phase 2: q,k accumulators of the head pair -> Q / K images
                mbar_wait(&bars[B_QKF], ph_qkf); ph_qkf ^= 1;
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tmem + lanebase + TC_QK + 32 * c, v);
                    tmem_ld_wait();
                    store_row_chunks(c < 2 ? qi : ki, row, (c & 1) * 4, v, s_vec + SRK_AV_BIAS_QK + pr * 128 + 32 * c, 1.0f);
                }
                tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(&bars[B_QKR]);

#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int h = 2 * pr + j;
                    // ---- phase 3: softmax of this row over the 64 keys of its own window
                    mbar_wait(&bars[B_SF], ph_sf); ph_sf ^= 1;
                    tc_fence_after();
                    uint32_t v0[32], v1[32];
                    tmem_ld32(tmem + lanebase + TC_S + 64 * half, v0);
                    tmem_ld32(tmem + lanebase + TC_S + 64 * half + 32, v1);
                    tmem_ld_wait();
                    const float* rpb = s_vec + SRK_AV_RPB + h * SRK_AV_RPB_STRIDE + rpb_base;
                    float s[64];
                    float mx = -INFINITY;
#pragma unroll
                    for (int jx = 0; jx < 64; ++jx) {
                        float val = __uint_as_float(jx < 32 ? v0[jx] : v1[jx - 32]) + rpb[-(15 * (jx >> 3) + (jx & 7))];
                        if (p.mask_mode == SRK_MASK_SHIFT) {
                            if (!(((mh >> (jx >> 3)) & (mw >> (jx & 7))) & 1u)) val += -100.0f * LOG2E;
                        }
                        s[jx] = val;
                    }
                    if (emask) {
#pragma unroll
                        for (int jx = 0; jx < 64; jx += 4) {
                            const float4 mk = __ldg(reinterpret_cast<const float4*>(emask + jx));
                            s[jx] += mk.x * LOG2E; s[jx + 1] += mk.y * LOG2E; s[jx + 2] += mk.z * LOG2E; s[jx + 3] += mk.w * LOG2E;
                        }
                    }
#pragma unroll
                    for (int jx = 0; jx < 64; ++jx) mx = fmaxf(mx, s[jx]);
                    float sum = 0.f;
                    uint32_t pw[32];
#pragma unroll
                    for (int jx = 0; jx < 64; jx += 2) {
                        const float e0 = exp2f(s[jx] - mx), e1 = exp2f(s[jx + 1] - mx);
                        const __nv_bfloat162 pb = __floats2bfloat162_rn(e0, e1);
                        sum += __low2float(pb) + __high2float(pb);       // normalise by what the MMA will see
                        pw[jx >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
                    }
                    inv_sum[h] = 1.0f / sum;
                    uint32_t zeros[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) zeros[i] = 0u;
                    tmem_st32(tmem + lanebase + TC_P + 32 * half, pw);
                    tmem_st32(tmem + lanebase + TC_P + 32 * (1 - half), zeros);
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&bars[B_PR]);
                }
            }

            // ---- phase 4: O accumulators / row sums -> O image (overwrites the x image)
            mbar_wait(&bars[B_OF], ph_of); ph_of ^= 1;
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 6; ++h) {
                uint32_t v[32];
                tmem_ld32(tmem + lanebase + TC_O + 32 * h, v);
                tmem_ld_wait();
                store_row_chunks(xa + (h >> 1) * ATOM_A, row, (h & 1) * 4, v, nullptr, inv_sum[h]);
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(&bars[B_OR]);

            // ---- phase 5: proj accumulators + bias + shortcut -> y (window reverse + un-shift in the store)
            mbar_wait(&bars[B_PJF], ph_pjf); ph_pjf ^= 1;
            tc_fence_after();
            store_rows_coalesced(tmem + TC_PROJ, lanebase, stage_buf, s_vec + SRK_AV_BIAS_PROJ, p.x, p.ld_in, p.y, p.ld_out,
                                 p.add_residual, row, cw, lane, tok_of_row);
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// K2: MLP half
// ------------------------------------------------------------------------------------------------
constexpr uint32_t M_XA = 0;                           // LN2(x) image [128 x 192]; later the store transposer
constexpr uint32_t M_H = M_XA + 3 * ATOM_A;            // gelu(fc1) image [128 x 384] = 6 k-atoms
constexpr uint32_t M_RING = M_H + 6 * ATOM_A;
constexpr uint32_t M_VEC = M_RING + RING_N * RING_STAGE;
constexpr uint32_t M_BAR = M_VEC + ((SRK_MLP_VEC_FLOATS * 4 + 127) / 128) * 128;
constexpr uint32_t M_END = M_BAR + 256;
constexpr uint32_t K2_SMEM = M_END + 1024;
static_assert(K2_SMEM <= 232448, "K2 shared memory exceeds 227 KB");
constexpr uint32_t TC_FC1 = 0;    // 384 cols (two 192-wide halves)
constexpr uint32_t TC_FC2 = 0;    // 192 cols, reuses fc1's columns once they are drained
enum { MB_FULL = 0, MB_EMPTY = 3, MB_XA = 6, MB_F1A = 7, MB_F1B = 8, MB_HR = 9, MB_F2 = 10, MB_COUNT = 11 };

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__global__ void __launch_bounds__(192, 1) swin_mlp_kernel(const MlpParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - raw);
    float* s_vec = reinterpret_cast<float*>(sm + M_VEC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + M_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + MB_COUNT + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < SRK_MLP_VEC_FLOATS; i += blockDim.x) s_vec[i] = p.vec[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING_N; ++i) { mbar_init(&bars[MB_FULL + i], 1); mbar_init(&bars[MB_EMPTY + i], 1); }
        mbar_init(&bars[MB_XA], 128); mbar_init(&bars[MB_F1A], 1); mbar_init(&bars[MB_F1B], 1);
        mbar_init(&bars[MB_HR], 128); mbar_init(&bars[MB_F2], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int s = 0; s < 12; ++s) {
                    mbar_wait(&bars[MB_EMPTY + stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars[MB_FULL + stage], RING_STAGE);
                    bulk_g2s(sm + M_RING + stage * RING_STAGE, p.wstream + static_cast<size_t>(s) * RING_STAGE, RING_STAGE,
                             &bars[MB_FULL + stage]);
                    if (++stage == RING_N) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, ph_xa = 0, ph_hr = 0;
            const uint32_t xa = sbase + M_XA, hi = sbase + M_H, ring = sbase + M_RING;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(&bars[MB_XA], ph_xa); ph_xa ^= 1;
                tc_fence_after();
                for (int nh = 0; nh < 2; ++nh) {      // fc1, hidden units [192 nh, 192 nh + 192)
                    for (int ka = 0; ka < 3; ++ka) {
                        mbar_wait(&bars[MB_FULL + stage], phase);
                        tc_fence_after();
                        const uint32_t a = xa + ka * ATOM_A, b = ring + stage * RING_STAGE;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_ss(tmem + TC_FC1 + 192 * nh, umma_desc_sw128(a + ks * 32), umma_desc_sw128(b + ks * 32),
                                    IDESC_128x192, (ka | ks) != 0);
                        umma_commit(&bars[MB_EMPTY + stage]);
                        if (++stage == RING_N) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&bars[nh == 0 ? MB_F1A : MB_F1B]);
                }
                mbar_wait(&bars[MB_HR], ph_hr); ph_hr ^= 1;
                tc_fence_after();
                for (int ka = 0; ka < 6; ++ka) {      // fc2 over the 384 (padded) hidden units
                    mbar_wait(&bars[MB_FULL + stage], phase);
                    tc_fence_after();
                    const uint32_t a = hi + ka * ATOM_A, b = ring + stage * RING_STAGE;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_ss(tmem + TC_FC2, umma_desc_sw128(a + ks * 32), umma_desc_sw128(b + ks * 32), IDESC_128x192,
                                (ka | ks) != 0);
                    umma_commit(&bars[MB_EMPTY + stage]);
                    if (++stage == RING_N) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars[MB_F2]);
            }
        }
        __syncwarp();
    } else {
        const int cw = warp - 2, q = warp & 3, row = q * 32 + lane;
        const uint32_t lanebase = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t xa = sbase + M_XA, hi = sbase + M_H;
        float* stage_buf = reinterpret_cast<float*>(sm + M_XA);
        uint32_t ph_f1a = 0, ph_f1b = 0, ph_f2 = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            auto tok_of_row = [&](int r) -> int64_t {
                const int64_t tk = static_cast<int64_t>(tile) * 128 + r;
                return tk < p.num_tokens ? tk : static_cast<int64_t>(-1);
            };
            ln_rows_to_image(p.x, p.ld_in, s_vec + SRK_MV_LN_W, s_vec + SRK_MV_LN_B, p.apply_ln, xa, cw, lane, tok_of_row);
            fence_proxy_async_smem();
            mbar_arrive(&bars[MB_XA]);
#pragma unroll
            for (int nh = 0; nh < 2; ++nh) {
                if (nh == 0) { mbar_wait(&bars[MB_F1A], ph_f1a); ph_f1a ^= 1; }
                else         { mbar_wait(&bars[MB_F1B], ph_f1b); ph_f1b ^= 1; }
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < 6; ++c) {          // hidden units 192 nh + 32 c .. + 32
                    uint32_t v[32];
                    tmem_ld32(tmem + lanebase + TC_FC1 + 192 * nh + 32 * c, v);
                    tmem_ld_wait();
                    const int j0 = 192 * nh + 32 * c;
                    const float* b1 = s_vec + SRK_MV_B1 + j0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(gelu_erf(__uint_as_float(v[i]) + b1[i]));
                    store_row_chunks(hi + (j0 >> 6) * ATOM_A, row, ((j0 & 63) >> 3), v, nullptr, 1.0f);
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(&bars[MB_HR]);

            mbar_wait(&bars[MB_F2], ph_f2); ph_f2 ^= 1;
            tc_fence_after();
            store_rows_coalesced(tmem + TC_FC2, lanebase, stage_buf, s_vec + SRK_MV_B2, p.x, p.ld_in, p.y, p.ld_out,
                                 p.add_residual, row, cw, lane, tok_of_row);
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

cudaError_t launch_swin_attn(const AttnParams& p, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(swin_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
    swin_attn_kernel<<<grid, 192, K1_SMEM, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_swin_mlp(const MlpParams& p, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(swin_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K2_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
    swin_mlp_kernel<<<grid, 192, K2_SMEM, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace srk
