// Generic window attention for sm_100a (tcgen05 + TMEM + bulk TMA): 256-query windows of any rectangular shape
// against KH x KW key windows, on bf16 q/k/v "planes" produced by token_linear_kernel (linear_kernel.cu).
//
//   HAT W-MSA / SW-MSA   hat_arch.py:166-197 + :281-301   <16,16,16,16>: 256 keys, relative position bias, shift mask
//   HAT OCAB             hat_arch.py:403-432              <16,16,24,24>: 576 keys from the zero-padded 24x24 overlap
//                                                          window (nn.Unfold k=24 s=16 p=4), online softmax in 3 chunks
//   DAT spatial windows  dat_arch.py:193-244              <8,32,8,32> and <32,8,32,8>: dynamic position bias table
//
// Plane layout: a plane holds 64 bf16 channels (= two padded heads) of every token: [token][128 B], the
// eight 16-byte chunks of a token row permuted by chunk ^ ((token + phase) & 7).  Because a window row is a run of
// consecutive tokens, ONE 1-D bulk TMA copy per window row drops it into shared memory as rows of a 128-byte-swizzled
// UMMA operand image -- roll, window partition and the OCAB unfold are only the source addresses of those copies
// (out-of-image OCAB keys are copied from a padding page).  Q and K are K-major operands (head hh at byte 64 hh of the
// row); V is consumed directly as an MN-major B operand (keys = K dimension), so no transpose exists anywhere.
//
// One persistent CTA per SM; work item = (window, head pair).  320 threads:
//   warp 0      : producer -- bulk copies of the Q / K / V window images and the pair's bias tables (double buffered
//                 when two sets fit in shared memory);
//   warp 1      : tcgen05.mma issuer.  Per head and query half g: S_g = Q_g K^T (M 128 x N keys, fp32 in TMEM),
//                 later O_g (+)= P_g V with P read from TMEM (bf16 pairs written over S by the row threads);
//   warps 2..9  : two groups of 128 row threads, group g = queries [128 g, 128 g + 128); thread <-> TMEM lane <-> query.
//                 Sweep 1: logits + position bias (+ mask) -> running max, written back; sweep 2: exp2, P as bf16
//                 pairs in place (the row sum comes out of the P v GEMM: padded dim 30 of every V head is 1).  With several key chunks the O accumulator is rescaled in TMEM (online
//                 softmax).  Finally O / rowsum -> bf16 plane rows (the proj GEMM's A operand) or fp32 rows.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "umma.cuh"

namespace srk {

#define WA_TL(dbgptr, slot, id) do { if ((dbgptr) != nullptr && (slot) < 8) (dbgptr)[(slot) * 64 + (id)] = clock64(); } while (0)

namespace {
constexpr float WA_LOG2E = 1.4426950408889634f;
enum { W_FULL = 0, W_EMPTY = 2, W_SF = 4, W_PR = 6, W_OF = 8, W_FREE = 10, W_COUNT = 12 };
constexpr uint32_t TC_OACC = 192;      // O accumulator columns inside a group's 256-column TMEM region

// MN-major B operand (rows = K index, 128 B per row = 64 N elements), SWIZZLE_128B: same bit layout as the K-major
// descriptor; SBO = 1024 B is the pitch of 8-row K groups, LBO (pitch of 64-element N blocks) is unused for N <= 64.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) { return umma_desc_sw128(smem_addr); }
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(int M, int N) { return umma_idesc_bf16(M, N) | (1u << 16); }

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
}  // namespace


// ---- sweep 1 of the softmax: logits (exp2 domain, already scaled) + position bias (+ mask) -> running max; the biased
//      logits are written back in place.  The TMEM load of piece pc + 1 is in flight while piece pc is processed.
//      MASKED (shifted-window mask, or an explicit mask `emrow`) is a separate instantiation so un-masked windows pay nothing
//      for it -- not even its code: the explicit-mask adds live in the MASKED variant only.
//      The un-masked variant is ROLLED (two 32-key pieces per iteration; the double-buffered register arrays stay static):
//      a quarter of the straight-line code.  ncu on the unrolled kernel: instruction-cache hit rate 82.7 %, 1.1 warp-cycles of
//      no-instruction stall per issued instruction -- this kernel's row threads run ~3 K instructions per head step.
//      Rolling needs a piece to cover whole key rows (32 % KW == 0: not OCAB's KW = 24), so that the bias pointer moves by a constant.
template <int KH, int KW, int SY, int NCH, bool MASKED>
__device__ __forceinline__ float softmax_sweep1(uint32_t tacc, const float* rp, int c, const uint32_t (&rowmask)[KH],
                                                const float* emrow, float mx) {
    constexpr int NP = NCH / 32;
    constexpr bool ROLL = !MASKED && (32 % KW == 0) && (NP % 2 == 0);
    uint32_t va[32], vb[32];
    tmem_ld32(tacc, va);
    if constexpr (ROLL) {
        // piece pc covers columns c NCH + 32 pc + e: key row yj = col / KW, xj = col % KW; the bias pointer moves by a constant per piece
        constexpr int STEP = SY * (32 / KW);                            // 32 / KW key rows per piece
        const float* rq = rp - SY * ((c * NCH) / KW);                  // (c NCH is a multiple of KW for these shapes)
        auto piece = [&](uint32_t (&v)[32], const float* r) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const int yj = e / KW, xj = e - yj * KW;
                const float sv = __uint_as_float(v[e]) + r[-(SY * yj + xj)];
                v[e] = __float_as_uint(sv);
                mx = fmaxf(mx, sv);
            }
        };
#pragma unroll 1
        for (int pp = 0; pp < NP; pp += 2) {
            tmem_ld_wait();
            tmem_ld32(tacc + 32 * (pp + 1), vb);
            piece(va, rq - STEP * pp);
            tmem_st32(tacc + 32 * pp, va);
            tmem_ld_wait();
            if (pp + 2 < NP) tmem_ld32(tacc + 32 * (pp + 2), va);
            piece(vb, rq - STEP * (pp + 1));
            tmem_st32(tacc + 32 * (pp + 1), vb);
        }
        tmem_st_wait();
        return mx;
    } else {
#pragma unroll
        for (int pc = 0; pc < NP; ++pc) {
            uint32_t(&v)[32] = (pc & 1) ? vb : va;
            tmem_ld_wait();
            if (pc + 1 < NP) tmem_ld32(tacc + 32 * (pc + 1), (pc & 1) ? va : vb);
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const int col = c * NCH + 32 * pc + e, yj = col / KW, xj = col - yj * KW;
                float sv = __uint_as_float(v[e]) + rp[-(SY * yj + xj)];
                if (MASKED) {
                    if (!((rowmask[yj] >> xj) & 1u)) sv -= 100.0f * WA_LOG2E;
                }
                v[e] = __float_as_uint(sv);
            }
            if (MASKED && emrow) {
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const float4 mk = __ldg(reinterpret_cast<const float4*>(emrow + c * NCH + 32 * pc + e));
                    v[e] = __float_as_uint(fmaf(mk.x, WA_LOG2E, __uint_as_float(v[e])));
                    v[e + 1] = __float_as_uint(fmaf(mk.y, WA_LOG2E, __uint_as_float(v[e + 1])));
                    v[e + 2] = __float_as_uint(fmaf(mk.z, WA_LOG2E, __uint_as_float(v[e + 2])));
                    v[e + 3] = __float_as_uint(fmaf(mk.w, WA_LOG2E, __uint_as_float(v[e + 3])));
                }
            }
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
            tmem_st32(tacc + 32 * pc, v);
        }
        tmem_st_wait();
        return mx;
    }
}

// ---- sweep 2: p = exp2(s - max) as bf16 pairs written over the S columns already consumed.  The row sum is not
//      accumulated here: padded dim 30 of every V head is 1, so the P v GEMM delivers sum_j p_ij in column 30 of O.
//      Rolled like sweep 1 (two pieces per iteration).
template <int NCH>
__device__ __forceinline__ void softmax_sweep2(uint32_t tacc, float mx) {
    constexpr int NP = NCH / 32;
    static_assert(NP % 2 == 0, "two pieces per iteration");
    uint32_t va[32], vb[32];
    tmem_ld32(tacc, va);
    auto piece = [&](const uint32_t (&v)[32], uint32_t dst) {
        uint32_t pw[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2)
            pw[e >> 1] = pack_bf16x2(ex2_approx(__uint_as_float(v[e]) - mx), ex2_approx(__uint_as_float(v[e + 1]) - mx));
        tmem_st16(dst, pw);
    };
#pragma unroll 1
    for (int pp = 0; pp < NP; pp += 2) {
        tmem_ld_wait();
        tmem_ld32(tacc + 32 * (pp + 1), vb);
        piece(va, tacc + 16 * pp);
        tmem_ld_wait();
        if (pp + 2 < NP) tmem_ld32(tacc + 32 * (pp + 2), va);
        piece(vb, tacc + 16 * (pp + 1));
    }
    tmem_st_wait();
}

template <int QH, int QW, int KH, int KW, int SY, int NCH>
struct WinAttnCfg {
    static constexpr int NQ = QH * QW, NK = KH * KW, NCHUNKS = NK / NCH;
    static constexpr int TABF = ((QH + KH - 1) * SY + 3) & ~3;          // floats per head of the strided bias table
    static constexpr uint32_t Q_OFF = 0, K_OFF = NQ * 128, V_OFF = K_OFF + NK * 128, T_OFF = V_OFF + NK * 128;
    static constexpr uint32_t SET_BYTES = T_OFF + ((2 * TABF * 4 + 1023) & ~1023);
    static constexpr int NSETS = (2 * SET_BYTES + 2048 <= 232448) ? 2 : 1;
    static constexpr uint32_t BAR_OFF = NSETS * SET_BYTES;
    static constexpr uint32_t SMEM = BAR_OFF + 256 + 1024;
    static constexpr uint32_t TX_BYTES = (NQ + 2 * NK) * 128 + 2 * TABF * 4;
    static_assert(NQ == 256, "256 queries per window (two M = 128 tiles)");
    static_assert(NK % NCH == 0 && NCH % 32 == 0 && NCH <= 256 && NCH >= 64, "key chunking");
    static_assert(NCHUNKS == 1 || NCH <= 192, "with several chunks the O accumulator must not alias S");
    static_assert(SMEM <= 232448, "shared memory");
};

template <int QH, int QW, int KH, int KW, int SY, int NCH>
__global__ void __launch_bounds__(320, 1) winattn_kernel(const WinAttnParams p) {
    using C = WinAttnCfg<QH, QW, KH, KW, SY, NCH>;
    constexpr int NK = C::NK, NCHUNKS = C::NCHUNKS, TABF = C::TABF, NSETS = C::NSETS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + W_COUNT + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[W_FULL + i], 1);  mbar_init(&bars[W_EMPTY + i], 257);
            mbar_init(&bars[W_SF + i], 1);    mbar_init(&bars[W_PR + i], 128);
            mbar_init(&bars[W_OF + i], 1);    mbar_init(&bars[W_FREE + i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                         // (kernels.h: launch_pdl) the prologue above touched nothing the previous kernel writes
    const int npairs = (p.n_heads + 1) >> 1;
    const int nw_img = p.nwy * p.nwx;

    if (warp == 0) {
        // ===================================================== producer: window images + bias tables
        uint32_t s = 0, ph = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const int pair = item % npairs, wg = item / npairs;
            const int b = wg / nw_img, w = wg - b * nw_img;
            const int wy = w / p.nwx, wx = w - wy * p.nwx;
            if (lane == 0) {
                mbar_wait(&bars[W_EMPTY + s], ph ^ 1);
                mbar_arrive_expect_tx(&bars[W_FULL + s], C::TX_BYTES);
            }
            __syncwarp();
            uint8_t* set = sm + s * C::SET_BYTES;
            uint64_t* full = &bars[W_FULL + s];
            const int64_t img0 = static_cast<int64_t>(b) * p.H * p.W;
            const uint8_t* qp = p.q_planes + pair * p.plane_stride;
            const uint8_t* kp = p.k_planes + pair * p.plane_stride;
            const uint8_t* vp = p.v_planes + pair * p.plane_stride;
            // query rows (and, for self-attention, the identical key / value rows): cyclic shift = wrapped source rows
            for (int ty = lane; ty < QH; ty += 32) {
                int y = wy * QH + ty + p.shift_y;
                if (y >= p.H) y -= p.H;
                int x0 = wx * QW + p.shift_x;
                if (x0 >= p.W) x0 -= p.W;
                const int run1 = (p.W - x0) < QW ? (p.W - x0) : QW;
                const int64_t t1 = (img0 + static_cast<int64_t>(y) * p.W + x0) * 128;
                const int64_t t2 = (img0 + static_cast<int64_t>(y) * p.W) * 128;
                const uint32_t d1 = ty * QW * 128, d2 = d1 + run1 * 128;
                bulk_g2s(set + C::Q_OFF + d1, qp + t1, run1 * 128, full);
                if (run1 < QW) bulk_g2s(set + C::Q_OFF + d2, qp + t2, (QW - run1) * 128, full);
                if (p.wrap) {
                    bulk_g2s(set + C::K_OFF + d1, kp + t1, run1 * 128, full);
                    bulk_g2s(set + C::V_OFF + d1, vp + t1, run1 * 128, full);
                    if (run1 < QW) {
                        bulk_g2s(set + C::K_OFF + d2, kp + t2, (QW - run1) * 128, full);
                        bulk_g2s(set + C::V_OFF + d2, vp + t2, (QW - run1) * 128, full);
                    }
                }
            }
            if (!p.wrap) {
                // overlapping key window (hat_arch.py:378 nn.Unfold with zero padding): clip against the image, zero-fill the rest
                for (int oy = lane; oy < KH; oy += 32) {
                    const int y = wy * QH + p.koff + oy;
                    const int xa = wx * QW + p.koff;                       // may be negative
                    const uint32_t d0 = oy * KW * 128;
                    int lo = xa < 0 ? -xa : 0;                             // zero columns on the left
                    int hi = xa + KW > p.W ? xa + KW - p.W : 0;            // zero columns on the right
                    if (y < 0 || y >= p.H) { lo = KW; hi = 0; }
                    const int mid = KW - lo - hi;
                    // padded keys: k = 0; v = 0 except the ones column (dim 30 of each head) that carries the softmax row sum --
                    // the pattern page is pre-swizzled per row & 7, so the copy starts at the destination row's phase
                    const uint8_t* vpad = p.zero_page + 4096;
                    if (lo > 0) {
                        bulk_g2s(set + C::K_OFF + d0, p.zero_page, lo * 128, full);
                        bulk_g2s(set + C::V_OFF + d0, vpad, lo * 128, full);
                    }
                    if (mid > 0) {
                        const int64_t t = (img0 + static_cast<int64_t>(y) * p.W + xa + lo) * 128;
                        bulk_g2s(set + C::K_OFF + d0 + lo * 128, kp + t, mid * 128, full);
                        bulk_g2s(set + C::V_OFF + d0 + lo * 128, vp + t, mid * 128, full);
                    }
                    if (hi > 0) {
                        bulk_g2s(set + C::K_OFF + d0 + (lo + mid) * 128, p.zero_page, hi * 128, full);
                        bulk_g2s(set + C::V_OFF + d0 + (lo + mid) * 128, vpad + ((lo + mid) & 7) * 128, hi * 128, full);
                    }
                }
            }
            if (lane == 0) bulk_g2s(set + C::T_OFF, p.tab + static_cast<int64_t>(2 * pair) * TABF, 2 * TABF * 4, full);
            if (++s == NSETS) { s = 0; ph ^= 1; }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================== MMA issuer (the whole warp, uniform: see umma_ss_w)
        {
            constexpr uint32_t IDESC_S = umma_idesc_bf16(128, NCH);
            constexpr uint32_t IDESC_PV = umma_idesc_bf16_bmn(128, 32);
            // Event-driven: each query-half group g walks its own sequence of (item, head, key chunk) steps -- S = Q K^T when its
            // TMEM region is free, P v when its P is written -- and the issuer serves whichever group is ready.  The groups are
            // started half a step apart (see the row threads) so one is in the LDS-bound sweep 1 while the other is in the
            // MUFU-bound sweep 2; a fixed service order would pull them back into lockstep.
            int item_g[2] = {static_cast<int>(blockIdx.x), static_cast<int>(blockIdx.x)};
            int k_g[2] = {0, 0}, hh_g[2] = {0, 0}, c_g[2] = {0, 0}, st_g[2] = {0, 0};
            int k_full = -1, done[2] = {0, 0};
            uint32_t ph_free[2] = {1, 1}, ph_pr[2] = {0, 0};
            const long long t_start = clock64();
            while (item_g[0] < p.n_items || item_g[1] < p.n_items) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (item_g[g] >= p.n_items) continue;
                    const int k = k_g[g];
                    const uint32_t s = k % NSETS;
                    if (k > k_full) {                       // first touch of this item's set: its images must have landed
                        if (!mbar_test_wait_w(&bars[W_FULL + s], (k / NSETS) & 1)) continue;
                        k_full = k;
                        tc_fence_after();
                    }
                    const uint32_t set = sbase + s * C::SET_BYTES;
                    const int pair = item_g[g] % npairs;
                    const int nh = (p.n_heads - 2 * pair) < 2 ? (p.n_heads - 2 * pair) : 2;
                    const int hh = hh_g[g], c = c_g[g];
                    if (st_g[g] == 0) {
                        if (c == 0) {       // the group has drained the previous O accumulator (it aliases S when NCH = 256)
                            if (!mbar_test_wait_w(&bars[W_FREE + g], ph_free[g])) continue;
                            ph_free[g] ^= 1;
                            tc_fence_after();
                        }
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            umma_ss_w(tmem + 256 * g, umma_desc_sw128(set + C::Q_OFF + g * 16384 + 64 * hh + 32 * ks),
                                    umma_desc_sw128(set + C::K_OFF + c * NCH * 128 + 64 * hh + 32 * ks), IDESC_S, ks);
                        umma_commit_w(&bars[W_SF + g]);
                        st_g[g] = 1;
                    } else {
                        if (!mbar_test_wait_w(&bars[W_PR + g], ph_pr[g])) continue;
                        ph_pr[g] ^= 1;
                        tc_fence_after();
#pragma unroll
                        for (int k4 = 0; k4 < NCH / 64; ++k4)       // 4 k-steps (64 keys) per call
                            umma_ts_w4<128>(tmem + 256 * g + TC_OACC, tmem + 256 * g + 32 * k4,
                                            umma_desc_sw128_mn(set + C::V_OFF + (c * NCH + 64 * k4) * 128 + 64 * hh), IDESC_PV, (c | k4) != 0);
                        umma_commit_w(&bars[W_OF + g]);
                        st_g[g] = 0;
                        if (++c_g[g] == NCHUNKS) {
                            c_g[g] = 0;
                            if (++hh_g[g] == nh) {          // this group is done with the item; the second one to finish releases the set
                                hh_g[g] = 0;
                                if (++done[k & 1] == 2) {
                                    done[k & 1] = 0;
                                    umma_commit_w(&bars[W_EMPTY + s]);
                                }
                                ++k_g[g];
                                item_g[g] += gridDim.x;
                            }
                        }
                    }
                }
                if (clock64() - t_start > SRK_WAIT_TIMEOUT_CYCLES) {
                    printf("srk: winattn MMA issuer timeout (block %d)\n", (int)blockIdx.x);
                    __trap();
                }
            }
        }
        __syncwarp();
    } else {
        // ===================================================== 2 x 128 row threads
        const int g = (warp - 2) >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const int qi = 128 * g + row;                       // query index inside the window
        const int yi = qi / QW, xi = qi - yi * QW;
        const uint32_t tacc = tmem + (static_cast<uint32_t>(q * 32) << 16) + 256 * g;
        uint32_t s = 0, ph = 0, ph_sf = 0, ph_of = 0;
        bool first_item = true;
        unsigned long long* dbg = (blockIdx.x == 0 && (threadIdx.x == 64 || threadIdx.x == 192)) ? p.dbg : nullptr;   // first lane of each group
        int slot = g;                                       // timeline slots: 2 * (head step) + group
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const int pair = item % npairs, wg = item / npairs;
            const int b = wg / nw_img, w = wg - b * nw_img;
            const int wy = w / p.nwx, wx = w - wy * p.nwx;
            const int nh = (p.n_heads - 2 * pair) < 2 ? (p.n_heads - 2 * pair) : 2;
            int y = wy * QH + yi + p.shift_y;
            if (y >= p.H) y -= p.H;
            int x = wx * QW + xi + p.shift_x;
            if (x >= p.W) x -= p.W;
            const int64_t tok = (static_cast<int64_t>(b) * p.H + y) * p.W + x;
            // closed form of the shifted-window mask (hat_arch.py:921-940, dat_arch.py:318-361): keys in another region get -100
            uint32_t rowmask[KH];                           // bit xj of rowmask[yj]: key (yj, xj) is in this query's region
            bool masked = false;
            if (p.mask_shift) {
                auto regy = [&](int pos) { return (pos >= p.H - QH ? 1 : 0) + (pos >= p.H - p.shift_y ? 1 : 0); };
                auto regx = [&](int pos) { return (pos >= p.W - QW ? 1 : 0) + (pos >= p.W - p.shift_x ? 1 : 0); };
                const int ry = regy(wy * QH + yi), rx = regx(wx * QW + xi);
                uint32_t mw = 0;
                for (int a = 0; a < KW; ++a) mw |= (regx(wx * QW + a) == rx ? 1u : 0u) << a;
                const uint32_t fw = KW == 32 ? 0xffffffffu : ((1u << KW) - 1u);
                masked = mw != fw;
#pragma unroll
                for (int a = 0; a < KH; ++a) {
                    const bool same = regy(wy * QH + a) == ry;
                    rowmask[a] = same ? mw : 0u;
                    masked = masked || !same;
                }
                masked = __any_sync(0xffffffffu, masked);   // warp-uniform choice of the sweep variant
            } else {
#pragma unroll
                for (int a = 0; a < KH; ++a) rowmask[a] = 0xffffffffu;
            }
            const float* emrow = nullptr;
            if (p.emask) {                                  // explicit mask: handled by the MASKED sweep variant
                emrow = p.emask + (static_cast<int64_t>(wg % p.emask_nw) * C::NQ + qi) * NK;
                masked = true;
            }

            mbar_wait(&bars[W_FULL + s], ph);               // the pair's bias tables have landed
            if (g == 1 && (first_item || NSETS == 1) && p.stagger > 0) {
                // start half a step behind group 0 (with one set both groups restart together at every item)
                const long long t0 = clock64();
                while (clock64() - t0 < p.stagger) {}
            }
            first_item = false;
            const float* tab_s = reinterpret_cast<const float*>(sm + s * C::SET_BYTES + C::T_OFF);
            const int base_i = p.c0 + SY * yi + xi;
#pragma unroll 1
            for (int hh = 0; hh < nh; ++hh) {
                const float* rp = tab_s + hh * TABF + base_i;
                float m_run = -1.0e30f;
                WA_TL(dbg, slot, 0);
#pragma unroll
                for (int c = 0; c < NCHUNKS; ++c) {
                    mbar_wait(&bars[W_SF + g], ph_sf); ph_sf ^= 1;
                    tc_fence_after();
                    WA_TL(dbg, slot, 1 + 4 * c);
                    const float mx = masked ? softmax_sweep1<KH, KW, SY, NCH, true>(tacc, rp, c, rowmask, emrow, m_run)
                                            : softmax_sweep1<KH, KW, SY, NCH, false>(tacc, rp, c, rowmask, emrow, m_run);
                    WA_TL(dbg, slot, 2 + 4 * c);
                    if (c > 0) {
                        // ---- online softmax: rescale the O accumulator (its column 30 is the running row sum); the previous
                        //      P v has completed
                        const float corr = ex2_approx(m_run - mx);
                        mbar_wait(&bars[W_OF + g], ph_of); ph_of ^= 1;
                        tc_fence_after();
                        uint32_t o[32];
                        tmem_ld32(tacc + TC_OACC, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * corr);
                        tmem_st32(tacc + TC_OACC, o);
                    }
                    m_run = mx;
                    softmax_sweep2<NCH>(tacc, mx);
                    tc_fence_before();
                    mbar_arrive(&bars[W_PR + g]);
                    WA_TL(dbg, slot, 3 + 4 * c);
                }
                // ---- O / rowsum -> output
                mbar_wait(&bars[W_OF + g], ph_of); ph_of ^= 1;
                tc_fence_after();
                WA_TL(dbg, slot, 20);
                uint32_t o[32];
                tmem_ld32(tacc + TC_OACC, o);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars[W_FREE + g]);             // S / O columns of this group may be overwritten
                const float inv = __frcp_rn(__uint_as_float(o[SRK_HEAD_DIM]));   // column 30 = sum_j p_ij (ones column of V)
                o[SRK_HEAD_DIM] = 0u;                                            // padded dims leave as zeros
                if (p.out_mode == 0) {
                    uint8_t* dst = p.o_planes + pair * p.o_plane_stride + tok * 128;
                    const uint32_t key = static_cast<uint32_t>(tok) & 7u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t w4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            w4[e] = pack_bf16x2(__uint_as_float(o[8 * k + 2 * e]) * inv, __uint_as_float(o[8 * k + 2 * e + 1]) * inv);
                        st_global_v4(dst + (((4 * hh + k) ^ key) << 4), w4[0], w4[1], w4[2], w4[3]);
                    }
                } else {
                    float* dst = p.o_rows + tok * p.o_ld + p.o_col0 + (2 * pair + hh) * SRK_HEAD_DIM;
#pragma unroll
                    for (int d = 0; d < SRK_HEAD_DIM; d += 2)
                        *reinterpret_cast<float2*>(dst + d) = make_float2(__uint_as_float(o[d]) * inv, __uint_as_float(o[d + 1]) * inv);
                }
                WA_TL(dbg, slot, 21);
                slot += 2;
            }
            mbar_arrive(&bars[W_EMPTY + s]);                // tables of this set no longer read
            if (++s == NSETS) { s = 0; ph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

static int wa_num_sms() { return device_num_sms(); }

template <int QH, int QW, int KH, int KW, int SY, int NCH>
static cudaError_t launch_cfg(const WinAttnParams& p, cudaStream_t stream) {
    using C = WinAttnCfg<QH, QW, KH, KW, SY, NCH>;
    static bool configured[SRK_MAX_DEVICES] = {};
    auto kern = winattn_kernel<QH, QW, KH, KW, SY, NCH>;
    if (cudaError_t e = configure_smem_once(configured, kern, C::SMEM); e != cudaSuccess) return e;
    const int grid = p.n_items < wa_num_sms() ? p.n_items : wa_num_sms();
    return launch_pdl(kern, grid, 320, C::SMEM, stream, p);
}

int winattn_table_floats(int kind) {
    switch (kind) {
        case SRK_WA_HAT_WMSA: return WinAttnCfg<16, 16, 16, 16, 48, 256>::TABF;
        case SRK_WA_HAT_OCAB: return WinAttnCfg<16, 16, 24, 24, 48, 192>::TABF;
        case SRK_WA_DAT_8x32: return WinAttnCfg<8, 32, 8, 32, 64, 256>::TABF;
        case SRK_WA_DAT_32x8: return WinAttnCfg<32, 8, 32, 8, 24, 256>::TABF;
    }
    return -1;
}

cudaError_t launch_winattn(int kind, const WinAttnParams& p, cudaStream_t stream) {
    switch (kind) {
        case SRK_WA_HAT_WMSA: return launch_cfg<16, 16, 16, 16, 48, 256>(p, stream);
        case SRK_WA_HAT_OCAB: return launch_cfg<16, 16, 24, 24, 48, 192>(p, stream);
        case SRK_WA_DAT_8x32: return launch_cfg<8, 32, 8, 32, 64, 256>(p, stream);
        case SRK_WA_DAT_32x8: return launch_cfg<32, 8, 32, 8, 24, 256>(p, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace srk
