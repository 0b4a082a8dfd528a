// Fused Swin-block kernels for sm_100a (tcgen05 + TMEM + bulk TMA), see DESIGN.md.
//
//   swin_attn_kernel : LN1 -> cyclic shift + window partition (index math in the loads) -> qkv GEMM
//                      -> per head S = q k^T (+ relative position bias, + shifted-window mask),
//                      softmax, P v -> proj GEMM -> + shortcut -> window reverse / un-shift store.
//                      Replaces network_swinir.py:244-276 (and :114-145 in SRK_MODE_WINDOWS).
//   swin_mlp_kernel  : LN2 -> fc1 -> GELU -> fc2 -> + shortcut.  Replaces :277, :24-30.
//
// One CTA per SM, persistent over 128-token tiles (= two 8x8 windows).  swin_attn_kernel: 512 threads, swin_mlp_kernel: 448 (or 320):
//   warp 0 lane 0 : weight producer -- streams the pre-swizzled weight slabs (packing.py: bf16, or fp16 in the variant translation
//                   units swin_kernels_f16.cu / _f16h.cu) from L2 into a 3-stage shared-memory ring with 1-D bulk TMA;
//   warp 1        : tcgen05.mma issuer (warp-uniform issue, all GEMMs accumulate in TMEM); runs ahead of the row threads,
//                   ordered only by mbarriers, so GEMMs of head h+1 overlap the softmax of head h;
//   warps 2..9    : 256 "row" threads in two groups g = 0,1.  Thread <-> TMEM lane <-> token row; the two groups split the
//                   accumulator columns of every epilogue (a TMEM lane quadrant is only reachable from warps with the same
//                   warp_id % 4) and the heads of the softmax (group g: heads g, g + 2, g + 4; one window row = 64 keys per thread);
//   warps 10..13  : swin_attn_kernel: utility warps (q|k epilogues, rows 0-63 of the next tile's LayerNorm);
//                   swin_mlp_kernel: LayerNorm warps one tile ahead;
//   warps 14, 15  : swin_attn_kernel: LayerNorm of rows 64-127 of the next tile.
// Code size matters here: the roles' loop bodies together exceed the instruction cache, so code that runs once per
// tile is kept small (rolled halves, shared out-of-line routines) and the clock64() timeline stamps exist only in the debug build.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "kernels.h"
#define SRK_OOL_TIMEOUT 1      // see umma.cuh: mbar_wait
#define SRK_OOL_WAIT 1         // spin loop of mbar_wait out of line (instruction-cache footprint)
#include "umma.cuh"
#include "rowops.cuh"

// swin_kernels_f16.cu compiles this file a second time with fp16 GEMM operands (umma.cuh: umma_idesc_op / pack_op2): same kernels
// under other names, without the layer kernel and the process-wide debug / tuning globals.
// (the variant translation units rename swin_attn_kernel / swin_mlp_kernel / launch_swin_attn / launch_swin_mlp before including
//  this file; SRK_ONLY_MLP leaves the attention kernel out)

namespace srk {

constexpr int NTHREADS = 448;          // K2: producer + MMA + 8 row warps + 4 LayerNorm warps
constexpr int NROWTHREADS = 256;
constexpr int K1_THREADS = 512;        // K1: + 4 utility warps (q|k epilogues, next-tile rows 0-63) + 2 LayerNorm warps (rows 64-127)
constexpr uint32_t VT_ATOM = 24576;     // 192 rows x 128 B: one k-atom (64 keys) of the V^T image
constexpr uint32_t RING_STAGE = 24576;  // largest weight slab (192 rows x 64 k)
constexpr int RING_N = 3;

constexpr uint32_t IDESC_128x128 = umma_idesc_op(128, 128);
constexpr uint32_t IDESC_128x192 = umma_idesc_op(128, 192);
constexpr uint32_t IDESC_64x64 = umma_idesc_op(64, 64);
constexpr uint32_t IDESC_64x32_BMN = umma_idesc_op(64, 32) | (1u << 16);     // B (= V) MN-major: [key][dim] rows
constexpr float LOG2E = 1.4426950408889634f;

// Optional per-CTA timeline (debug): CTA 0 writes clock64() at event `id` of its `it`-th tile.
#ifdef SRK_NO_TIMELINE
#define SRK_TL(dbgptr, it, id) do { (void)(dbgptr); } while (0)
#else
#define SRK_TL(dbgptr, it, id) do { if ((dbgptr) != nullptr && blockIdx.x == 0 && (it) < 8) (dbgptr)[(it) * 64 + (id)] = clock64(); } while (0)
#endif
#define SRK_TL0(dbgptr, id) do { if (threadIdx.x == 64) SRK_TL(dbgptr, 0, id); } while (0)
#ifndef SRK_F16_OPERANDS
unsigned long long* g_timeline = nullptr;
int g_stagger_attn = 0, g_stagger_mlp = 0, g_stagger_winattn = 1500;
int g_pdl = 1;
#endif

// All CTAs of a launch run the same phase sequence; started together they hit their memory phases (tile load, tile
// store) at the same time and leave HBM / L2 idle in between.  Skewing the start of CTA i by (i mod 4) * `cycles`
// spreads the memory phases of neighbouring SMs over the tile period.
__device__ __forceinline__ void stagger_start(int cycles) {
    const long long wait = static_cast<long long>(blockIdx.x & 3) * cycles;
    if (wait > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < wait) {}
    }
}

// ------------------------------------------------------------------------------------------------
// K1: attention half
// ------------------------------------------------------------------------------------------------
constexpr uint32_t A_XA = 0;                          // normalised x image [128 x 192]
constexpr uint32_t A_VT = A_XA + 3 * ATOM_A;          // V image [128 tokens x 192 dims]; then the O image; then the store staging
constexpr uint32_t A_QKI = A_VT + 2 * VT_ATOM;        // 2 x [q_h | k_h] images [128 x (32+32)] (one per softmax group)
constexpr uint32_t A_RING = A_QKI + 2 * ATOM_A;       // weight ring
// swin_attn_kernel: the store staging (128 rows x 720 B, tile-row order = two windows of 64 tokens) is ONE contiguous region -- the V^T
// and q|k images plus a 12 KB tail right behind them -- so that a whole 8 x 8 window leaves with one tensor-map TMA store (below);
// the weight ring follows the tail.  (The layer kernel keeps the older map: A_RING, rows 28..31 of each quadrant in a separate tail.)
constexpr uint32_t K_TAIL = A_QKI + 2 * ATOM_A;
constexpr uint32_t K_RING = K_TAIL + 12 * 1024;
constexpr uint32_t A_VEC = K_RING + RING_N * RING_STAGE;
constexpr uint32_t A_BAR = A_VEC + ((SRK_ATTN_VEC_FLOATS * 4 + 127) / 128) * 128;
constexpr uint32_t A_END = A_BAR + 256;
constexpr uint32_t K1_SMEM = A_END + 1024;            // + alignment slack
static_assert(K1_SMEM <= 232448, "K1 shared memory exceeds 227 KB");
static_assert(128 * 720 <= 2 * VT_ATOM + 2 * ATOM_A + 12 * 1024 && 3 * ATOM_A <= 2 * VT_ATOM, "O image / store staging must fit");

// TMEM columns (fp32, 128 lanes).  S and P v run as two M = 64 UMMAs per head, one per window: an M = 64 accumulator
// occupies lanes {0-15, 32-47, 64-79, 96-111} and the second window's interleaves at lane offset 16, so both windows of
// the tile share the same 64 columns (no block-diagonal waste, no zero fill).  Softmax/O row of lane L:
// window (L % 32) / 16, token 16 * (L / 32) + L % 16.
constexpr uint32_t TC_O = 0;                          // O accumulator, 6 heads x 32; later the proj accumulator (192)
constexpr uint32_t TC_S0 = 192, TC_S1 = 256;          // S = q k^T (64 keys); P (bf16 pairs) aliases cols 0..31
constexpr uint32_t TC_QK0 = 384, TC_QK1 = 448;        // [q_h | k_h] accumulators (64 cols), double buffered by head parity
constexpr uint32_t TC_V = 192;                        // V accumulator [128 tokens x 192 dims] (before the first S of the tile)
constexpr uint32_t TC_PROJ = 0;
constexpr uint32_t LANE16 = 16u << 16;                // TMEM lane offset of the second window's M = 64 tile

enum {  // K1 barrier slots
    B_FULL = 0, B_EMPTY = 3, B_XA = 6, B_VTF = 7, B_VTD = 8, B_QKF0 = 9, B_QKF1 = 10, B_QKR0 = 11, B_QKR1 = 12, B_SF0 = 13, B_SF1 = 14,
    B_PR0 = 15, B_PR1 = 16, B_OF = 17, B_OR = 18, B_PJF = 19, B_DRAIN = 20, B_XAFREE = 21, B_COUNT = 22
};

// Progress-counter report of a finished tile (umma.cuh), called by the 128 threads of softmax group 0 -- the threads that issued the
// tile's bulk stores.  Out of line: inlined into the row-warp loops it changed their code generation even when never executed.
static __device__ __noinline__ void signal_progress(int* counter, int add) {
    bulk_wait0();                       // this thread's copies have completed (global writes performed)
    named_bar_sync(6, 128);             // ... and those of the other issuing threads
    if (threadIdx.x == 64) {
        __threadfence();
        red_release_gpu_add(counter, add);
    }
}

#ifndef SRK_ONLY_MLP
struct TileGeom {
    int64_t base[2];
    int y0[2], x0[2], img[2];
    bool valid[2];
};
__device__ __forceinline__ void set_tile_geom(const AttnParams& p, int tile, TileGeom& geo) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        const int gw = tile * 2 + hf;
        geo.valid[hf] = gw < p.total_windows;
        if (p.mode == SRK_MODE_WINDOWS) {
            geo.base[hf] = static_cast<int64_t>(gw) * 64; geo.y0[hf] = 0; geo.x0[hf] = 0; geo.img[hf] = 0;
        } else {
            const int b = gw / p.nw_img, w = gw - b * p.nw_img;
            const int wy = w / p.nwx, wx = w - wy * p.nwx;
            geo.img[hf] = b;
            geo.base[hf] = static_cast<int64_t>(b) * p.H * p.W;
            geo.y0[hf] = wy * 8 + p.shift; geo.x0[hf] = wx * 8 + p.shift;
        }
    }
}

// swin_attn_kernel: the same with the two divisions by launch constants done as multiply-high (host-computed magic numbers,
// AttnParams::div_*): set_tile_geom is inlined at four places of the kernel and an integer division is ~40 instructions.
__device__ __forceinline__ int fast_div(int n, uint32_t magic, uint32_t shift) {       // n / d for 0 <= n < 2^31, see div_magic()
    return static_cast<int>((__umulhi(static_cast<uint32_t>(n), magic) + static_cast<uint32_t>(n)) >> shift);
}
__device__ __forceinline__ void set_tile_geom_k1(const AttnParams& p, int tile, TileGeom& geo) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        const int gw = tile * 2 + hf;
        geo.valid[hf] = gw < p.total_windows;
        if (p.mode == SRK_MODE_WINDOWS) {
            geo.base[hf] = static_cast<int64_t>(gw) * 64; geo.y0[hf] = 0; geo.x0[hf] = 0; geo.img[hf] = 0;
        } else {
            const int b = fast_div(gw, p.div_img_m, p.div_img_s), w = gw - b * p.nw_img;
            const int wy = fast_div(w, p.div_nwx_m, p.div_nwx_s), wx = w - wy * p.nwx;
            geo.img[hf] = b;
            geo.base[hf] = static_cast<int64_t>(b) * p.H * p.W;
            geo.y0[hf] = wy * 8 + p.shift; geo.x0[hf] = wx * 8 + p.shift;
        }
    }
}

// token index of tile row r (-1: padding window).  Selects instead of geo.x[hf]: a runtime index would put geo in local memory.
__device__ __forceinline__ int64_t tile_tok(const AttnParams& p, const TileGeom& geo, int r) {
    const bool hi = r >= 64;
    const int tt = r & 63;
    if (!(hi ? geo.valid[1] : geo.valid[0])) return static_cast<int64_t>(-1);
    const int64_t base = hi ? geo.base[1] : geo.base[0];
    if (p.mode == SRK_MODE_WINDOWS) return base + tt;
    int yy = (hi ? geo.y0[1] : geo.y0[0]) + (tt >> 3);
    if (yy >= p.H) yy -= p.H;
    int xx = (hi ? geo.x0[1] : geo.x0[0]) + (tt & 7);
    if (xx >= p.W) xx -= p.W;
    return base + static_cast<int64_t>(yy) * p.W + xx;
}

// Row addresses of a 16-row batch (tile rows row0 .. row0 + 15, row0 a multiple of 16: two window rows of 8 tokens) for
// ln_rows_to_image_p: this lane's row in pass p is token (ty0 + (p >> 2), 2 (p & 3) + (lane >> 4)) of the window, so
// two row bases and four column offsets cover all eight passes -- one add per pass instead of the full index arithmetic.
struct RowSrc16 {
    const float* yrow[2];
    int xo[4];
    __device__ __forceinline__ const float* ptr(int pass) const { return yrow[pass >> 2] ? yrow[pass >> 2] + xo[pass & 3] : nullptr; }
};
__device__ __forceinline__ RowSrc16 make_row_src16(const AttnParams& p, const float* xb, const TileGeom& geo, int row0, int lane) {
    RowSrc16 rs;
    const bool hi = row0 >= 64;
    const int tt0 = row0 & 63, sub = lane >> 4;
    const int64_t base = hi ? geo.base[1] : geo.base[0];
    if (!(hi ? geo.valid[1] : geo.valid[0])) {
        rs.yrow[0] = rs.yrow[1] = nullptr;
#pragma unroll
        for (int k = 0; k < 4; ++k) rs.xo[k] = 0;
    } else if (p.mode == SRK_MODE_WINDOWS) {
        rs.yrow[0] = xb + (base + tt0 + sub) * p.ld_in;
        rs.yrow[1] = rs.yrow[0] + 8 * p.ld_in;
#pragma unroll
        for (int k = 0; k < 4; ++k) rs.xo[k] = 2 * k * p.ld_in;
    } else {
        int ya = (hi ? geo.y0[1] : geo.y0[0]) + (tt0 >> 3);
        if (ya >= p.H) ya -= p.H;
        int yb = ya + 1;
        if (yb >= p.H) yb -= p.H;
        rs.yrow[0] = xb + (base + static_cast<int64_t>(ya) * p.W) * p.ld_in;
        rs.yrow[1] = xb + (base + static_cast<int64_t>(yb) * p.W) * p.ld_in;
        const int x0 = (hi ? geo.x0[1] : geo.x0[0]) + sub;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int xx = x0 + 2 * k;
            if (xx >= p.W) xx -= p.W;
            rs.xo[k] = xx * p.ld_in;
        }
    }
    return rs;
}

// One out-of-line copy of the 16-row LayerNorm -> image routine for the three roles that use it.
static __device__ __noinline__ void k1_ln16_to_image(const float* y0, const float* y1, int xo0, int xo1, int xo2, int xo3, int apply_ln,
                                                     uint32_t xa, int cw8, int lane) {
    RowSrc16 rs;
    rs.yrow[0] = y0; rs.yrow[1] = y1; rs.xo[0] = xo0; rs.xo[1] = xo1; rs.xo[2] = xo2; rs.xo[3] = xo3;
    ln_rows_to_image_p(apply_ln, xa, cw8, lane, [&](int pass) { return rs.ptr(pass); });
}
__device__ __forceinline__ void k1_ln16(const AttnParams& p, const float* xb, const TileGeom& geo, uint32_t xa, int cw8, int lane) {
    const RowSrc16 rs = make_row_src16(p, xb, geo, 16 * cw8, lane);
    k1_ln16_to_image(rs.yrow[0], rs.yrow[1], rs.xo[0], rs.xo[1], rs.xo[2], rs.xo[3], p.apply_ln, xa, cw8, lane);
}

// MMA-issue building blocks of K1, out of line: the kernel is ~200 KB of SASS, its first tile runs from a cold
// instruction cache (2.5x slower than the steady state), and these bodies were inlined at 3 call sites each.
// All are called by the whole MMA warp with uniform arguments (see umma_ss_w).  `cur` = ring stage | phase << 8.
static __device__ __noinline__ uint32_t k1_gemm_k192(uint64_t* bars, uint32_t ring, uint32_t cur, uint32_t d_tmem, uint32_t img,
                                                     uint32_t img_is_a, uint32_t idesc) {      // K = 192: 3 ring slabs x 4 k-steps
    uint32_t stage = cur & 0xffu, phase = cur >> 8;
#pragma unroll 1
    for (int ka = 0; ka < 3; ++ka) {
        mbar_wait(&bars[B_FULL + stage], phase);
        tc_fence_after();
        const uint32_t w = ring + stage * RING_STAGE, im = img + ka * ATOM_A;
        const uint32_t a = img_is_a ? im : w, b = img_is_a ? w : im;
        umma_ss_w4(d_tmem, umma_desc_sw128(a), umma_desc_sw128(b), idesc, ka != 0);
        umma_commit_w(&bars[B_EMPTY + stage]);
        if (++stage == RING_N) { stage = 0; phase ^= 1; }
    }
    return stage | (phase << 8);
}
// 4-D tensor-map TMA store (plain or fp32 reduce-add) of one staged 8 x 8 x 180 window box at (x, y) of image b; whole warp, one lane issues
__device__ __forceinline__ void k1_store_window(const CUtensorMap* tmap, uint32_t src, int x, int y, int b, int add) {
    if (add)
        asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                     "@e cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%1, %2, %3, %4}], [%5];\n\t}" ::"l"(reinterpret_cast<uint64_t>(tmap)),
                     "r"(0), "r"(x), "r"(y), "r"(b), "r"(src) : "memory");
    else
        asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                     "@e cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];\n\t}" ::"l"(reinterpret_cast<uint64_t>(tmap)),
                     "r"(0), "r"(x), "r"(y), "r"(b), "r"(src) : "memory");
}

__global__ void __launch_bounds__(K1_THREADS, 1) swin_attn_kernel(const AttnParams p, const __grid_constant__ CUtensorMap tmap_y) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - raw);
    float* s_vec = reinterpret_cast<float*>(sm + A_VEC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + A_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    SRK_TL0(p.dbg, 13);
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < SRK_ATTN_VEC_FLOATS; i += blockDim.x) s_vec[i] = p.vec[i];      // constants: before the PDL wait
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING_N; ++i) { mbar_init(&bars[B_FULL + i], 1); mbar_init(&bars[B_EMPTY + i], 1); }
        mbar_init(&bars[B_XA], 192); mbar_init(&bars[B_XAFREE], 1);        mbar_init(&bars[B_VTF], 1);    mbar_init(&bars[B_VTD], NROWTHREADS);
        mbar_init(&bars[B_QKF0], 1);         mbar_init(&bars[B_QKF1], 1);   mbar_init(&bars[B_QKR0], 128); mbar_init(&bars[B_QKR1], 128);
        mbar_init(&bars[B_SF0], 1);          mbar_init(&bars[B_SF1], 1);
        mbar_init(&bars[B_PR0], 128);        mbar_init(&bars[B_PR1], 128);  mbar_init(&bars[B_OF], 1);
        mbar_init(&bars[B_OR], NROWTHREADS); mbar_init(&bars[B_PJF], 1);    mbar_init(&bars[B_DRAIN], 128);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    SRK_TL0(p.dbg, 14);
    // order after the previous kernel: the whole grid (griddepcontrol.wait), or -- with progress counters -- tile by tile below
    const bool flag_wait = p.prog_wait != nullptr && p.wait_target > 0;
    if (warp != 0 && !flag_wait) pdl_wait();      // (the weight producer streams constant slabs either way)
    SRK_TL0(p.dbg, 15);
    // whole warp: returns the residual stream's base once the image of `tile` has been finished by the previous kernel
    auto tile_ready = [&](int tile) -> const float* {
        return p.x + progress_wait(flag_wait ? p.prog_wait + fast_div(tile * 2, p.div_img_m, p.div_img_s) : nullptr, p.wait_target);
    };

    if (warp == 0) {
        // ===================================================== weight producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                uint32_t off = 0;
#pragma unroll 1
                for (int s = 0; s < 15; ++s) {      // 3 x 24 KB (Wv k-atoms), 9 x 16 KB (3 head pairs x 3 k-atoms of [q|k|q|k]), 3 x 24 KB (proj)
                    const uint32_t bytes = (s < 3 || s >= 12) ? 24576u : 16384u;
                    mbar_wait(&bars[B_EMPTY + stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars[B_FULL + stage], bytes);
                    bulk_g2s(sm + K_RING + stage * RING_STAGE, p.wstream + off, bytes, &bars[B_FULL + stage]);
                    off += bytes;
                    if (++stage == RING_N) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================== MMA issuer (warp-uniform: see umma_ss_w)
        {
            unsigned long long* mdbg = lane == 0 ? p.dbg : nullptr;
            uint32_t ph_xa = 0, ph_vtd = 0, ph_qkr[2] = {0, 0}, ph_pr[2] = {0, 0}, ph_or = 0;
            const uint32_t xa = sbase + A_XA, vt = sbase + A_VT, qki = sbase + A_QKI, ring = sbase + K_RING;
            uint32_t cur = 0;                   // weight ring cursor (stage | phase << 8)
            auto gemm_k192 = [&](uint32_t d_tmem, uint32_t img, bool img_is_a, uint32_t idesc) {
                cur = k1_gemm_k192(bars, ring, cur, d_tmem, img, img_is_a ? 1u : 0u, idesc);
            };
            auto gemm_qk_pair = [&]() {         // [q_h | k_h | q_h+1 | k_h+1] = xhat * W^T for a pair of heads: one N = 128 GEMM
                gemm_k192(tmem + TC_QK0, xa, true, IDESC_128x128);
                umma_commit_w(&bars[B_QKF0]);
                umma_commit_w(&bars[B_QKF1]);
            };
            auto issue_s = [&](int h) {        // S_w = q_h k_h^T per window w: two M = 64, N = 64 UMMAs sharing 64 columns
                const uint32_t img = qki + (h & 1) * ATOM_A;
                const uint32_t scol = tmem + ((h & 1) ? TC_S1 : TC_S0);
#pragma unroll
                for (int w = 0; w < 2; ++w)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        umma_ss_w(scol + w * LANE16, umma_desc_sw128(img + w * 8192 + ks * 32), umma_desc_sw128(img + w * 8192 + 64 + ks * 32),
                                IDESC_64x64, ks != 0);
                umma_commit_w(&bars[B_SF0 + (h & 1)]);
            };
            auto issue_pv = [&](int h) {       // O_h = P v_h per window: A = P (TMEM, aliases S), B = V rows (keys) of window w, dims of head h (MN-major)
                const uint32_t pcol = tmem + ((h & 1) ? TC_S1 : TC_S0);
#pragma unroll
                for (int w = 0; w < 2; ++w)
                    umma_ts_w4<128>(tmem + TC_O + 32 * h + w * LANE16, pcol + w * LANE16,
                                    umma_desc_sw128(vt + (h >> 1) * ATOM_A + w * 8192 + (h & 1) * 64), IDESC_64x32_BMN, 0);
            };
            int it = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                SRK_TL(mdbg, it, 32);
                mbar_wait(&bars[B_XA], ph_xa); ph_xa ^= 1;
                tc_fence_after();
                SRK_TL(mdbg, it, 33);
                // ---- V = xhat * Wv^T [128 tokens x 192 dims]: A = x image, B = Wv slab (192 rows)
                gemm_k192(tmem + TC_V, xa, true, IDESC_128x192);
                umma_commit_w(&bars[B_VTF]);
                SRK_TL(mdbg, it, 34);
                // ---- heads 0, 1 (their accumulator does not overlap V's); S needs V drained (S0 / S1 alias it)
                gemm_qk_pair();
                mbar_wait(&bars[B_VTD], ph_vtd); ph_vtd ^= 1;
                tc_fence_after();
#pragma unroll 1
                for (int h = 0; h < 6; ++h) {
                    mbar_wait(&bars[B_QKR0 + (h & 1)], ph_qkr[h & 1]); ph_qkr[h & 1] ^= 1;   // image h written, accumulator half h & 1 drained
                    tc_fence_after();
                    SRK_TL(mdbg, it, 35 + h);
                    issue_s(h);
                    SRK_TL(mdbg, it, 44 + h);
                    if ((h & 1) && h < 5) gemm_qk_pair();                  // both halves drained: the next pair of heads
                    if (h == 3) umma_commit_w(&bars[B_XAFREE]);            // last GEMM reading the x image: free once it completes
                    if (h >= 1) {
                        mbar_wait(&bars[B_PR0 + ((h - 1) & 1)], ph_pr[(h - 1) & 1]); ph_pr[(h - 1) & 1] ^= 1;
                        tc_fence_after();
                        issue_pv(h - 1);
                    }
                }
                mbar_wait(&bars[B_PR1], ph_pr[1]); ph_pr[1] ^= 1;
                tc_fence_after();
                issue_pv(5);
                umma_commit_w(&bars[B_OF]);
                SRK_TL(mdbg, it, 41);
                mbar_wait(&bars[B_OR], ph_or); ph_or ^= 1;
                tc_fence_after();
                SRK_TL(mdbg, it, 42);
                // ---- proj: A = O image (in the V^T region), B = Wproj slab (192 rows)
                gemm_k192(tmem + TC_PROJ, vt, true, IDESC_128x192);
                umma_commit_w(&bars[B_PJF]);
                SRK_TL(mdbg, it, 43);
            }
        }
        __syncwarp();
    } else if (warp >= 14) {
        // ===================================================== 2 LayerNorm warps: rows [64, 128) of the next tile's x image.
        // 16 rows per warp are loaded and normalised into registers while the heads of the current tile run, dumped as
        // soon as the last q|k GEMM has read the image, and followed by the other 16.
        const int lw = warp - 14;
        const uint32_t xa = sbase + A_XA;
        uint32_t ph_free = 0;
        TileGeom geo;
        if (static_cast<int>(blockIdx.x) < p.n_tiles) mbar_arrive(&bars[B_XA]);      // first tile: the row warps do all of it
        for (int tile = blockIdx.x; tile + static_cast<int>(gridDim.x) < p.n_tiles; tile += gridDim.x) {
            const float* xb = tile_ready(tile + gridDim.x);
            set_tile_geom_k1(p, tile + gridDim.x, geo);
            // (no register prefetch of the rows while the image is still in use: the image is free half way through the heads, the next
            //  tile needs it ~10 K cycles later, and the prefetch variant was another 470 instructions of once-per-tile code)
            mbar_wait(&bars[B_XAFREE], ph_free); ph_free ^= 1;
            k1_ln16(p, xb, geo, xa, 4 + 2 * lw, lane);
            k1_ln16(p, xb, geo, xa, 4 + 2 * lw + 1, lane);
            fence_proxy_async_smem();
            mbar_arrive(&bars[B_XA]);
        }
    } else if (warp >= 10) {
        // ===================================================== 128 utility threads: q|k epilogues and the next tile's x image
        const int cwu = warp - 10;                  // 0..3: 32-row slice of the LN phase
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lanebase = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t xa = sbase + A_XA, qki = sbase + A_QKI;
        uint32_t ph_qkf[2] = {0, 0};
        TileGeom geo;
        int uit = 0;
        unsigned long long* udbg = threadIdx.x == 320 ? p.dbg : nullptr;
        auto ln_tile = [&](int tile) {               // gather + normalise rows [0, 64) of the next tile -> x image (16 rows per warp)
            const float* xb = tile_ready(tile);
            set_tile_geom_k1(p, tile, geo);
            k1_ln16(p, xb, geo, xa, cwu, lane);
            fence_proxy_async_smem();
            mbar_arrive(&bars[B_XA]);
        };
        // (the first tile is normalised by the 8 row warps, idle at kernel start: twice the loads in flight while HBM is cold)
        uint32_t ph_drain = 0;
        bool first_tile = true;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            if (!first_tile) {      // the previous tile's output rows (staged over the V^T / q|k images) have left shared memory
                mbar_wait(&bars[B_DRAIN], ph_drain); ph_drain ^= 1;
            }
            first_tile = false;
#pragma unroll 1
            for (int h = 0; h < 6; ++h) {
                // ---- q,k accumulators of head h -> [q_h | k_h] image (h & 1).  The k bias is dropped: it shifts every
                //      logit of a row by the same amount, which the softmax cancels.
                mbar_wait(&bars[B_QKF0 + (h & 1)], ph_qkf[h & 1]); ph_qkf[h & 1] ^= 1;
                tc_fence_after();
                SRK_TL(udbg, uit, 50 + h);
                const uint32_t img = qki + (h & 1) * ATOM_A;
                const uint32_t acc = tmem + lanebase + ((h & 1) ? TC_QK1 : TC_QK0);
                uint32_t v[32];
                tmem_ld32(acc, v);
                tmem_ld_wait();
                store_row_chunks<true, false>(img, row, 0, v, s_vec + SRK_AV_BIAS_Q + 32 * h, 1.0f);
                tmem_ld32(acc + 32, v);
                tmem_ld_wait();
                store_row_chunks<false, false>(img, row, 4, v, nullptr, 1.0f);
                tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(&bars[B_QKR0 + (h & 1)]);
                SRK_TL(udbg, uit, 56 + h);
            }
            // ---- the q|k GEMM of head 5 is complete, so nothing reads the x image any more: build the next tile's
            if (tile + static_cast<int>(gridDim.x) < p.n_tiles) ln_tile(tile + gridDim.x);
            SRK_TL(udbg, uit, 62);
            ++uit;
        }
    } else {
        // ===================================================== 256 row threads (two softmax groups)
        const int cw8 = warp - 2;
        const int g = cw8 >> 2;                     // group: softmax of heads h = g (mod 2); column half elsewhere
        const int q = warp & 3;                     // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;              // accumulator row == token row of the tile
        const uint32_t lanebase = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t vt = sbase + A_VT;
        // softmax / O row of this lane (M = 64 accumulator layout): window `half`, token t of that window
        const int half = lane >> 4, t = 16 * q + (lane & 15);
        const int srow = 64 * half + t;             // the same row in tile order (row of the q|k / O images)
        const int rpb_base = (t >> 3) * 15 + (t & 7) + 112;
        uint32_t ph_vtf = 0, ph_sf = 0, ph_of = 0, ph_pjf = 0;
        uint64_t* const bar_sf = &bars[B_SF0 + g];
        uint64_t* const bar_pr = &bars[B_PR0 + g];
        const uint32_t scol = g ? TC_S1 : TC_S0;

        TileGeom geo;
        auto tok_of_row = [&](int r) -> int64_t { return tile_tok(p, geo, r); };
        stagger_start(p.stagger);
        if (static_cast<int>(blockIdx.x) < p.n_tiles) {      // first tile's x image (later ones: the utility warps, one tile ahead)
            const float* xb = tile_ready(blockIdx.x);
            set_tile_geom_k1(p, blockIdx.x, geo);
            k1_ln16(p, xb, geo, sbase + A_XA, cw8, lane);
            fence_proxy_async_smem();
            named_bar_sync(1, NROWTHREADS);
            if (g == 0) mbar_arrive(&bars[B_XA]);
        }

        // Per-tile row state: window geometry and the mask bits of this row (closed form of calculate_mask,
        // network_swinir.py:216-237).  ~300 dependent scalar instructions: computed for the NEXT tile while the proj GEMM
        // runs (the row warps would only wait), not between the store and the next tile's first epilogue.
        struct RowState { TileGeom geo; uint32_t mh, mw; const float* emask; };
        auto prep = [&](int tile, RowState& st) {
            set_tile_geom_k1(p, tile, st.geo);
            const int gw_row = tile * 2 + half;
            st.mh = 0xffu; st.mw = 0xffu;
            if (p.mask_mode == SRK_MASK_SHIFT) {
                const int w = gw_row - fast_div(gw_row, p.div_img_m, p.div_img_s) * p.nw_img;
                const int wy = fast_div(w, p.div_nwx_m, p.div_nwx_s), wx = w - wy * p.nwx;
                auto reg = [&](int pos, int L) { return (pos >= L - 8 ? 1 : 0) + (pos >= L - p.shift ? 1 : 0); };
                const int rh = reg(wy * 8 + (t >> 3), p.H), rw = reg(wx * 8 + (t & 7), p.W);
                st.mh = 0; st.mw = 0;
#pragma unroll 1
                for (int a = 0; a < 8; ++a) {
                    st.mh |= (reg(wy * 8 + a, p.H) == rh ? 1u : 0u) << a;
                    st.mw |= (reg(wx * 8 + a, p.W) == rw ? 1u : 0u) << a;
                }
            }
            st.emask = nullptr;
            if (p.mask_mode == SRK_MASK_EXPLICIT && gw_row < p.total_windows)
                st.emask = p.mask + (static_cast<int64_t>(gw_row % p.mask_nw) * 64 + t) * 64;
        };
        RowState nxt;
        if (static_cast<int>(blockIdx.x) < p.n_tiles) prep(blockIdx.x, nxt);
        // progress counters: group 0 (the threads that issue the bulk stores) reports a tile once its writes have completed --
        // not right after the store but after the next tile's first softmax, when the copies are long done and nothing stalls
        auto signal_tile = [&](int tile) { signal_progress(p.prog_sig + fast_div(tile * 2, p.div_img_m, p.div_img_s), min(2, p.total_windows - 2 * tile)); };

        int it = 0;
        unsigned long long* dbg = threadIdx.x == 64 ? p.dbg : nullptr;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            SRK_TL(dbg, it, 0);
            geo = nxt.geo;
            const uint32_t mh = nxt.mh, mw = nxt.mw;
            const float* emask = nxt.emask;
            const bool masked = (mh & mw) != 0xffu;

            // ---- phase 1: V accumulators -> V image [token][dim] (3 k-atoms of 64 dims; thread = token row, group g = dims 96 g ..).
            //      The v bias is folded into the proj bias at pack time (softmax rows sum to one).
            mbar_wait(&bars[B_VTF], ph_vtf); ph_vtf ^= 1;
            tc_fence_after();
            SRK_TL(dbg, it, 2);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int d0 = 96 * g + 32 * c;
                uint32_t v[32];
                tmem_ld32(tmem + lanebase + TC_V + d0, v);
                tmem_ld_wait();
                store_row_chunks<false, false>(vt + (d0 >> 6) * ATOM_A, row, (d0 & 63) >> 3, v, nullptr, 1.0f);
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(&bars[B_VTD]);
            SRK_TL(dbg, it, 3);

            float inv_sum0 = 0.f, inv_sum1 = 0.f, inv_sum2 = 0.f;      // (scalars: the loops over hh are not unrolled -- code size)
#pragma unroll 1
            for (int hh = 0; hh < 3; ++hh) {
                const int h = 2 * hh + g;
                // ---- softmax of this row over the 64 keys of its own window (exp2 domain; log2 e folded into Wq, rpb)
                mbar_wait(bar_sf, ph_sf); ph_sf ^= 1;
                tc_fence_after();
                SRK_TL(dbg, it, 5 + 3 * hh);
                uint32_t v0[32], v1[32];
                tmem_ld32(tmem + lanebase + scol, v0);
                tmem_ld32(tmem + lanebase + scol + 32, v1);
                tmem_ld_wait();
                const float* rpb = s_vec + SRK_AV_RPB + h * SRK_AV_RPB_STRIDE + rpb_base;
                float s[64];
#pragma unroll
                for (int jx = 0; jx < 64; ++jx)
                    s[jx] = __uint_as_float(jx < 32 ? v0[jx] : v1[jx - 32]) + rpb[-(15 * (jx >> 3) + (jx & 7))];
                if (masked) {
#pragma unroll
                    for (int jx = 0; jx < 64; ++jx)
                        if (!(((mh >> (jx >> 3)) & (mw >> (jx & 7))) & 1u)) s[jx] += -100.0f * LOG2E;
                }
                if (emask) {
#pragma unroll
                    for (int jx = 0; jx < 64; jx += 4) {
                        const float4 mk = __ldg(reinterpret_cast<const float4*>(emask + jx));
                        s[jx] = fmaf(mk.x, LOG2E, s[jx]); s[jx + 1] = fmaf(mk.y, LOG2E, s[jx + 1]);
                        s[jx + 2] = fmaf(mk.z, LOG2E, s[jx + 2]); s[jx + 3] = fmaf(mk.w, LOG2E, s[jx + 3]);
                    }
                }
                float mx = s[0];
#pragma unroll
                for (int jx = 1; jx < 64; ++jx) mx = fmaxf(mx, s[jx]);
                float sum0 = 0.f, sum1 = 0.f;
                uint32_t pw[32];
#pragma unroll
                for (int jx = 0; jx < 64; jx += 2) {
                    const float e0 = ex2_approx(s[jx] - mx), e1 = ex2_approx(s[jx + 1] - mx);
                    sum0 += e0; sum1 += e1;
                    pw[jx >> 1] = pack_op2(e0, e1);
                }
                {
                    const float is = __frcp_rn(sum0 + sum1);
                    if (hh == 0) inv_sum0 = is; else if (hh == 1) inv_sum1 = is; else inv_sum2 = is;
                }
                tmem_st32(tmem + lanebase + scol, pw);                        // P aliases the first half of the S columns
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(bar_pr);
                SRK_TL(dbg, it, 6 + 3 * hh);
                if (p.prog_sig != nullptr && g == 0 && hh == 0 && it > 0) signal_tile(tile - static_cast<int>(gridDim.x));
            }

            // ---- phase 4: O accumulators / row sums -> O image (in the V^T region); this group's own heads
            mbar_wait(&bars[B_OF], ph_of); ph_of ^= 1;
            tc_fence_after();
            SRK_TL(dbg, it, 23);
#pragma unroll 1
            for (int hh = 0; hh < 3; ++hh) {
                const int h = 2 * hh + g;
                uint32_t v[32];
                tmem_ld32(tmem + lanebase + TC_O + 32 * h, v);
                tmem_ld_wait();
                store_row_chunks<false, true>(vt + (h >> 1) * ATOM_A, srow, (h & 1) * 4, v, nullptr, hh == 0 ? inv_sum0 : (hh == 1 ? inv_sum1 : inv_sum2));
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(&bars[B_OR]);
            SRK_TL(dbg, it, 24);
            if (tile + static_cast<int>(gridDim.x) < p.n_tiles) prep(tile + gridDim.x, nxt);

            // ---- phase 5: proj accumulators + bias -> staged rows -> bulk (reduce-add) store; window reverse + un-shift
            //      are the destination addresses of the copies.  Staging = V^T + q|k image regions + a small tail.
            mbar_wait(&bars[B_PJF], ph_pjf); ph_pjf ^= 1;
            tc_fence_after();
            SRK_TL(dbg, it, 26);
            if (p.use_tmap) {
                // SRK_MODE_IMAGE: a window's 64 staged rows are an 8 x 8 x 180 box of the (B, H, W, ld) output: ONE tensor-map store per
                // window (a bulk copy per 8-token window row cost ~600 cycles of issue each, 16+ per tile: 2 - 3 K of a 21 K cycle tile).
                stage_rows_and_bulk_store(tmem + TC_PROJ, lanebase, sm + A_VT, sm + A_VT, 32, s_vec + SRK_AV_BIAS_PROJ, p.y, p.ld_out,
                                          p.add_residual, q, g, lane, tok_of_row, 0, nullptr, nullptr, 0, 0, false);
                named_bar_sync(2 + (q >> 1), 128);          // both lane quadrants (and both column groups) of this window are staged
                if (g == 0) {
                    const int w = q >> 1;
                    const int x0 = w ? geo.x0[1] : geo.x0[0], y0 = w ? geo.y0[1] : geo.y0[0];
                    if (x0 + 8 > p.W || y0 + 8 > p.H) {
                        // a shifted window that wraps around the image edge: the copy engine rejects negative start coordinates for
                        // stores, so its rows leave as before (a bulk copy per contiguous run: 2 x 8 pieces per wrapped window)
                        issue_row_runs(sm + A_VT + (q * 32 + lane) * ROW_BYTES, 32, p.y, p.ld_out, p.add_residual, q, lane, tok_of_row);
                    } else if ((q & 1) == 0 && (w ? geo.valid[1] : geo.valid[0])) {
                        k1_store_window(&tmap_y, sbase + A_VT + static_cast<uint32_t>(w) * 64u * ROW_BYTES, x0, y0, w ? geo.img[1] : geo.img[0],
                                        p.add_residual);
                    }
                    bulk_commit();
                    __syncwarp();
                }
            } else {
                stage_rows_and_bulk_store(tmem + TC_PROJ, lanebase, sm + A_VT, sm + A_VT, 32, s_vec + SRK_AV_BIAS_PROJ, p.y, p.ld_out,
                                          p.add_residual, q, g, lane, tok_of_row, 0, (dbg != nullptr && blockIdx.x == 0 && it < 8) ? dbg + it * 64 : nullptr);
            }
            tc_fence_before();
            SRK_TL(dbg, it, 28);
            // the copies drain while the next tile's V^T / q|k GEMMs run; nobody may write the V^T or q|k images before that
            if (g == 0) {
                bulk_wait_read0();
                SRK_TL(dbg, it, 29);
                mbar_arrive(&bars[B_DRAIN]);        // -> utility warps (q|k images)
            }
            named_bar_sync(1, NROWTHREADS);         // -> both groups (V^T image)
            SRK_TL(dbg, it, 27);
        }
        if (p.prog_sig != nullptr && g == 0 && it > 0) signal_tile(static_cast<int>(blockIdx.x) + (it - 1) * static_cast<int>(gridDim.x));
    }
    SRK_TL0(p.dbg, 16);
    tc_fence_before();
    __syncthreads();
    SRK_TL0(p.dbg, 17);
    if (warp == 1) tmem_dealloc(tmem, 512);
}

#endif  // !SRK_ONLY_MLP

// ------------------------------------------------------------------------------------------------
// K2: MLP half
// ------------------------------------------------------------------------------------------------
constexpr uint32_t M_XA = 0;                           // normalised x image [128 x 192]
constexpr uint32_t M_STAGE = M_XA + 3 * ATOM_A;        // output rows staged for the bulk store (private: copies drain asynchronously)
constexpr uint32_t M_RING = M_STAGE + 6 * ATOM_A;
constexpr uint32_t M_VEC = M_RING + RING_N * RING_STAGE;
constexpr uint32_t M_BAR = M_VEC + ((SRK_MLP_VEC_FLOATS * 4 + 127) / 128) * 128;
constexpr uint32_t M_END = M_BAR + 256;
constexpr uint32_t K2_SMEM = M_END + 1024;
static_assert(K2_SMEM <= 232448, "K2 shared memory exceeds 227 KB");
static_assert(128 * 720 <= 6 * ATOM_A, "store staging size");
// TMEM: fc1 accumulators of one 128-unit hidden chunk, double buffered.  gelu(fc1) is written back over them as bf16 pairs
// (group g: fp32 columns [64 g, 64 g + 64) -> packed columns [64 g, 64 g + 32)) and is the A operand of fc2 straight from TMEM.
constexpr uint32_t TC_F1A = 0, TC_F1B = 128;
constexpr uint32_t TC_F2 = 256;                // fc2 accumulator, 192 cols
enum { MB_FULL = 0, MB_EMPTY = 3, MB_XA = 6, MB_F1A = 7, MB_F1B = 8, MB_HR0 = 9, MB_HR1 = 10, MB_HR2 = 11, MB_F2 = 12, MB_XAFREE = 13, MB_COUNT = 14 };

// LNW = true: 448 threads, 4 dedicated LayerNorm warps run one tile ahead (pays off from ~3 tiles per CTA); LNW = false: 320
// threads, the 8 row warps normalise the next tile while fc2 runs (more registers for the row path, better for 1-2 tiles per CTA).
// Out-of-line MMA-issue blocks of K2 (whole MMA warp, uniform arguments; `cur` = ring stage | phase << 8): see K1.
static __device__ __noinline__ uint32_t k2_fc1_chunk(uint64_t* bars, uint32_t ring, uint32_t cur, uint32_t acc, uint32_t xa, uint64_t* done_bar) {
    uint32_t stage = cur & 0xffu, phase = cur >> 8;
#pragma unroll 1
    for (int ka = 0; ka < 3; ++ka) {
        mbar_wait(&bars[MB_FULL + stage], phase);
        tc_fence_after();
        const uint32_t a = xa + ka * ATOM_A, b = ring + stage * RING_STAGE;
        umma_ss_w4(acc, umma_desc_sw128(a), umma_desc_sw128(b), IDESC_128x128, ka != 0);
        umma_commit_w(&bars[MB_EMPTY + stage]);
        if (++stage == RING_N) { stage = 0; phase ^= 1; }
    }
    umma_commit_w(done_bar);
    return stage | (phase << 8);
}
static __device__ __noinline__ uint32_t k2_fc2_chunk(uint64_t* bars, uint32_t ring, uint32_t cur, uint32_t acc, uint32_t h_tmem, uint32_t first) {
    uint32_t stage = cur & 0xffu, phase = cur >> 8;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        mbar_wait(&bars[MB_FULL + stage], phase);
        tc_fence_after();
        const uint32_t b = ring + stage * RING_STAGE;
        umma_ts_w4(acc, h_tmem + 64 * half, umma_desc_sw128(b), IDESC_128x192, !(first && half == 0));
        umma_commit_w(&bars[MB_EMPTY + stage]);
        if (++stage == RING_N) { stage = 0; phase ^= 1; }
    }
    return stage | (phase << 8);
}

template <bool LNW>
__global__ void __launch_bounds__(LNW ? NTHREADS : 320, 1) swin_mlp_kernel(const MlpParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - raw);
    float* s_vec = reinterpret_cast<float*>(sm + M_VEC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + M_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + MB_COUNT + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    SRK_TL0(p.dbg, 13);
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < SRK_MLP_VEC_FLOATS; i += blockDim.x) s_vec[i] = p.vec[i];       // constants: before the PDL wait
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING_N; ++i) { mbar_init(&bars[MB_FULL + i], 1); mbar_init(&bars[MB_EMPTY + i], 1); }
        mbar_init(&bars[MB_XA], 128);          mbar_init(&bars[MB_F1A], 1);           mbar_init(&bars[MB_F1B], 1);
        mbar_init(&bars[MB_XAFREE], 1);
        mbar_init(&bars[MB_HR0], NROWTHREADS); mbar_init(&bars[MB_HR1], NROWTHREADS); mbar_init(&bars[MB_HR2], NROWTHREADS);
        mbar_init(&bars[MB_F2], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    SRK_TL0(p.dbg, 14);
    const bool flag_wait = p.prog_wait != nullptr && p.wait_target > 0;      // see swin_attn_kernel
    if (warp != 0 && !flag_wait) pdl_wait();      // (the weight producer streams constant slabs either way)
    SRK_TL0(p.dbg, 15);
    auto tile_image = [&](int tile) { return static_cast<int>((static_cast<int64_t>(tile) * 128) / p.tokens_per_image); };
    // whole warp: returns the residual stream's base once the image of `tile` has been finished by the previous kernel
    auto tile_ready = [&](int tile) -> const float* {
        return p.x + progress_wait(flag_wait ? p.prog_wait + tile_image(tile) : nullptr, p.wait_target);
    };

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                uint32_t off = 0;
#pragma unroll 1
                for (int s = 0; s < 15; ++s) {      // MMA order: fc1 c0, c1 (6 x 16 KB) | fc2 k-atoms 0,1 (2 x 24 KB) | fc1 c2 (3 x 16 KB) | fc2 k-atoms 2..5
                    const uint32_t bytes = (s < 6 || (s >= 8 && s < 11)) ? 16384u : 24576u;
                    mbar_wait(&bars[MB_EMPTY + stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars[MB_FULL + stage], bytes);
                    bulk_g2s(sm + M_RING + stage * RING_STAGE, p.wstream + off, bytes, &bars[MB_FULL + stage]);
                    off += bytes;
                    if (++stage == RING_N) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================== MMA issuer (warp-uniform: see umma_ss_w)
        {
            uint32_t cur = 0, ph_xa = 0, ph_hr[3] = {0, 0, 0};     // cur: weight ring cursor (stage | phase << 8)
            uint32_t nchunk = 0;                              // fc1 chunk counter -> TMEM buffer parity
            const uint32_t xa = sbase + M_XA, ring = sbase + M_RING;
            auto fc1_chunk = [&]() {                            // 128 hidden units: A = x image, B = W1 slab (128 rows)
                const uint32_t buf = nchunk & 1;
                cur = k2_fc1_chunk(bars, ring, cur, tmem + (buf ? TC_F1B : TC_F1A), xa, &bars[buf ? MB_F1B : MB_F1A]);
                ++nchunk;
            };
            auto fc2_chunk = [&](uint32_t buf, bool first) {    // K = 128 hidden units of one chunk: A = H (TMEM), B = two W2 slabs
                cur = k2_fc2_chunk(bars, ring, cur, tmem + TC_F2, tmem + (buf ? TC_F1B : TC_F1A), first ? 1u : 0u);
            };
            unsigned long long* mdbg = lane == 0 ? p.dbg : nullptr;
            int mit = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++mit) {
                SRK_TL(mdbg, mit, 32);
                mbar_wait(&bars[MB_XA], ph_xa); ph_xa ^= 1;
                tc_fence_after();
                SRK_TL(mdbg, mit, 33);
                const uint32_t b0 = nchunk & 1;
                fc1_chunk();                                                   // chunk 0 -> buffer b0
                fc1_chunk();                                                   // chunk 1 -> buffer b0 ^ 1
                SRK_TL(mdbg, mit, 34);
                mbar_wait(&bars[MB_HR0], ph_hr[0]); ph_hr[0] ^= 1;             // H of chunk 0 written over its accumulators
                tc_fence_after();
                SRK_TL(mdbg, mit, 35);
                fc2_chunk(b0, true);
                fc1_chunk();                                                   // chunk 2 -> buffer b0 (after fc2 consumed H of chunk 0)
                umma_commit_w(&bars[MB_XAFREE]);                               // all fc1 GEMMs issued: the x image is free once they complete
                SRK_TL(mdbg, mit, 36);
                mbar_wait(&bars[MB_HR1], ph_hr[1]); ph_hr[1] ^= 1;
                tc_fence_after();
                SRK_TL(mdbg, mit, 37);
                fc2_chunk(b0 ^ 1, false);
                SRK_TL(mdbg, mit, 38);
                mbar_wait(&bars[MB_HR2], ph_hr[2]); ph_hr[2] ^= 1;
                tc_fence_after();
                SRK_TL(mdbg, mit, 39);
                fc2_chunk(b0, false);
                umma_commit_w(&bars[MB_F2]);
                SRK_TL(mdbg, mit, 40);
            }
        }
        __syncwarp();
    } else if (LNW && warp >= 10) {
        // ===================================================== 4 LayerNorm warps: one tile ahead of the GEMMs.  The next tile's rows are
        // loaded and normalised into registers while fc1 still reads the x image, then dumped into it (rowops.cuh)
        const int lw = warp - 10;
        const uint32_t xa = sbase + M_XA;
        uint32_t ph_free = 0;
        // the first tile is normalised by the 8 row warps (twice the loads in flight at kernel start); these warps start on the second
        for (int tile = blockIdx.x + gridDim.x; tile < p.n_tiles; tile += gridDim.x) {
            const float* xb = tile_ready(tile);
            auto tok_of_row = [&](int r) -> int64_t {
                const int64_t tk = static_cast<int64_t>(tile) * 128 + r;
                return tk < p.num_tokens ? tk : static_cast<int64_t>(-1);
            };
            uint2 h0[4][3], h1[4][3], h2[4][3], h3[4][3];
            ln_rows_hold<4>(xb, p.ld_in, p.apply_ln, 32 * lw, lane, tok_of_row, h0);
            ln_rows_hold<4>(xb, p.ld_in, p.apply_ln, 32 * lw + 8, lane, tok_of_row, h1);
            ln_rows_hold<4>(xb, p.ld_in, p.apply_ln, 32 * lw + 16, lane, tok_of_row, h2);
            ln_rows_hold<4>(xb, p.ld_in, p.apply_ln, 32 * lw + 24, lane, tok_of_row, h3);
            mbar_wait(&bars[MB_XAFREE], ph_free); ph_free ^= 1;        // the previous tile's fc1 GEMMs have read the x image
            ln_rows_dump<4>(xa, 32 * lw, lane, h0);
            ln_rows_dump<4>(xa, 32 * lw + 8, lane, h1);
            ln_rows_dump<4>(xa, 32 * lw + 16, lane, h2);
            ln_rows_dump<4>(xa, 32 * lw + 24, lane, h3);
            fence_proxy_async_smem();
            mbar_arrive(&bars[MB_XA]);
        }
    } else {
        const int cw8 = warp - 2, g = cw8 >> 2, q = warp & 3;
        const uint32_t lanebase = static_cast<uint32_t>(q * 32) << 16;
        uint32_t ph_f1[2] = {0, 0}, ph_f2 = 0, nchunk = 0;
        stagger_start(p.stagger);
        if (static_cast<int>(blockIdx.x) < p.n_tiles) {         // first tile: all 8 row warps normalise it (see the LayerNorm warps)
            const int tile = blockIdx.x;
            const float* xb = tile_ready(tile);
            auto tok_of_row = [&](int r) -> int64_t {
                const int64_t tk = static_cast<int64_t>(tile) * 128 + r;
                return tk < p.num_tokens ? tk : static_cast<int64_t>(-1);
            };
            ln_rows_to_image(xb, p.ld_in, p.apply_ln, sbase + M_XA, cw8, lane, tok_of_row);
            fence_proxy_async_smem();
            named_bar_sync(1, NROWTHREADS);
            if (g == 0) mbar_arrive(&bars[MB_XA]);
        }
        auto signal_tile = [&](int tile) { signal_progress(p.prog_sig + tile_image(tile), 1); };      // group 0, once the tile's stores completed
        int it = 0;
        unsigned long long* dbg = threadIdx.x == 64 ? p.dbg : nullptr;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            SRK_TL(dbg, it, 0);
            auto tok_of_row = [&](int r) -> int64_t {
                const int64_t tk = static_cast<int64_t>(tile) * 128 + r;
                return tk < p.num_tokens ? tk : static_cast<int64_t>(-1);
            };
            // ---- fc1 accumulators -> +b1 -> GELU -> H image; group g owns hidden units 128 c + 64 g .. + 64 (H atom 2c+g)
#pragma unroll 1
            for (int c = 0; c < 3; ++c) {
                const uint32_t buf = nchunk & 1;
                ++nchunk;
                mbar_wait(&bars[buf ? MB_F1B : MB_F1A], ph_f1[buf]); ph_f1[buf] ^= 1;
                tc_fence_after();
                SRK_TL(dbg, it, 1 + 2 * c);
                {
                    const uint32_t col = (buf ? TC_F1B : TC_F1A) + 64 * g;
                    uint32_t v0[32], v1[32];
                    tmem_ld32(tmem + lanebase + col, v0);
                    tmem_ld32(tmem + lanebase + col + 32, v1);
                    tmem_ld_wait();
                    const float4* b1 = reinterpret_cast<const float4*>(s_vec + SRK_MV_B1 + 128 * c + 64 * g);
                    uint32_t hw[32];                               // 64 hidden units of this row as bf16 pairs
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = b1[i];
                        hw[2 * i] = gelu_pack2(__uint_as_float(v0[4 * i + 0]) + b.x, __uint_as_float(v0[4 * i + 1]) + b.y);
                        hw[2 * i + 1] = gelu_pack2(__uint_as_float(v0[4 * i + 2]) + b.z, __uint_as_float(v0[4 * i + 3]) + b.w);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = b1[8 + i];
                        hw[16 + 2 * i] = gelu_pack2(__uint_as_float(v1[4 * i + 0]) + b.x, __uint_as_float(v1[4 * i + 1]) + b.y);
                        hw[16 + 2 * i + 1] = gelu_pack2(__uint_as_float(v1[4 * i + 2]) + b.z, __uint_as_float(v1[4 * i + 3]) + b.w);
                    }
                    tmem_st32(tmem + lanebase + col, hw);          // H aliases the first half of this thread's own accumulator columns
                    tmem_st_wait();
                }
                tc_fence_before();
                mbar_arrive(&bars[MB_HR0 + c]);
                SRK_TL(dbg, it, 2 + 2 * c);
            }
            SRK_TL(dbg, it, 7);
            if (p.prog_sig != nullptr && g == 0 && it > 0) signal_tile(tile - static_cast<int>(gridDim.x));   // previous tile: its copies are long done
            if (!LNW && tile + static_cast<int>(gridDim.x) < p.n_tiles) {
                // ---- the x image is free (all fc1 GEMMs of this tile are complete): normalise the next tile while fc2 runs
                const int nt = tile + gridDim.x;
                const float* xb = tile_ready(nt);
                auto tok_next = [&](int r) -> int64_t {
                    const int64_t tk = static_cast<int64_t>(nt) * 128 + r;
                    return tk < p.num_tokens ? tk : static_cast<int64_t>(-1);
                };
                ln_rows_to_image(xb, p.ld_in, p.apply_ln, sbase + M_XA, cw8, lane, tok_next);
                fence_proxy_async_smem();
                named_bar_sync(1, NROWTHREADS);
                if (g == 0) mbar_arrive(&bars[MB_XA]);
            }
            // ---- fc2 accumulators + b2 -> staged rows (private staging) -> bulk (reduce-add) store, drained asynchronously
            mbar_wait(&bars[MB_F2], ph_f2); ph_f2 ^= 1;
            tc_fence_after();
            SRK_TL(dbg, it, 8);
            if (g == 0) bulk_wait_read0();          // the previous tile's copies no longer read the private staging rows
            named_bar_sync(2 + q, 64);
            stage_rows_and_bulk_store(tmem + TC_F2, lanebase, sm + M_STAGE, sm + M_STAGE, 32, s_vec + SRK_MV_B2, p.y, p.ld_out,
                                      p.add_residual, q, g, lane, tok_of_row);
            tc_fence_before();
            SRK_TL(dbg, it, 9);
        }
        bulk_wait_read0();          // shared memory must outlive the bulk copies that read it
        if (p.prog_sig != nullptr && g == 0 && it > 0) signal_tile(static_cast<int>(blockIdx.x) + (it - 1) * static_cast<int>(gridDim.x));
    }
    SRK_TL0(p.dbg, 16);
    tc_fence_before();
    __syncthreads();
    SRK_TL0(p.dbg, 17);
    if (warp == 1) tmem_dealloc(tmem, 512);
}

#ifndef SRK_F16_OPERANDS
// ------------------------------------------------------------------------------------------------
// KL: one persistent kernel per BasicLayer (network_swinir.py:349-416): the attention and MLP halves of ALL blocks of the group
// ------------------------------------------------------------------------------------------------
// Work items in global order: for block b: [attention tile 0 .. T-1][MLP tile 0 .. T-1], T = tokens / 128.  CTA c processes items
// c, c + G, c + 2G, ... (G = grid).  An item waits, per tile, for the image it belongs to (the image progress counters of the
// two-kernel form, umma.cuh) -- every wait is on items earlier in the global order, every CTA walks its items in order and all CTAs
// are resident (grid <= #SMs, one CTA per SM), so the smallest unfinished item can always proceed.
// Why: with one launch per half-block, each of the 12 launches of a 6-block group pays a prologue, a cold first tile (LayerNorm of the
// first tile exposed, empty weight ring) and the quantisation of 3.46 tiles per CTA to 4 rounds -- measured: swin_attn_kernel 61 us
// back to back against 37 us of steady-state tile time.  Here the pipelines never drain: the next item's LayerNorm (utility + LayerNorm
// warps), its weights (producer) and its first GEMMs (MMA warp) overlap the current item whatever the two item types are.
// The roles are swin_attn_kernel's (16 warps).  MLP items reuse them: the utility warps (no q|k epilogues to do) and the LayerNorm
// warps build the next item's x image, the row warps run the GELU epilogues and the store.  TMEM of an MLP item is remapped
// (fc1 buffers at columns 192 / 320, fc2 accumulator at 0) so that the first GEMMs of either item type never touch columns [0, 192),
// which the row warps may still be draining for the previous item (proj / fc2 accumulator).
constexpr uint32_t L_VECA = A_RING + RING_N * RING_STAGE;
constexpr uint32_t L_VECM = L_VECA + ((SRK_ATTN_VEC_FLOATS * 4 + 127) / 128) * 128;
constexpr uint32_t L_TAIL = L_VECM + ((SRK_MLP_VEC_FLOATS * 4 + 127) / 128) * 128;
constexpr uint32_t L_BAR = L_TAIL + 16 * 720;
constexpr uint32_t L_END = L_BAR + 512;
constexpr uint32_t KL_SMEM = L_END + 1024;
static_assert(KL_SMEM <= 232448, "KL shared memory exceeds 227 KB");
constexpr uint32_t TL_F1A = 192, TL_F1B = 320, TL_F2 = 0;          // MLP items (see above)
enum { LB_F1A = B_COUNT, LB_F1B, LB_HR0, LB_HR1, LB_HR2, LB_F2, LB_COUNT };

struct LayerItem { int type, blk, tile; };
__device__ __forceinline__ LayerItem layer_item(const LayerParams& p, int i) {
    const int ph = i / p.T;
    LayerItem it;
    it.type = ph & 1; it.blk = ph >> 1; it.tile = i - ph * p.T;
    return it;
}
// per-item view of the geometry helpers' parameter block
__device__ __forceinline__ void layer_attn_params(const LayerParams& p, const LayerItem& it, AttnParams& a) {
    a.H = p.H; a.W = p.W; a.nwx = p.nwx; a.nw_img = p.nw_img; a.ld_in = p.ld; a.ld_out = p.ld; a.apply_ln = 1; a.add_residual = 1;
    a.total_windows = 2 * p.T; a.n_tiles = p.T; a.mask = nullptr; a.mask_nw = 1;
    if (it.type == 0) { a.mode = SRK_MODE_IMAGE; a.shift = p.blk[it.blk].shift; a.mask_mode = a.shift > 0 ? SRK_MASK_SHIFT : SRK_MASK_NONE; }
    else              { a.mode = SRK_MODE_WINDOWS; a.shift = 0; a.mask_mode = SRK_MASK_NONE; }        // rows = tokens 128 tile + r
}
// whole warp: the residual stream's base once everything item `it` reads has been written (block 0's attention: nothing to wait for)
__device__ __forceinline__ const float* layer_item_ready(const LayerParams& p, const LayerItem& it) {
    const int img = it.tile / p.tiles_per_image;
    const int* ctr = it.type == 0 ? p.progress + p.B + img : p.progress + img;
    const int target = it.type == 0 ? it.blk * p.tiles_per_image : (it.blk + 1) * p.nw_img;
    return p.y + progress_wait(target > 0 ? ctr : nullptr, target);
}
static __device__ __noinline__ void layer_load_vec(float* dst, const float* src, int n, int lane_in_group, int group_threads) {
    for (int i = lane_in_group; i < n; i += group_threads) dst[i] = __ldg(src + i);
}
static __device__ __noinline__ uint32_t kl_fc1_chunk(uint64_t* bars, uint32_t ring, uint32_t cur, uint32_t acc, uint32_t xa, uint64_t* done_bar) {
    uint32_t stage = cur & 0xffu, phase = cur >> 8;
#pragma unroll 1
    for (int ka = 0; ka < 3; ++ka) {
        mbar_wait(&bars[B_FULL + stage], phase);
        tc_fence_after();
        umma_ss_w4(acc, umma_desc_sw128(xa + ka * ATOM_A), umma_desc_sw128(ring + stage * RING_STAGE), IDESC_128x128, ka != 0);
        umma_commit_w(&bars[B_EMPTY + stage]);
        if (++stage == RING_N) { stage = 0; phase ^= 1; }
    }
    umma_commit_w(done_bar);
    return stage | (phase << 8);
}
static __device__ __noinline__ uint32_t kl_fc2_chunk(uint64_t* bars, uint32_t ring, uint32_t cur, uint32_t acc, uint32_t h_tmem, uint32_t first) {
    uint32_t stage = cur & 0xffu, phase = cur >> 8;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        mbar_wait(&bars[B_FULL + stage], phase);
        tc_fence_after();
        umma_ts_w4(acc, h_tmem + 64 * half, umma_desc_sw128(ring + stage * RING_STAGE), IDESC_128x192, !(first && half == 0));
        umma_commit_w(&bars[B_EMPTY + stage]);
        if (++stage == RING_N) { stage = 0; phase ^= 1; }
    }
    return stage | (phase << 8);
}

__global__ void __launch_bounds__(K1_THREADS, 1) swin_layer_kernel(const LayerParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - raw);
    float* s_veca = reinterpret_cast<float*>(sm + L_VECA);
    float* s_vecm = reinterpret_cast<float*>(sm + L_VECM);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + LB_COUNT + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, first = blockIdx.x;

    pdl_launch_dependents();
    // constants of the first attention and MLP items (block 0); later blocks: reloaded by the utility warps (see there)
    for (int i = threadIdx.x; i < SRK_ATTN_VEC_FLOATS; i += blockDim.x) s_veca[i] = p.blk[0].attn_vec[i];
    for (int i = threadIdx.x; i < SRK_MLP_VEC_FLOATS; i += blockDim.x) s_vecm[i] = p.blk[0].mlp_vec[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING_N; ++i) { mbar_init(&bars[B_FULL + i], 1); mbar_init(&bars[B_EMPTY + i], 1); }
        mbar_init(&bars[B_XA], 192); mbar_init(&bars[B_XAFREE], 1);        mbar_init(&bars[B_VTF], 1);    mbar_init(&bars[B_VTD], NROWTHREADS);
        mbar_init(&bars[B_QKF0], 1);         mbar_init(&bars[B_QKF1], 1);   mbar_init(&bars[B_QKR0], 128); mbar_init(&bars[B_QKR1], 128);
        mbar_init(&bars[B_SF0], 1);          mbar_init(&bars[B_SF1], 1);
        mbar_init(&bars[B_PR0], 128);        mbar_init(&bars[B_PR1], 128);  mbar_init(&bars[B_OF], 1);
        mbar_init(&bars[B_OR], NROWTHREADS); mbar_init(&bars[B_PJF], 1);    mbar_init(&bars[B_DRAIN], 128);
        mbar_init(&bars[LB_F1A], 1);         mbar_init(&bars[LB_F1B], 1);   mbar_init(&bars[LB_F2], 1);
        mbar_init(&bars[LB_HR0], NROWTHREADS); mbar_init(&bars[LB_HR1], NROWTHREADS); mbar_init(&bars[LB_HR2], NROWTHREADS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp != 0) pdl_wait();          // (the weight producer streams constant slabs either way)

    if (warp == 0) {
        // ===================================================== weight producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int i = first; i < p.n_items; i += G) {
                const LayerItem it = layer_item(p, i);
                const uint8_t* w = it.type == 0 ? p.blk[it.blk].attn_w : p.blk[it.blk].mlp_w;
                uint32_t off = 0;
                for (int s = 0; s < 15; ++s) {      // slab sizes in MMA order: see swin_attn_kernel / swin_mlp_kernel
                    const uint32_t bytes = it.type == 0 ? ((s < 3 || s >= 12) ? 24576u : 16384u) : ((s < 6 || (s >= 8 && s < 11)) ? 16384u : 24576u);
                    mbar_wait(&bars[B_EMPTY + stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars[B_FULL + stage], bytes);
                    bulk_g2s(sm + A_RING + stage * RING_STAGE, w + off, bytes, &bars[B_FULL + stage]);
                    off += bytes;
                    if (++stage == RING_N) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================== MMA issuer (warp-uniform: see umma_ss_w)
        uint32_t ph_xa = 0, ph_vtd = 0, ph_qkr[2] = {0, 0}, ph_pr[2] = {0, 0}, ph_or = 0, ph_hr[3] = {0, 0, 0}, nchunk = 0;
        const uint32_t xa = sbase + A_XA, vt = sbase + A_VT, qki = sbase + A_QKI, ring = sbase + A_RING;
        uint32_t cur = 0;                   // weight ring cursor (stage | phase << 8)
        auto gemm_k192 = [&](uint32_t d_tmem, uint32_t img, bool img_is_a, uint32_t idesc) {
            cur = k1_gemm_k192(bars, ring, cur, d_tmem, img, img_is_a ? 1u : 0u, idesc);
        };
        auto gemm_qk_pair = [&]() {
            gemm_k192(tmem + TC_QK0, xa, true, IDESC_128x128);
            umma_commit_w(&bars[B_QKF0]);
            umma_commit_w(&bars[B_QKF1]);
        };
        auto issue_s = [&](int h) {
            const uint32_t img = qki + (h & 1) * ATOM_A;
            const uint32_t scol = tmem + ((h & 1) ? TC_S1 : TC_S0);
#pragma unroll
            for (int w = 0; w < 2; ++w)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                    umma_ss_w(scol + w * LANE16, umma_desc_sw128(img + w * 8192 + ks * 32), umma_desc_sw128(img + w * 8192 + 64 + ks * 32),
                              IDESC_64x64, ks != 0);
            umma_commit_w(&bars[B_SF0 + (h & 1)]);
        };
        auto issue_pv = [&](int h) {
            const uint32_t pcol = tmem + ((h & 1) ? TC_S1 : TC_S0);
#pragma unroll
            for (int w = 0; w < 2; ++w)
                umma_ts_w4<128>(tmem + TC_O + 32 * h + w * LANE16, pcol + w * LANE16,
                                umma_desc_sw128(vt + (h >> 1) * ATOM_A + w * 8192 + (h & 1) * 64), IDESC_64x32_BMN, 0);
        };
        unsigned long long* mdbg = lane == 0 ? p.dbg : nullptr;
        int mit = 0;
        for (int i = first; i < p.n_items; i += G, ++mit) {
            const LayerItem it = layer_item(p, i);
            SRK_TL(mdbg, mit, 32);
            mbar_wait(&bars[B_XA], ph_xa); ph_xa ^= 1;
            tc_fence_after();
            SRK_TL(mdbg, mit, 33);
            if (it.type == 0) {
                gemm_k192(tmem + TC_V, xa, true, IDESC_128x192);
                umma_commit_w(&bars[B_VTF]);
                gemm_qk_pair();
                mbar_wait(&bars[B_VTD], ph_vtd); ph_vtd ^= 1;
                tc_fence_after();
#pragma unroll 1
                for (int h = 0; h < 6; ++h) {
                    mbar_wait(&bars[B_QKR0 + (h & 1)], ph_qkr[h & 1]); ph_qkr[h & 1] ^= 1;
                    tc_fence_after();
                    issue_s(h);
                    if ((h & 1) && h < 5) gemm_qk_pair();
                    if (h == 3) umma_commit_w(&bars[B_XAFREE]);
                    if (h >= 1) {
                        mbar_wait(&bars[B_PR0 + ((h - 1) & 1)], ph_pr[(h - 1) & 1]); ph_pr[(h - 1) & 1] ^= 1;
                        tc_fence_after();
                        issue_pv(h - 1);
                    }
                }
                mbar_wait(&bars[B_PR1], ph_pr[1]); ph_pr[1] ^= 1;
                tc_fence_after();
                issue_pv(5);
                umma_commit_w(&bars[B_OF]);
                mbar_wait(&bars[B_OR], ph_or); ph_or ^= 1;
                tc_fence_after();
                gemm_k192(tmem + TC_PROJ, vt, true, IDESC_128x192);
                umma_commit_w(&bars[B_PJF]);
            } else {
                const uint32_t b0 = nchunk & 1;
                auto f1buf = [&](uint32_t buf) { return tmem + (buf ? TL_F1B : TL_F1A); };
                auto fc1 = [&]() {
                    const uint32_t buf = nchunk & 1;
                    cur = kl_fc1_chunk(bars, ring, cur, f1buf(buf), xa, &bars[buf ? LB_F1B : LB_F1A]);
                    ++nchunk;
                };
                fc1();                                                           // chunk 0 -> buffer b0
                fc1();                                                           // chunk 1 -> buffer b0 ^ 1
                SRK_TL(mdbg, mit, 34);
                mbar_wait(&bars[LB_HR0], ph_hr[0]); ph_hr[0] ^= 1;               // H of chunk 0 written over its accumulators;
                tc_fence_after();                                                // (the row warps are past the previous item: columns [0,192) are free)
                SRK_TL(mdbg, mit, 35);
                cur = kl_fc2_chunk(bars, ring, cur, tmem + TL_F2, f1buf(b0), 1u);
                fc1();                                                           // chunk 2 -> buffer b0
                umma_commit_w(&bars[B_XAFREE]);                                  // all fc1 GEMMs issued: the x image is free once they complete
                SRK_TL(mdbg, mit, 36);
                mbar_wait(&bars[LB_HR1], ph_hr[1]); ph_hr[1] ^= 1;
                tc_fence_after();
                SRK_TL(mdbg, mit, 37);
                cur = kl_fc2_chunk(bars, ring, cur, tmem + TL_F2, f1buf(b0 ^ 1), 0u);
                SRK_TL(mdbg, mit, 38);
                mbar_wait(&bars[LB_HR2], ph_hr[2]); ph_hr[2] ^= 1;
                tc_fence_after();
                SRK_TL(mdbg, mit, 39);
                cur = kl_fc2_chunk(bars, ring, cur, tmem + TL_F2, f1buf(b0), 0u);
                umma_commit_w(&bars[LB_F2]);
                SRK_TL(mdbg, mit, 40);
            }
        }
        __syncwarp();
    } else if (warp >= 14) {
        // ===================================================== 2 LayerNorm warps: rows [64, 128) of the NEXT item's x image
        const int lw = warp - 14;
        const uint32_t xa = sbase + A_XA;
        uint32_t ph_free = 0;
        TileGeom geo;
        AttnParams pa{};
        if (first < p.n_items) mbar_arrive(&bars[B_XA]);        // first item: the row warps build all of it
        for (int i = first; i + G < p.n_items; i += G) {
            const LayerItem nx = layer_item(p, i + G);
            layer_attn_params(p, nx, pa);
            const float* xb = layer_item_ready(p, nx);
            set_tile_geom(pa, nx.tile, geo);
            uint2 hb[8][3];
            {
                const RowSrc16 rs = make_row_src16(pa, xb, geo, 64 + 32 * lw, lane);
                ln_rows_hold_p<8>(1, lane, [&](int pass) { return rs.ptr(pass); }, hb);
            }
            mbar_wait(&bars[B_XAFREE], ph_free); ph_free ^= 1;
            ln_rows_dump<8>(xa, 64 + 32 * lw, lane, hb);
            k1_ln16(pa, xb, geo, xa, 4 + 2 * lw + 1, lane);
            fence_proxy_async_smem();
            mbar_arrive(&bars[B_XA]);
        }
    } else if (warp >= 10) {
        // ===================================================== 128 utility threads: q|k epilogues (attention items), rows [0, 64) of the
        //                                                       next item's x image, reload of the per-block constants
        const int cwu = warp - 10;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lanebase = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t xa = sbase + A_XA, qki = sbase + A_QKI;
        uint32_t ph_qkf[2] = {0, 0}, ph_drain = 0;
        TileGeom geo;
        AttnParams pa{};
        int n = 0, blk_a = 0, blk_m = 0;            // blocks whose constants are in s_veca / s_vecm
        for (int i = first; i < p.n_items; i += G, ++n) {
            const LayerItem it = layer_item(p, i);
            const bool has_next = i + G < p.n_items;
            const LayerItem nx = layer_item(p, has_next ? i + G : i);
            if (it.type == 0) {
                // the previous item's output rows (staged over the V^T / q|k images) have left shared memory.  B_DRAIN completes once per
                // item transition (`drain` of the row warps); every phase is consumed in order (attention items: here, MLP items: at
                // their end), so a parity wait can never be two phases ahead of the barrier
                if (n > 0) { mbar_wait(&bars[B_DRAIN], ph_drain); ph_drain ^= 1; }
#pragma unroll 1
                for (int h = 0; h < 6; ++h) {
                    mbar_wait(&bars[B_QKF0 + (h & 1)], ph_qkf[h & 1]); ph_qkf[h & 1] ^= 1;
                    tc_fence_after();
                    const uint32_t img = qki + (h & 1) * ATOM_A;
                    const uint32_t acc = tmem + lanebase + ((h & 1) ? TC_QK1 : TC_QK0);
                    uint32_t v[32];
                    tmem_ld32(acc, v);
                    tmem_ld_wait();
                    store_row_chunks<true, false>(img, row, 0, v, s_veca + SRK_AV_BIAS_Q + 32 * h, 1.0f);
                    tmem_ld32(acc + 32, v);
                    tmem_ld_wait();
                    store_row_chunks<false, false>(img, row, 4, v, nullptr, 1.0f);
                    tc_fence_before();
                    fence_proxy_async_smem();
                    mbar_arrive(&bars[B_QKR0 + (h & 1)]);
                }
                // the q|k GEMM of head 5 is complete, so nothing reads the x image any more
                if (has_next) {
                    layer_attn_params(p, nx, pa);
                    const float* xb = layer_item_ready(p, nx);
                    set_tile_geom(pa, nx.tile, geo);
                    k1_ln16(pa, xb, geo, xa, cwu, lane);
                }
            } else if (has_next) {
                // MLP item: nothing else to do -- load and normalise the next item's rows into registers right away, dump them once the
                // fc1 GEMMs have read the x image (phase n of B_XAFREE: one completion per item)
                layer_attn_params(p, nx, pa);
                const float* xb = layer_item_ready(p, nx);
                set_tile_geom(pa, nx.tile, geo);
                uint2 hb[8][3];
                {
                    const RowSrc16 rs = make_row_src16(pa, xb, geo, 16 * cwu, lane);
                    ln_rows_hold_p<8>(1, lane, [&](int pass) { return rs.ptr(pass); }, hb);
                }
                mbar_wait(&bars[B_XAFREE], static_cast<uint32_t>(n & 1));
                ln_rows_dump<8>(xa, 16 * cwu, lane, hb);
            }
            if (has_next) {
                // constants of the next item's block: its type's previous user is an EARLIER item of this CTA (items of one type are
                // separated by the row warps' sequential order), and its first reader waits behind B_XA, which this arrive releases
                if (nx.type == 0 && nx.blk != blk_a) { layer_load_vec(s_veca, p.blk[nx.blk].attn_vec, SRK_ATTN_VEC_FLOATS, threadIdx.x - 320, 128); blk_a = nx.blk; }
                if (nx.type == 1 && nx.blk != blk_m) { layer_load_vec(s_vecm, p.blk[nx.blk].mlp_vec, SRK_MLP_VEC_FLOATS, threadIdx.x - 320, 128); blk_m = nx.blk; }
                fence_proxy_async_smem();
                mbar_arrive(&bars[B_XA]);
            }
            if (it.type == 1 && n > 0) { mbar_wait(&bars[B_DRAIN], ph_drain); ph_drain ^= 1; }
        }
    } else {
        // ===================================================== 256 row threads (two groups)
        const int cw8 = warp - 2;
        const int g = cw8 >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lanebase = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t vt = sbase + A_VT;
        const int half = lane >> 4, t = 16 * q + (lane & 15);
        const int srow = 64 * half + t;
        const int rpb_base = (t >> 3) * 15 + (t & 7) + 112;
        uint32_t ph_vtf = 0, ph_sf = 0, ph_of = 0, ph_pjf = 0, ph_f1[2] = {0, 0}, ph_f2 = 0, nchunk = 0;
        uint64_t* const bar_sf = &bars[B_SF0 + g];
        uint64_t* const bar_pr = &bars[B_PR0 + g];
        const uint32_t scol = g ? TC_S1 : TC_S0;
        TileGeom geo;
        AttnParams pa{};

        if (first < p.n_items) {            // first item's x image (later ones: the utility / LayerNorm warps, one item ahead)
            const LayerItem it0 = layer_item(p, first);
            layer_attn_params(p, it0, pa);
            const float* xb = layer_item_ready(p, it0);
            set_tile_geom(pa, it0.tile, geo);
            k1_ln16(pa, xb, geo, sbase + A_XA, cw8, lane);
            fence_proxy_async_smem();
            named_bar_sync(1, NROWTHREADS);
            if (g == 0) mbar_arrive(&bars[B_XA]);
        }
        struct RowState { TileGeom geo; uint32_t mh, mw; };
        auto prep = [&](const LayerItem& it, RowState& st) {           // window geometry / mask bits of this row for an attention item
            AttnParams a{};
            layer_attn_params(p, it, a);
            set_tile_geom(a, it.tile, st.geo);
            const int gw_row = it.tile * 2 + half;
            st.mh = 0xffu; st.mw = 0xffu;
            if (a.mask_mode == SRK_MASK_SHIFT) {
                const int w = gw_row % p.nw_img;
                const int wy = w / p.nwx, wx = w - wy * p.nwx;
                auto reg = [&](int pos, int Ln) { return (pos >= Ln - 8 ? 1 : 0) + (pos >= Ln - a.shift ? 1 : 0); };
                const int rh = reg(wy * 8 + (t >> 3), p.H), rw = reg(wx * 8 + (t & 7), p.W);
                st.mh = 0; st.mw = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    st.mh |= (reg(wy * 8 + k, p.H) == rh ? 1u : 0u) << k;
                    st.mw |= (reg(wx * 8 + k, p.W) == rw ? 1u : 0u) << k;
                }
            }
        };
        RowState nxt;
        if (first < p.n_items) prep(layer_item(p, first), nxt);
        // progress report of a finished item by group 0 (the threads that issued its bulk stores): deferred to the middle of the next
        // item (its copies are long done by then, nothing stalls) unless the next item of this CTA is in another phase -- it might
        // depend on this very item -- or there is none
        int* pend_ctr = nullptr;
        int pend_add = 0;
        // The previous item's bulk copies may still read the staging region (V^T + q|k images + tail).  They are waited for lazily, right
        // before the region is written again -- the V epilogue of an attention item, the second GELU chunk of an MLP item (by then the
        // copies, 2.5 - 4 K cycles, are long done; not later: the utility warps count B_DRAIN phases item by item) -- instead of at the
        // end of the item, where the wait was exposed (an MLP item: 3 K of 17.5 K cycles).
        bool drain_pending = false;
        auto drain = [&]() {
            if (drain_pending) {
                if (g == 0) {
                    bulk_wait_read0();
                    mbar_arrive(&bars[B_DRAIN]);        // -> utility warps (q|k images of the next attention item)
                }
                named_bar_sync(1, NROWTHREADS);         // -> both groups
                drain_pending = false;
            }
        };
        auto flush_signal = [&]() {
            if (pend_ctr != nullptr) { signal_progress(pend_ctr, pend_add); pend_ctr = nullptr; }
        };

        int n = 0;
        unsigned long long* dbg = threadIdx.x == 64 ? p.dbg : nullptr;
        for (int i = first; i < p.n_items; i += G, ++n) {
            const LayerItem it = layer_item(p, i);
            const bool has_next = i + G < p.n_items;
            const LayerItem nx = layer_item(p, has_next ? i + G : i);
            SRK_TL(dbg, n, 0);
            if (dbg != nullptr && blockIdx.x == 0 && n == 0) dbg[1024] = clock64();
            if (it.type == 0) {
                // ------------------------------------------------------------------ attention item (swin_attn_kernel's row loop)
                layer_attn_params(p, it, pa);
                geo = nxt.geo;
                const uint32_t mh = nxt.mh, mw = nxt.mw;
                const bool masked = (mh & mw) != 0xffu;
                auto tok_of_row = [&](int r) -> int64_t { return tile_tok(pa, geo, r); };
                mbar_wait(&bars[B_VTF], ph_vtf); ph_vtf ^= 1;
                tc_fence_after();
                drain();                                 // the V image is written over the previous item's staged rows
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int d0 = 96 * g + 32 * c;
                    uint32_t v[32];
                    tmem_ld32(tmem + lanebase + TC_V + d0, v);
                    tmem_ld_wait();
                    store_row_chunks<false, false>(vt + (d0 >> 6) * ATOM_A, row, (d0 & 63) >> 3, v, nullptr, 1.0f);
                }
                tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(&bars[B_VTD]);
                float inv_sum0 = 0.f, inv_sum1 = 0.f, inv_sum2 = 0.f;
#pragma unroll 1
                for (int hh = 0; hh < 3; ++hh) {
                    const int h = 2 * hh + g;
                    mbar_wait(bar_sf, ph_sf); ph_sf ^= 1;
                    tc_fence_after();
                    uint32_t v0[32], v1[32];
                    tmem_ld32(tmem + lanebase + scol, v0);
                    tmem_ld32(tmem + lanebase + scol + 32, v1);
                    tmem_ld_wait();
                    const float* rpb = s_veca + SRK_AV_RPB + h * SRK_AV_RPB_STRIDE + rpb_base;
                    float s[64];
#pragma unroll
                    for (int jx = 0; jx < 64; ++jx)
                        s[jx] = __uint_as_float(jx < 32 ? v0[jx] : v1[jx - 32]) + rpb[-(15 * (jx >> 3) + (jx & 7))];
                    if (masked) {
#pragma unroll
                        for (int jx = 0; jx < 64; ++jx)
                            if (!(((mh >> (jx >> 3)) & (mw >> (jx & 7))) & 1u)) s[jx] += -100.0f * LOG2E;
                    }
                    float mx = s[0];
#pragma unroll
                    for (int jx = 1; jx < 64; ++jx) mx = fmaxf(mx, s[jx]);
                    float sum0 = 0.f, sum1 = 0.f;
                    uint32_t pw[32];
#pragma unroll
                    for (int jx = 0; jx < 64; jx += 2) {
                        const float e0 = ex2_approx(s[jx] - mx), e1 = ex2_approx(s[jx + 1] - mx);
                        sum0 += e0; sum1 += e1;
                        pw[jx >> 1] = pack_op2(e0, e1);
                    }
                    {
                        const float is = __frcp_rn(sum0 + sum1);
                        if (hh == 0) inv_sum0 = is; else if (hh == 1) inv_sum1 = is; else inv_sum2 = is;
                    }
                    tmem_st32(tmem + lanebase + scol, pw);
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(bar_pr);
                    if (g == 0 && hh == 0) flush_signal();
                }
                mbar_wait(&bars[B_OF], ph_of); ph_of ^= 1;
                tc_fence_after();
#pragma unroll 1
                for (int hh = 0; hh < 3; ++hh) {
                    const int h = 2 * hh + g;
                    uint32_t v[32];
                    tmem_ld32(tmem + lanebase + TC_O + 32 * h, v);
                    tmem_ld_wait();
                    store_row_chunks<false, true>(vt + (h >> 1) * ATOM_A, srow, (h & 1) * 4, v, nullptr, hh == 0 ? inv_sum0 : (hh == 1 ? inv_sum1 : inv_sum2));
                }
                tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(&bars[B_OR]);
                if (has_next && nx.type == 0) prep(nx, nxt);          // ~300 dependent scalar instructions, under the proj GEMM
                mbar_wait(&bars[B_PJF], ph_pjf); ph_pjf ^= 1;
                tc_fence_after();
                SRK_TL(dbg, n, 26);
                stage_rows_and_bulk_store(tmem + TC_PROJ, lanebase, sm + A_VT, sm + L_TAIL, 28, s_veca + SRK_AV_BIAS_PROJ, p.y, p.ld,
                                          1, q, g, lane, tok_of_row, 0, (dbg != nullptr && blockIdx.x == 0 && n < 8) ? dbg + n * 64 : nullptr);
                tc_fence_before();
                SRK_TL(dbg, n, 28);
                if (g == 0) { pend_ctr = p.progress + (it.tile * 2) / p.nw_img; pend_add = 2; }
            } else {
                // ------------------------------------------------------------------ MLP item (swin_mlp_kernel's row loop)
                const int tile = it.tile;
                auto tok_of_row = [&](int r) -> int64_t { return static_cast<int64_t>(tile) * 128 + r; };
#pragma unroll 1
                for (int c = 0; c < 3; ++c) {
                    const uint32_t buf = nchunk & 1;
                    ++nchunk;
                    mbar_wait(&bars[buf ? LB_F1B : LB_F1A], ph_f1[buf]); ph_f1[buf] ^= 1;
                    tc_fence_after();
                    SRK_TL(dbg, n, 1 + 2 * c);
                    {
                        const uint32_t col = (buf ? TL_F1B : TL_F1A) + 64 * g;
                        uint32_t v0[32], v1[32];
                        tmem_ld32(tmem + lanebase + col, v0);
                        tmem_ld32(tmem + lanebase + col + 32, v1);
                        tmem_ld_wait();
                        const float4* b1 = reinterpret_cast<const float4*>(s_vecm + SRK_MV_B1 + 128 * c + 64 * g);
                        uint32_t hw[32];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float4 b = b1[k];
                            hw[2 * k] = gelu_pack2(__uint_as_float(v0[4 * k + 0]) + b.x, __uint_as_float(v0[4 * k + 1]) + b.y);
                            hw[2 * k + 1] = gelu_pack2(__uint_as_float(v0[4 * k + 2]) + b.z, __uint_as_float(v0[4 * k + 3]) + b.w);
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float4 b = b1[8 + k];
                            hw[16 + 2 * k] = gelu_pack2(__uint_as_float(v1[4 * k + 0]) + b.x, __uint_as_float(v1[4 * k + 1]) + b.y);
                            hw[16 + 2 * k + 1] = gelu_pack2(__uint_as_float(v1[4 * k + 2]) + b.z, __uint_as_float(v1[4 * k + 3]) + b.w);
                        }
                        tmem_st32(tmem + lanebase + col, hw);
                        tmem_st_wait();
                    }
                    tc_fence_before();
                    mbar_arrive(&bars[LB_HR0 + c]);
                    SRK_TL(dbg, n, 2 + 2 * c);
                    if (c == 1) drain();
                }
                if (g == 0) flush_signal();
                if (has_next && nx.type == 0) prep(nx, nxt);
                mbar_wait(&bars[LB_F2], ph_f2); ph_f2 ^= 1;
                tc_fence_after();
                SRK_TL(dbg, n, 8);
                stage_rows_and_bulk_store(tmem + TL_F2, lanebase, sm + A_VT, sm + L_TAIL, 28, s_vecm + SRK_MV_B2, p.y, p.ld, 1, q, g, lane, tok_of_row);
                tc_fence_before();
                SRK_TL(dbg, n, 9);
                if (g == 0) { pend_ctr = p.progress + p.B + tile / p.tiles_per_image; pend_add = 1; }
            }
            drain_pending = true;                       // (see `drain`)
            // no next item, or the next item of this CTA is in another phase (it may depend on this very item): report now
            if (g == 0 && (!has_next || (i + G) / p.T != i / p.T)) flush_signal();
            SRK_TL(dbg, n, 63);
            if (dbg != nullptr && blockIdx.x == 0 && n < 127) dbg[1025 + n] = clock64();      // (tools/timeline_layer.py allocates 2048 entries)
        }
        if (g == 0) bulk_wait_read0();      // shared memory must outlive the bulk copies that read it
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

#endif  // !SRK_F16_OPERANDS (layer kernel)

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static int num_sms() { return device_num_sms(); }
#ifndef SRK_ONLY_MLP
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
#endif

#ifndef SRK_ONLY_MLP
cudaError_t launch_swin_attn(const AttnParams& p_in, cudaStream_t stream) {
    static bool configured[SRK_MAX_DEVICES] = {};
    if (cudaError_t e = configure_smem_once(configured, swin_attn_kernel, K1_SMEM); e != cudaSuccess) return e;
    AttnParams p = p_in;
    auto div_magic = [](int d, uint32_t& m, uint32_t& sh) {      // n / d == (umulhi(n, m) + n) >> sh for 0 <= n < 2^31 (round-up method)
        sh = 0;
        while ((1ll << sh) < d) ++sh;
        m = static_cast<uint32_t>(((1ull << 32) * ((1ull << sh) - static_cast<uint64_t>(d))) / static_cast<uint64_t>(d) + 1);
    };
    div_magic(p.nw_img, p.div_img_m, p.div_img_s);
    div_magic(p.nwx, p.div_nwx_m, p.div_nwx_s);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    p.use_tmap = 0;
    if (p.mode == SRK_MODE_IMAGE) {
        // the output as a 4-D fp32 tensor (180 x W x H x B, row pitch ld_out), box = one window (a failed encode keeps the bulk copies)
        const int B = p.total_windows / p.nw_img;
        const cuuint64_t ld = static_cast<cuuint64_t>(p.ld_out) * sizeof(float);
        const cuuint64_t gdim[4] = {SRK_DIM, static_cast<cuuint64_t>(p.W), static_cast<cuuint64_t>(p.H), static_cast<cuuint64_t>(B)};
        const cuuint64_t gstr[3] = {ld, ld * p.W, ld * p.W * p.H};
        const cuuint32_t box[4] = {SRK_DIM, 8, 8, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        if (EncodeTiledFn enc = encode_tiled_fn())
            if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.y, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                p.use_tmap = 1;
    }
    const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
    return launch_pdl2(swin_attn_kernel, grid, K1_THREADS, K1_SMEM, stream, p, tmap);
}
#endif

cudaError_t launch_swin_mlp(const MlpParams& p, cudaStream_t stream) {
    static bool configured_a[SRK_MAX_DEVICES] = {}, configured_b[SRK_MAX_DEVICES] = {};
    if (cudaError_t e = configure_smem_once(configured_a, swin_mlp_kernel<true>, K2_SMEM); e != cudaSuccess) return e;
    if (cudaError_t e = configure_smem_once(configured_b, swin_mlp_kernel<false>, K2_SMEM); e != cudaSuccess) return e;
    const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
    if (p.n_tiles > 2 * grid) return launch_pdl(swin_mlp_kernel<true>, grid, NTHREADS, K2_SMEM, stream, p);
    return launch_pdl(swin_mlp_kernel<false>, grid, 320, K2_SMEM, stream, p);
}

#ifndef SRK_F16_OPERANDS
cudaError_t launch_swin_layer(const LayerParams& p, cudaStream_t stream) {
    static bool configured[SRK_MAX_DEVICES] = {};
    if (cudaError_t e = configure_smem_once(configured, swin_layer_kernel, KL_SMEM); e != cudaSuccess) return e;
    const int grid = p.T < num_sms() ? p.T : num_sms();
    return launch_pdl(swin_layer_kernel, grid, K1_THREADS, KL_SMEM, stream, p);
}
#endif

}  // namespace srk
