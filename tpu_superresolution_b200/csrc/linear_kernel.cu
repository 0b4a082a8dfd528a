// Token-wise linear layers on tcgen05 for the HAT / DAT paths (sm_100a): y[tok, :] = act(A[tok, :] W^T + b).
//
//   A operand   SRK_LIN_A_ROWS  : fp32 token rows [tok][ld] (the residual stream) with optional LayerNorm -- the
//                                 (x - mean) * rstd image is built by the row threads, the affine is folded into W, b;
//               SRK_LIN_A_PLANES: bf16 planes [k-atom][tok][128 B] (chunk ^ (tok & 7) swizzle) written by a previous
//                                 kernel: one 16 KB bulk TMA copy per k-atom drops 128 tokens into the operand image.
//   output      SRK_LIN_OUT_PLANES: bf16 planes of 64 columns each (q / k / v head pairs, hidden units), swizzle phase
//                                 per plane (so a later kernel can bulk-copy window rows as UMMA operand rows);
//               SRK_LIN_OUT_ROWS  : fp32 rows [tok][ld_out], chunk c = columns [180 c, 180 c + 180), plain or added into y
//                                 with cp.reduce.async.bulk (residual update in place).
// Replaces nn.Linear call sites hat_arch.py:179 (qkv), :195 (proj), :401, :436 (OCAB qkv / proj) and dat_arch.py:371,
// :435, :483, :526, :79-88 (qkv / proj / fc1 / fc2).
//
// One persistent CTA per SM over 128-token tiles, N processed in chunks of 192 columns with double-buffered TMEM
// accumulators (the epilogue of chunk c overlaps the MMAs of chunk c + 1).  448 threads: warp 0 producer (weight
// slabs through a 3-stage ring, A planes), warp 1 MMA issuer, warps 2..9 = 256 row threads (thread <-> TMEM lane <->
// token row; the two groups split the columns of every chunk), warps 10..13 = LayerNorm warps that load and normalise the
// NEXT tile's rows into registers while this tile's GEMMs still read the operand image, then dump them into it.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "rowops.cuh"
#include "umma.cuh"

namespace srk {

namespace {
constexpr int LIN_THREADS = 512;          // producer, MMA issuer, 8 row warps, 6 LayerNorm warps (16 warps: still 128 registers per thread)
constexpr int LIN_RING_N = 3;
constexpr uint32_t LIN_SLAB = 24576;             // 192 rows x 64 k
constexpr uint32_t LIN_NC = 192;                 // columns per chunk
constexpr uint32_t LIN_IDESC = umma_idesc_bf16(128, 192);
constexpr uint32_t L_A = 0;                      // A image: k_atoms k-atoms of 16 KB (48 KB or 96 KB); the rest follows at run time:
// ring (72 KB) | bias vector (6 KB) | barriers | [k_atoms = 3 and OUT_ROWS: 90 KB of fp32 row staging].  With k_atoms = 6 the
// staging rows alias the A image, which is dead once the single chunk's GEMM has completed.
constexpr uint32_t LIN_VEC_BYTES = 1536 * 4;
__host__ __device__ constexpr uint32_t lin_off_ring(int k_atoms) { return k_atoms * ATOM_A; }
__host__ __device__ constexpr uint32_t lin_off_vec(int k_atoms) { return lin_off_ring(k_atoms) + LIN_RING_N * LIN_SLAB; }
__host__ __device__ constexpr uint32_t lin_off_bar(int k_atoms) { return lin_off_vec(k_atoms) + LIN_VEC_BYTES; }
__host__ __device__ constexpr uint32_t lin_off_stage(int k_atoms) { return lin_off_bar(k_atoms) + 256; }
// OUT_PLANES: the three 64-column planes of a chunk are staged as [plane][128 rows][128 B] -- exactly the global layout of 128
// consecutive tokens -- and leave with one 16 KB bulk copy per plane; double buffered when the A image is 48 KB.
constexpr uint32_t LIN_PSTAGE = 3 * ATOM_A;
__host__ __device__ constexpr int lin_n_pstage(int k_atoms) { return k_atoms == 3 ? 2 : 1; }
__host__ __device__ constexpr uint32_t lin_smem(int k_atoms, bool separate_stage, bool planes_out = false) {
    return lin_off_stage(k_atoms) + (separate_stage ? 128 * 720 : 0) + (planes_out ? lin_n_pstage(k_atoms) * LIN_PSTAGE : 0) + 1024;
}
static_assert(lin_smem(3, true) <= 232448 && lin_smem(6, false) <= 232448 && lin_smem(3, false, true) <= 232448 &&
              lin_smem(6, false, true) <= 232448, "token_linear shared memory");
enum { LB_FULL = 0, LB_EMPTY = 3, LB_AFULL = 6, LB_AEMPTY = 7, LB_ACCF = 8, LB_ACCE = 10, LB_DRAIN = 12, LB_COUNT = 13 };

}  // namespace

__global__ void __launch_bounds__(LIN_THREADS, 1) token_linear_kernel(const __grid_constant__ LinearParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - raw);
    const uint32_t L_RING = lin_off_ring(p.k_atoms);
    float* s_vec = reinterpret_cast<float*>(sm + lin_off_vec(p.k_atoms));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + lin_off_bar(p.k_atoms));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + LB_COUNT + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool a_rows = p.a_mode == SRK_LIN_A_ROWS;
    const bool out_rows = p.out_mode == SRK_LIN_OUT_ROWS;
    const bool stage_alias = out_rows && p.k_atoms == 6;           // staging rows live in the A image (dead after the GEMM)
    uint8_t* const stage = stage_alias ? sm + L_A : sm + lin_off_stage(p.k_atoms);

    pdl_launch_dependents();            // (see kernels.h: launch_pdl) the prologue below touches only constants
    for (int i = threadIdx.x; i < p.n_chunks * (int)LIN_NC; i += blockDim.x) s_vec[i] = p.bias[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < LIN_RING_N; ++i) { mbar_init(&bars[LB_FULL + i], 1); mbar_init(&bars[LB_EMPTY + i], 1); }
        mbar_init(&bars[LB_AFULL], a_rows ? 192 : 1);        // row input: 6 LayerNorm warps (first tile: 4 row warps + LayerNorm warps 4, 5)
        mbar_init(&bars[LB_AEMPTY], 1);
        mbar_init(&bars[LB_ACCF], 1);     mbar_init(&bars[LB_ACCF + 1], 1);
        mbar_init(&bars[LB_ACCE], 256);   mbar_init(&bars[LB_ACCE + 1], 256);
        mbar_init(&bars[LB_DRAIN], 128);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                         // everything the previous kernel wrote is visible from here on

    if (warp == 0) {
        // ===================================================== producer
        if (lane == 0) {
            uint32_t stage_i = 0, phase = 0, ph_ae = 1, ph_dr = 1;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                if (!a_rows) {
                    mbar_wait(&bars[LB_AEMPTY], ph_ae); ph_ae ^= 1;           // previous tile's MMAs have read the A image
                    if (stage_alias) { mbar_wait(&bars[LB_DRAIN], ph_dr); ph_dr ^= 1; }   // ... and its output rows have left it
                    const int64_t tok0 = static_cast<int64_t>(tile) * 128;
                    const int64_t left = p.num_tokens - tok0;
                    const uint32_t rows = left < 128 ? static_cast<uint32_t>(left) : 128u;
                    mbar_arrive_expect_tx(&bars[LB_AFULL], rows * 128u * p.k_atoms);
                    for (int ka = 0; ka < p.k_atoms; ++ka)
                        bulk_g2s(sm + L_A + ka * ATOM_A, p.a_planes + ka * p.a_plane_stride + tok0 * 128, rows * 128u, &bars[LB_AFULL]);
                }
                const uint8_t* w = p.wstream;
                for (int s = 0; s < p.n_chunks * p.k_atoms; ++s) {
                    mbar_wait(&bars[LB_EMPTY + stage_i], phase ^ 1);
                    mbar_arrive_expect_tx(&bars[LB_FULL + stage_i], LIN_SLAB);
                    bulk_g2s(sm + L_RING + stage_i * LIN_SLAB, w, LIN_SLAB, &bars[LB_FULL + stage_i]);
                    w += LIN_SLAB;
                    if (++stage_i == LIN_RING_N) { stage_i = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================== MMA issuer (the whole warp, uniform: see umma_ss_w)
        {
            uint32_t stage_i = 0, phase = 0, ph_af = 0, ph_acce[2] = {1, 1}, nacc = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(&bars[LB_AFULL], ph_af); ph_af ^= 1;
                tc_fence_after();
                for (int c = 0; c < p.n_chunks; ++c) {
                    const uint32_t buf = nacc & 1;
                    ++nacc;
                    mbar_wait(&bars[LB_ACCE + buf], ph_acce[buf]); ph_acce[buf] ^= 1;      // epilogue has drained this accumulator
                    tc_fence_after();
                    for (int ka = 0; ka < p.k_atoms; ++ka) {
                        mbar_wait(&bars[LB_FULL + stage_i], phase);
                        tc_fence_after();
                        const uint32_t a = sbase + L_A + ka * ATOM_A, b = sbase + L_RING + stage_i * LIN_SLAB;
                        umma_ss_w4(tmem + 256 * buf, umma_desc_sw128(a), umma_desc_sw128(b), LIN_IDESC, ka != 0);
                        umma_commit_w(&bars[LB_EMPTY + stage_i]);
                        if (++stage_i == LIN_RING_N) { stage_i = 0; phase ^= 1; }
                    }
                    umma_commit_w(&bars[LB_ACCF + buf]);
                }
                umma_commit_w(&bars[LB_AEMPTY]);
            }
        }
        __syncwarp();
    } else if (warp >= 10) {
        // ===================================================== 6 LayerNorm warps (fp32 row input): run one tile ahead of the GEMMs
        if (a_rows) {
            const int lw = warp - 10;                 // 0..5
            uint32_t ph_ae = 0;
            // the first tile is normalised by the 8 row warps (twice the loads in flight at kernel start); these warps start on the second.
            // Rows per warp: 24 (warps 0-3) or 16 (warps 4, 5) -- with four warps of 32 rows the LayerNorm (6-9 K cycles per tile) was
            // slower than the GEMM + epilogue of a 180-wide output (5 K).
            const int nb = lw < 4 ? 3 : 2;
            const int r0 = lw < 4 ? 24 * lw : 96 + 16 * (lw - 4);
            if (lw >= 4 && static_cast<int>(blockIdx.x) < p.n_tiles) mbar_arrive(&bars[LB_AFULL]);      // first tile: see the barrier's count
            for (int tile = blockIdx.x + gridDim.x; tile < p.n_tiles; tile += gridDim.x) {
                auto tok_of_row = [&](int r) -> int64_t {
                    const int64_t tk = static_cast<int64_t>(tile) * 128 + r;
                    return tk < p.num_tokens ? tk : static_cast<int64_t>(-1);
                };
                uint2 h0[4][3], h1[4][3], h2[4][3];               // up to 24 rows x 192 channels / 32 lanes, bf16: 72 registers
                ln_rows_hold<4>(p.x, p.ld_in, p.apply_ln, r0, lane, tok_of_row, h0);
                ln_rows_hold<4>(p.x, p.ld_in, p.apply_ln, r0 + 8, lane, tok_of_row, h1);
                if (nb == 3) ln_rows_hold<4>(p.x, p.ld_in, p.apply_ln, r0 + 16, lane, tok_of_row, h2);
                mbar_wait(&bars[LB_AEMPTY], ph_ae); ph_ae ^= 1;             // the previous tile's MMAs have read the A image
                ln_rows_dump<4>(sbase + L_A, r0, lane, h0);
                ln_rows_dump<4>(sbase + L_A, r0 + 8, lane, h1);
                if (nb == 3) ln_rows_dump<4>(sbase + L_A, r0 + 16, lane, h2);
                fence_proxy_async_smem();
                mbar_arrive(&bars[LB_AFULL]);
            }
        }
    } else {
        // ===================================================== 256 row threads
        const int cw8 = warp - 2, g = cw8 >> 2, q = warp & 3, row = q * 32 + lane;
        const uint32_t lanebase = static_cast<uint32_t>(q * 32) << 16;
        uint32_t ph_accf[2] = {0, 0}, nacc = 0, nplane = 0;
        if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) p.dbg[62] = clock64();
        if (a_rows && static_cast<int>(blockIdx.x) < p.n_tiles) {      // first tile: all 8 row warps normalise it (see the LayerNorm warps)
            const int tile = blockIdx.x;
            auto tok_of_row = [&](int r) -> int64_t {
                const int64_t tk = static_cast<int64_t>(tile) * 128 + r;
                return tk < p.num_tokens ? tk : static_cast<int64_t>(-1);
            };
            ln_rows_to_image(p.x, p.ld_in, p.apply_ln, sbase + L_A, cw8, lane, tok_of_row);
            fence_proxy_async_smem();
            named_bar_sync(1, 256);
            if (g == 0) mbar_arrive(&bars[LB_AFULL]);
        }
        bool first_tile = true;
        unsigned long long* dbg = (blockIdx.x == 0 && threadIdx.x == 64) ? p.dbg : nullptr;
        int it = 0;
        if (dbg) dbg[63] = clock64();
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            if (dbg && it < 8) dbg[it * 64 + 0] = clock64();
            const int64_t tok = static_cast<int64_t>(tile) * 128 + row;
            auto tok_of_row = [&](int r) -> int64_t {
                const int64_t tk = static_cast<int64_t>(tile) * 128 + r;
                return tk < p.num_tokens ? tk : static_cast<int64_t>(-1);
            };
            for (int c = 0; c < p.n_chunks; ++c) {
                const uint32_t buf = nacc & 1;
                ++nacc;
                mbar_wait(&bars[LB_ACCF + buf], ph_accf[buf]); ph_accf[buf] ^= 1;
                tc_fence_after();
                if (dbg && it < 8) dbg[it * 64 + 1 + 3 * c] = clock64();
                if (dbg && it < 8) dbg[it * 64 + 2 + 3 * c] = clock64();
                const uint32_t acc = tmem + lanebase + 256 * buf;
                if (!out_rows) {
                    // ---- accumulator + bias (+ GELU) -> bf16 -> staged plane rows -> one bulk copy per plane
                    const int n_ps = lin_n_pstage(p.k_atoms);
                    uint8_t* pst = sm + ((lin_off_stage(p.k_atoms) + 1023u) & ~1023u) + (nplane % n_ps) * LIN_PSTAGE;
                    ++nplane;
                    if (threadIdx.x == 64) {                  // the issuing thread: the copies that last read this buffer are done
                        if (n_ps == 2) bulk_wait_read1(); else bulk_wait_read0();
                    }
                    named_bar_sync(1, 256);
#pragma unroll 1
                    for (int pi = 0; pi < 3; ++pi) {
                        const int piece = 3 * g + pi;                       // 32 columns of the chunk
                        uint32_t v[32];
                        tmem_ld32(acc + 32 * piece, v);
                        tmem_ld_wait();
                        const float4* bb = reinterpret_cast<const float4*>(s_vec + c * LIN_NC + 32 * piece);
                        uint32_t pw[16];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float4 b4 = bb[k];
                            float f0 = __uint_as_float(v[4 * k]) + b4.x, f1 = __uint_as_float(v[4 * k + 1]) + b4.y;
                            float f2 = __uint_as_float(v[4 * k + 2]) + b4.z, f3 = __uint_as_float(v[4 * k + 3]) + b4.w;
                            if (p.act == SRK_LIN_ACT_GELU) { f0 = gelu_fast(f0); f1 = gelu_fast(f1); f2 = gelu_fast(f2); f3 = gelu_fast(f3); }
                            pw[2 * k] = pack_bf16x2(f0, f1);
                            pw[2 * k + 1] = pack_bf16x2(f2, f3);
                        }
                        const int plane = c * 3 + (piece >> 1);
                        const uint32_t key = static_cast<uint32_t>(tok + (((p.plane_phase_mask >> plane) & 1u) << 2)) & 7u;
                        const uint32_t dst = smem_u32(pst) + (piece >> 1) * ATOM_A + row * 128;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            st_shared_v4(dst + (((4 * (piece & 1) + k) ^ key) << 4), pw[4 * k], pw[4 * k + 1], pw[4 * k + 2], pw[4 * k + 3]);
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1, 256);
                    if (threadIdx.x == 64) {
                        const int64_t tok0 = static_cast<int64_t>(tile) * 128;
                        const int64_t left = p.num_tokens - tok0;
                        const uint32_t bytes = (left < 128 ? static_cast<uint32_t>(left) : 128u) * 128u;
#pragma unroll
                        for (int pl = 0; pl < 3; ++pl)
                            bulk_s2g(p.out_planes + (c * 3 + pl) * p.out_plane_stride + tok0 * 128, smem_u32(pst) + pl * ATOM_A, bytes);
                        bulk_commit();
                    }
                } else {
                    if (!(first_tile && c == 0) && !stage_alias) {      // the previous chunk's copies no longer read the staging rows
                        if (g == 0) bulk_wait_read0();
                        named_bar_sync(2 + q, 64);
                    }
                    // chunk c = output columns [180 c, 180 c + 180) of the ld_out-wide rows
                    stage_rows_and_bulk_store(acc, 0u, stage, stage, 32, s_vec + c * LIN_NC, p.y + c * SRK_DIM, p.ld_out, p.add_residual, q, g,
                                              lane, tok_of_row, p.act == SRK_LIN_ACT_GELU, nullptr, p.use_tmap ? &p.tmap_out : nullptr,
                                              c * SRK_DIM, tile * 128);
                    if (stage_alias && g == 0) {
                        bulk_wait_read0();
                        mbar_arrive(&bars[LB_DRAIN]);
                    }
                }
                tc_fence_before();
                mbar_arrive(&bars[LB_ACCE + buf]);
                if (dbg && it < 8) dbg[it * 64 + 3 + 3 * c] = clock64();
            }
            first_tile = false;
        }
        bulk_wait_read0();          // shared memory must outlive the bulk copies that read it
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn lin_encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    return fn;
}

cudaError_t launch_token_linear(const LinearParams& p_in, cudaStream_t stream) {
    LinearParams p = p_in;
    p.use_tmap = 0;
    if (p.out_mode == SRK_LIN_OUT_ROWS && p.ld_out != SRK_DIM && p.num_tokens < (1ll << 31)) {
        // (a failed encode just keeps the per-row copies)
        if (EncodeTiledFn enc = lin_encode_tiled_fn()) {
            const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(p.ld_out), static_cast<cuuint64_t>(p.num_tokens)};
            const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(p.ld_out) * sizeof(float)};
            const cuuint32_t box[2] = {SRK_DIM, 32};
            const cuuint32_t estr[2] = {1, 1};
            if (enc(&p.tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p.y, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                p.use_tmap = 1;
        }
    }
    static bool configured[SRK_MAX_DEVICES] = {};
    if (cudaError_t e = configure_smem_once(configured, token_linear_kernel, 232448); e != cudaSuccess) return e;
    const int sms = device_num_sms();
    const bool need_stage = p.out_mode == SRK_LIN_OUT_ROWS && p.k_atoms == 3;
    const uint32_t smem = lin_smem(p.k_atoms, need_stage, p.out_mode == SRK_LIN_OUT_PLANES) + 1024;
    const int grid = p.n_tiles < sms ? p.n_tiles : sms;
    return launch_pdl(token_linear_kernel, grid, LIN_THREADS, smem, stream, p);
}

}  // namespace srk
