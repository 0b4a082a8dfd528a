// Row-thread helpers shared by the fused Swin kernels (swin_kernels.cu) and the token-linear kernel (linear_kernel.cu):
// LayerNorm -> bf16 operand image, accumulator -> image chunks, accumulator -> staged fp32 rows -> bulk (reduce-add) store, GELU.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "srk.h"
#include "umma.cuh"

namespace srk {

constexpr uint32_t ATOM_A = 16384;      // 128 rows x 128 B: one k-atom of a 128-row operand image

// ------------------------------------------------------------------------------------------------
// shared helpers for the 256 row threads
// ------------------------------------------------------------------------------------------------
// (x - mean) * rstd of 16 token rows per warp -> bf16 SW128 image at `xa` (3 k-atoms).  LayerNorm's affine
// (gamma, beta) is folded into the following GEMM's weights / bias at pack time (packing.py).
// Half-warp per token: 16 lanes x 3 float4 cover the 180 channels (45 float4) fully coalesced.  The 12 loads of a
// lane for 8 rows are issued before their first use; two such halves per call (see below).
// Core form: `row_ptr(pass)` returns the global address of the token row this lane's half-warp handles in `pass`
// (image row cw8 * 16 + 2 * pass + (lane >> 4)), or nullptr for a padding row.  A single warp runs dependent scalar
// code at one instruction per ~6 cycles, so callers keep the per-pass address arithmetic to an add or two.
template <typename PtrFn>
__device__ __forceinline__ void ln_rows_to_image_p(int apply_ln, uint32_t xa, int cw8, int lane, PtrFn row_ptr) {
    const int l16 = lane & 15;
    const bool live2 = l16 < 13;            // float4 index l16 + 32 < 45
    // Two rolled halves of four passes (8 rows each): half the straight-line code of an 8-pass body.  This routine runs once per tile
    // and role, so its instructions are fetched cold every time (ncu: the kernel's no-instruction stalls concentrate here; the first
    // tile of a launch spent ~14 K cycles in it) -- the second half now hits the instruction cache, at the price of two exposed
    // memory latencies (12 loads in flight per lane) instead of one.
#pragma unroll 1
    for (int hb = 0; hb < 2; ++hb) {
        float4 v[4][3];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
            const float* rp = row_ptr(4 * hb + pp);
            const float4* src = reinterpret_cast<const float4*>(rp) + l16;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            v[pp][0] = rp ? ld_cg_f4(src) : z;              // through L2 (see umma.cuh: image progress counters)
            v[pp][1] = rp ? ld_cg_f4(src + 16) : z;
            v[pp][2] = (rp && live2) ? ld_cg_f4(src + 32) : z;
        }
        // statistics of all passes first, then the 4 butterfly rounds over all passes at once: the shuffles of a round are
        // independent, so their latency overlaps instead of serialising dependent steps
        float s[4], q[4];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
            float4(&w)[3] = v[pp];
            s[pp] = 0.f; q[pp] = 0.f;
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
                s[pp] += (w[jj].x + w[jj].y) + (w[jj].z + w[jj].w);
                q[pp] = fmaf(w[jj].x, w[jj].x, q[pp]); q[pp] = fmaf(w[jj].y, w[jj].y, q[pp]);
                q[pp] = fmaf(w[jj].z, w[jj].z, q[pp]); q[pp] = fmaf(w[jj].w, w[jj].w, q[pp]);
            }
        }
        if (apply_ln) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    s[pp] += __shfl_xor_sync(0xffffffffu, s[pp], o);
                    q[pp] += __shfl_xor_sync(0xffffffffu, q[pp], o);
                }
            }
        }
        const uint32_t r0 = cw8 * 16 + 8 * hb + (lane >> 4);
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
            float4(&w)[3] = v[pp];
            if (apply_ln) {
                const float mean = s[pp] * (1.0f / SRK_DIM);
                const float var = fmaxf(fmaf(-mean, mean, q[pp] * (1.0f / SRK_DIM)), 0.f);
                const float rstd = rsqrtf(var + 1e-5f);
                const float nm = -mean * rstd;
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    w[jj].x = fmaf(w[jj].x, rstd, nm); w[jj].y = fmaf(w[jj].y, rstd, nm);
                    w[jj].z = fmaf(w[jj].z, rstd, nm); w[jj].w = fmaf(w[jj].w, rstd, nm);
                }
                if (!live2) w[2] = make_float4(0.f, 0.f, 0.f, 0.f);      // padded channels 180..191 stay exactly zero
            }
            const uint32_t r = r0 + 2 * pp;
            const uint32_t off = xa + sw128_off(r, l16 >> 1) + (l16 & 1) * 8;
#pragma unroll
            for (int jj = 0; jj < 3; ++jj)      // channel 4f = 64*jj + 4*l16 -> atom jj, chunk l16>>1, byte (l16&1)*8
                st_shared_v2(off + jj * ATOM_A, pack_op2(w[jj].x, w[jj].y), pack_op2(w[jj].z, w[jj].w));
        }
    }
}
template <typename TokFn>
__device__ __forceinline__ void ln_rows_to_image(const float* __restrict__ x, int ld, int apply_ln, uint32_t xa, int cw8,
                                                 int lane, TokFn tok_of_row) {
    ln_rows_to_image_p(apply_ln, xa, cw8, lane, [&](int pass) -> const float* {
        const int64_t tok = tok_of_row(cw8 * 16 + pass * 2 + (lane >> 4));
        return tok >= 0 ? x + tok * ld : nullptr;
    });
}


// Split form of ln_rows_to_image for dedicated LayerNorm warps that run AHEAD of the GEMMs: ln_rows_hold() loads and
// normalises NPASS x 2 token rows (half-warp per row) and keeps them as packed bf16 in registers (6 per row pair) while
// the operand image is still being read by the previous tile's MMAs; ln_rows_dump() writes them into the image once it is
// free -- the global-memory latency of the next tile is then off the critical path, only the dump (a few hundred cycles) is on it.
template <int NPASS, typename PtrFn>
__device__ __forceinline__ void ln_rows_hold_p(int apply_ln, int lane, PtrFn row_ptr, uint2 (&held)[NPASS][3]) {
    const int l16 = lane & 15;
    const bool live2 = l16 < 13;
    float4 v[NPASS][3];
#pragma unroll
    for (int pass = 0; pass < NPASS; ++pass) {
        const float* rp = row_ptr(pass);
        const float4* src = reinterpret_cast<const float4*>(rp) + l16;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        v[pass][0] = rp ? ld_cg_f4(src) : z;              // through L2 (see umma.cuh: image progress counters)
        v[pass][1] = rp ? ld_cg_f4(src + 16) : z;
        v[pass][2] = (rp && live2) ? ld_cg_f4(src + 32) : z;
    }
    float s[NPASS], q[NPASS];
#pragma unroll
    for (int pass = 0; pass < NPASS; ++pass) {
        float4(&w)[3] = v[pass];
        s[pass] = 0.f; q[pass] = 0.f;
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
            s[pass] += (w[jj].x + w[jj].y) + (w[jj].z + w[jj].w);
            q[pass] = fmaf(w[jj].x, w[jj].x, q[pass]); q[pass] = fmaf(w[jj].y, w[jj].y, q[pass]);
            q[pass] = fmaf(w[jj].z, w[jj].z, q[pass]); q[pass] = fmaf(w[jj].w, w[jj].w, q[pass]);
        }
    }
    if (apply_ln) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
            for (int pass = 0; pass < NPASS; ++pass) {
                s[pass] += __shfl_xor_sync(0xffffffffu, s[pass], o);
                q[pass] += __shfl_xor_sync(0xffffffffu, q[pass], o);
            }
        }
    }
#pragma unroll
    for (int pass = 0; pass < NPASS; ++pass) {
        float4(&w)[3] = v[pass];
        if (apply_ln) {
            const float mean = s[pass] * (1.0f / SRK_DIM);
            const float var = fmaxf(fmaf(-mean, mean, q[pass] * (1.0f / SRK_DIM)), 0.f);
            const float rstd = rsqrtf(var + 1e-5f);
            const float nm = -mean * rstd;
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
                w[jj].x = fmaf(w[jj].x, rstd, nm); w[jj].y = fmaf(w[jj].y, rstd, nm);
                w[jj].z = fmaf(w[jj].z, rstd, nm); w[jj].w = fmaf(w[jj].w, rstd, nm);
            }
            if (!live2) w[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) held[pass][jj] = make_uint2(pack_op2(w[jj].x, w[jj].y), pack_op2(w[jj].z, w[jj].w));
    }
}
template <int NPASS, typename TokFn>
__device__ __forceinline__ void ln_rows_hold(const float* __restrict__ x, int ld, int apply_ln, int row0, int lane, TokFn tok_of_row,
                                             uint2 (&held)[NPASS][3]) {
    ln_rows_hold_p<NPASS>(apply_ln, lane, [&](int pass) -> const float* {
        const int64_t tok = tok_of_row(row0 + pass * 2 + (lane >> 4));
        return tok >= 0 ? x + tok * ld : nullptr;
    }, held);
}
template <int NPASS>
__device__ __forceinline__ void ln_rows_dump(uint32_t xa, int row0, int lane, const uint2 (&held)[NPASS][3]) {
    const int l16 = lane & 15;
#pragma unroll
    for (int pass = 0; pass < NPASS; ++pass) {
        const uint32_t r = row0 + 2 * pass + (lane >> 4);
        const uint32_t off = xa + sw128_off(r, l16 >> 1) + (l16 & 1) * 8;
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) st_shared_v2(off + jj * ATOM_A, held[pass][jj].x, held[pass][jj].y);
    }
}

// 32 fp32 accumulators (+ per-column bias from smem | * scale) -> 4 x 16-byte bf16 chunks of one image row.
template <bool HAS_BIAS, bool HAS_SCALE>
__device__ __forceinline__ void store_row_chunks(uint32_t img_atom, uint32_t row, uint32_t c16base, const uint32_t (&v)[32],
                                                 const float* bias, float scale) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * k + e]);
        if (HAS_BIAS) {
            const float4 b0 = reinterpret_cast<const float4*>(bias)[2 * k], b1 = reinterpret_cast<const float4*>(bias)[2 * k + 1];
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
        }
        if (HAS_SCALE) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] *= scale;
        }
        st_shared_v4(img_atom + sw128_off(row, c16base + k), pack_op2(f[0], f[1]), pack_op2(f[2], f[3]),
                     pack_op2(f[4], f[5]), pack_op2(f[6], f[7]));
    }
}

// gelu(x) = x Phi(x).  Phi(x) = 0.5 (1 + erf(x / sqrt 2)) is evaluated as 0.5 (1 + tanh(x (c1 + c3 u + c5 u^2))),
// u = min(x^2, 64): minimax fit, |error| <= 2.6e-5 on the GELU output for all x, plus the MUFU.TANH error (2^-11 rel.)
// -- both far below the bf16 rounding (2^-9 rel.) applied to the result right after.  One MUFU per element.
__device__ __forceinline__ float gelu_fast(float x) {
    const float u2 = fminf(x * x, 64.0f);
    const float qv = x * fmaf(u2, fmaf(u2, -3.51517176e-04f, 3.70056486e-02f), 7.97507881e-01f);
    const float hx = 0.5f * x;
    return fmaf(hx, tanh_approx(qv), hx);
}


// Two GELU results as a packed operand pair.  With SRK_HALF_GELU (fp16 operands only) the whole evaluation runs on packed halves:
// 9 instructions and ONE MUFU per PAIR instead of 8 + 1 MUFU per element.  The fp16 arithmetic costs accuracy (max abs 2.3e-3 near
// x = 2.8, rms 4e-4 on |x| < 3), but the bf16 operand path rounds the fp32 result to 8 bits right after (rms 2e-3), so this is still
// five times closer to the exact GELU than the bf16 path; the tight mode keeps the fp32 evaluation.
__device__ __forceinline__ uint32_t gelu_pack2(float x0, float x1) {
#if defined(SRK_HALF_GELU) && defined(SRK_F16_OPERANDS)
    const uint32_t xu = pack_f16x2(x0, x1);
    const __half2 x = *reinterpret_cast<const __half2*>(&xu);
    const __half2 u2 = __hmin2(__hmul2(x, x), __float2half2_rn(64.0f));
    const __half2 pl = __hfma2(u2, __hfma2(u2, __float2half2_rn(-3.51517176e-04f), __float2half2_rn(3.70056486e-02f)), __float2half2_rn(7.97507881e-01f));
    const __half2 q = __hmul2(x, pl);
    uint32_t tu;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(tu) : "r"(*reinterpret_cast<const uint32_t*>(&q)));
    const __half2 t = *reinterpret_cast<const __half2*>(&tu);
    const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
    const __half2 r = __hfma2(hx, t, hx);
    return *reinterpret_cast<const uint32_t*>(&r);
#else
    return pack_op2(gelu_fast(x0), gelu_fast(x1));
#endif
}

constexpr uint32_t ROW_BYTES = SRK_DIM * 4;     // 720
// Bulk copies of the 32 staged rows of lane quadrant q (this lane's row at `my_row`): one copy per maximal run of tokens that are
// contiguous in memory (and in the staging buffer).  Whole warp; the caller commits the bulk group.
template <typename TokFn>
__device__ __forceinline__ void issue_row_runs(uint8_t* my_row, int main_rows, float* __restrict__ y, int ld_out, int add_residual, int q,
                                               int lane, TokFn tok_of_row) {
    const int64_t tok = tok_of_row(q * 32 + lane);
    const int64_t prev = __shfl_up_sync(0xffffffffu, tok, 1);
    const bool valid = tok >= 0;
    const bool start = valid && (lane == 0 || lane == main_rows || ld_out != SRK_DIM || prev < 0 || tok != prev + 1);
    const uint32_t m_start = __ballot_sync(0xffffffffu, start);
    const uint32_t m_stop = m_start | ~__ballot_sync(0xffffffffu, valid);      // next start or first invalid row ends a run
    // Issue from CONVERGENT code with warp-uniform operands: inside `if (start)` the compiler wraps the bulk-copy instruction
    // (its operands live in uniform registers) in an ELECT / R2UR.BROADCAST waterfall loop, one trip per run.  Here every lane
    // computes its run, then the warp walks the run starts; the values of the run's first lane are broadcast and one elected
    // lane issues the copy (all copies of a warp therefore belong to that lane's bulk groups).
    const uint32_t after = lane == 31 ? 0u : (m_stop >> (lane + 1));
    const int len = after ? __ffs(after) : (32 - lane);
    const uint32_t my_bytes = static_cast<uint32_t>(len) * ROW_BYTES;
    const uint32_t my_src = smem_u32(my_row);
    float* const my_dst = y + (valid ? tok : 0) * ld_out;
    for (uint32_t m = m_start; m != 0; m &= m - 1) {
        const int l = __ffs(m) - 1;
        const uint32_t bytes = __shfl_sync(0xffffffffu, my_bytes, l);
        const uint32_t src = __shfl_sync(0xffffffffu, my_src, l);
        const uint64_t dst = __shfl_sync(0xffffffffu, reinterpret_cast<uint64_t>(my_dst), l);
        if (add_residual)
            asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                         "@e cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;\n\t}" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
        else
            asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                         "@e cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\t}" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    }
}

// Final epilogue shared by K1/K2: accumulator (+bias) -> fp32 rows staged in shared memory in the exact global row
// layout (720 B per token row) -> the TMA engine writes them out with one bulk copy per contiguous run of tokens.
// With add_residual the copy is `cp.reduce.async.bulk ... add.f32`: the residual stream is updated in place
// (y += delta), so the shortcut is never loaded into the SM.  Thread = accumulator row; the two groups fill columns
// [96 g, 96 g + 96) of the same rows.  `half` / `nhalf` let a caller with < 92 KB of staging do the 32 rows of each
// lane quadrant in two passes of 16.
// Staging geometry: rows 0 .. main_rows-1 of every lane quadrant live at stage_main + (q * main_rows + r) * 720, the
// remaining rows at stage_tail + (q * (32 - main_rows) + r - main_rows) * 720 (K1 has no single 92 KB hole).
template <typename TokFn>
__device__ __forceinline__ void stage_rows_and_bulk_store(uint32_t tmem_acc, uint32_t lanebase, uint8_t* stage_main, uint8_t* stage_tail,
                                                          int main_rows, const float* s_bias, float* __restrict__ y, int ld_out,
                                                          int add_residual, int q, int g, int lane, TokFn tok_of_row, int act_gelu = 0,
                                                          unsigned long long* tl = nullptr, const void* tmap = nullptr, int tcol = 0,
                                                          int trow = 0, bool issue = true) {
    uint8_t* const my_row = lane < main_rows ? stage_main + (q * main_rows + lane) * ROW_BYTES
                                             : stage_tail + (q * (32 - main_rows) + lane - main_rows) * ROW_BYTES;
    float* dst = reinterpret_cast<float*>(my_row);
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        const int c = 3 * g + ci;
        uint32_t v[32];
        tmem_ld32(tmem_acc + lanebase + 32 * c, v);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (32 * c + 4 * k < SRK_DIM) {
                const float4 b = reinterpret_cast<const float4*>(s_bias + 32 * c)[k];
                float4 o;
                o.x = __uint_as_float(v[4 * k + 0]) + b.x;
                o.y = __uint_as_float(v[4 * k + 1]) + b.y;
                o.z = __uint_as_float(v[4 * k + 2]) + b.z;
                o.w = __uint_as_float(v[4 * k + 3]) + b.w;
                if (act_gelu) { o.x = gelu_fast(o.x); o.y = gelu_fast(o.y); o.z = gelu_fast(o.z); o.w = gelu_fast(o.w); }
                *reinterpret_cast<float4*>(dst + 32 * c + 4 * k) = o;
            }
        }
    }
    if (tl) tl[50] = clock64();
    fence_proxy_async_smem();                   // generic-proxy smem writes -> visible to the bulk copy engine
    if (!issue) return;                         // the caller synchronises and issues the copies (swin_attn_kernel: one box per window)
    named_bar_sync(2 + q, 64);                  // both groups of this lane quadrant have written their columns
    if (tl) tl[51] = clock64();
    if (g == 0 && tmap != nullptr) {
        // rows wider than the 180 staged columns: ONE tensor-map TMA store of this quadrant's 32 x 180 box at (column tcol, row trow +
        // 32 q) of the 2-D output (rows past the last token are clipped by the copy engine); main_rows must be 32.  Per-row bulk
        // copies (720 B each, ~600 cycles of issue per copy) made the 180 -> 720 layer of DAT's SGFN cost 119 us per launch.
        const uint32_t src = smem_u32(stage_main + q * 32 * ROW_BYTES);
        const int r0 = trow + 32 * q;
        if (add_residual)
            asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                         "@e cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%1, %2}], [%3];\n\t}" ::"l"(reinterpret_cast<uint64_t>(tmap)),
                         "r"(tcol), "r"(r0), "r"(src) : "memory");
        else
            asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                         "@e cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\n\t}" ::"l"(reinterpret_cast<uint64_t>(tmap)),
                         "r"(tcol), "r"(r0), "r"(src) : "memory");
        bulk_commit();
        __syncwarp();
    } else if (g == 0) {
        issue_row_runs(my_row, main_rows, y, ld_out, add_residual, q, lane, tok_of_row);
        bulk_commit();
        __syncwarp();
    }
}


}  // namespace srk