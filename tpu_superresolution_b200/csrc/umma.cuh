// Blackwell (sm_100a) primitives used by the fused window-attention kernels:
// mbarrier, 1-D bulk TMA (cp.async.bulk), tcgen05 alloc / mma / commit / ld / st and the
// UMMA shared-memory / instruction descriptors.  Raw PTX, no CUTLASS dependency.
//
// Operand layout convention used everywhere in this repo ("SW128 K-major image"):
//   a [rows x K] bf16 operand is cut into K/64 "k-atoms"; one k-atom is rows x 128 B,
//   row r at byte r*128, and the 16-byte chunk c (0..7) of row r is stored at chunk
//   position c ^ (r & 7) (the 128-byte swizzle).  Every k-atom base is 1024-byte aligned.
//   UMMA descriptor: SWIZZLE_128B, SBO = 1024 B (8-row group pitch), LBO unused (=1).
//   A k-step (16 bf16 = 32 B) inside the atom is addressed by adding 32 B to the start.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace srk {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time; a poller of several barriers must not)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// warp-uniform probe: every lane gets the same answer (a phase that any lane saw complete is complete)
__device__ __forceinline__ bool mbar_test_wait_w(uint64_t* bar, uint32_t parity) {
    return __any_sync(0xffffffffu, mbar_test_wait(bar, parity)) != 0;
}
#ifndef SRK_WAIT_TIMEOUT_CYCLES
#define SRK_WAIT_TIMEOUT_CYCLES (4000000000ll)   // ~2 s at 1.9 GHz: a deadlock traps instead of hanging the GPU
#endif
// With SRK_OOL_TIMEOUT (defined by a translation unit before this header) the timeout report is an out-of-line call: an inlined
// printf site costs ~25 instructions at each of the ~20 waits of a kernel, which matters for the 130+ KB fused Swin kernels (cold
// instruction cache on the first tile).  The window-attention kernel is faster with the inline form (the call costs it stack spills).
static __device__ __noinline__ void mbar_timeout_trap(uint32_t bar, uint32_t parity) {
    printf("srk: mbarrier wait timeout (block %d thread %d bar@%u parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, bar, parity);
    __trap();
}
#ifdef SRK_OOL_TIMEOUT
// the whole spin loop out of line: ~12 instructions less at each of the ~60 waits of a fused kernel (instruction-cache footprint)
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && clock64() - t0 > SRK_WAIT_TIMEOUT_CYCLES) mbar_timeout_trap(bar, parity);
    }
}
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
#ifdef SRK_OOL_WAIT
    mbar_wait_slow(smem_u32(bar), parity);
    return;
#endif
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
#ifndef SRK_OOL_TIMEOUT
        if (clock64() - t0 > SRK_WAIT_TIMEOUT_CYCLES) {
            printf("srk: mbarrier wait timeout (block %d thread %d bar@%u parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
#else
        if (clock64() - t0 > SRK_WAIT_TIMEOUT_CYCLES) mbar_timeout_trap(smem_u32(bar), parity);
#endif
    }
}

// ------------------------------------------------------------------ proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {   // generic-proxy smem writes -> visible to UMMA / TMA
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------ 1-D bulk TMA (global -> shared)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------ bulk TMA (shared -> global), plain or fp32 reduce-add
__device__ __forceinline__ void bulk_s2g(void* gdst, uint32_t smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_s2g_add_f32(void* gdst, uint32_t smem_src, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// all bulk copies committed by this thread have completed (their global writes are performed), not just read their source
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------ fine-grained ordering between kernels (image progress counters)
// Consecutive fused Swin kernels run under programmatic dependent launch; instead of waiting for the whole previous grid
// (griddepcontrol.wait) a kernel may wait, tile by tile, for the image its tile belongs to: the producer kernel adds to a
// per-image counter once a tile's writes have completed, the consumer polls it with acquire semantics and then reads the
// tile through L2 (ld.global.cg: the SM's L1 may still hold lines of the residual stream from before the update).
// (not volatile, no memory clobber: the compiler may batch and hoist these like ordinary loads.  What keeps them behind a
//  progress wait is a data dependence: progress_wait() returns an opaque zero that the caller adds to the base pointer.)
__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {
#ifdef SRK_LN_LDG
    return __ldg(p);
#endif
    float4 v;
    asm("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// whole warp: returns once *p >= target (p == nullptr: no wait).  The result is always 0, but opaque to the compiler: add it
// to the base pointer of the loads that must come after the wait.
__device__ __forceinline__ int progress_wait(const int* p, int target) {
    if (p != nullptr) {
        if ((threadIdx.x & 31) == 0 && ld_acquire_gpu(p) < target) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(p) < target) {
                __nanosleep(40);
                if (clock64() - t0 > SRK_WAIT_TIMEOUT_CYCLES) mbar_timeout_trap(0xfffffffeu, static_cast<uint32_t>(target));
            }
        }
        __syncwarp();
    }
    int z;
    asm volatile("mov.u32 %0, 0;" : "=r"(z)::"memory");
    return z;
}

// ------------------------------------------------------------------ TMEM allocation (one full warp)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B (see header comment).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);   // [0,14)  start address >> 4
    d |= static_cast<uint64_t>(1) << 16;                       // [16,30) leading byte offset (unused, =1)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;               // [32,46) stride byte offset: 8 rows * 128 B
    d |= static_cast<uint64_t>(1) << 46;                       // [46,48) descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;                       // [61,64) SWIZZLE_128B
    return d;
}
// Instruction descriptor: kind::f16, A = B = bf16, D = fp32, both operands K-major, dense.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}
// ... A = B = fp16 (a_format = b_format = 0)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// Operand type of the translation unit: bf16 (default), or fp16 with SRK_F16_OPERANDS -- the "tight" precision mode: an 11-bit
// significand like TF32's; safe in range because every GEMM input here is LayerNorm output, a softmax probability or a
// weight-bounded projection of those (swin_kernels_f16.cu compiles swin_kernels.cu a second time with it).
#ifdef SRK_F16_OPERANDS
__host__ __device__ constexpr uint32_t umma_idesc_op(int M, int N) { return umma_idesc_f16(M, N); }
#else
__host__ __device__ constexpr uint32_t umma_idesc_op(int M, int N) { return umma_idesc_bf16(M, N); }
#endif
// byte offset of 16-byte chunk c16 of row r inside one SW128 k-atom
__device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t c16) { return r * 128u + ((c16 ^ (r & 7u)) << 4); }

// ------------------------------------------------------------------ MMA issue (one thread)
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T   (A: 128 lanes x K/2 columns, two bf16 per 32-bit column)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Warp-uniform forms: the WHOLE warp executes the call with identical operands and one elected lane issues the
// instruction.  Inside an `if (lane == 0)` region the compiler must assume divergent operands and wraps every
// tcgen05.mma (its descriptors live in uniform registers) in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop
// -- ~15 dependent instructions, ~90 cycles per MMA, which made the issuing thread the bottleneck of the fused
// kernels.  Issued from convergent code the descriptors are computed on the uniform datapath directly.
__device__ __forceinline__ void umma_ss_w(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Four consecutive k-steps (K = 64: one SW128 k-atom) of an SS MMA with one election: descriptors advance by
// 32 bytes (>> 4 = 2) per step.  `accumulate_first` applies to step 0; steps 1..3 always accumulate.
__device__ __forceinline__ void umma_ss_w4(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate_first) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "add.u64 a1, %1, 2;\n\t add.u64 b1, %2, 2;\n\t"
        "add.u64 a2, %1, 4;\n\t add.u64 b2, %2, 4;\n\t"
        "add.u64 a3, %1, 6;\n\t add.u64 b3, %2, 6;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, 1;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first)
        : "memory");
}
// The TS form: A = packed bf16 in TMEM, 8 columns (16 k-values) per step; B advances by BSTEP descriptor units
// (16 bytes each) per step: 2 for a K-major SW128 operand, 128 for the MN-major V operand (16 key rows of 128 B).
template <int BSTEP = 2>
__device__ __forceinline__ void umma_ts_w4(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate_first) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        ".reg .b32 a1, a2, a3;\n\t"
        ".reg .b64 b1, b2, b3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "add.u32 a1, %1, 8;\n\t add.u64 b1, %2, %5;\n\t"
        "add.u32 a2, %1, 16;\n\t add.u64 b2, %2, %6;\n\t"
        "add.u32 a3, %1, 24;\n\t add.u64 b3, %2, %7;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %3, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %3, 1;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate_first), "n"(BSTEP), "n"(2 * BSTEP), "n"(3 * BSTEP)
        : "memory");
}
__device__ __forceinline__ void umma_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar))
        : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------ TMEM <-> registers (warp-collective)
// 32x32b: lane t of the warp touches TMEM lane (quadrant base + t); 32 consecutive columns -> 32 registers.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// launch_dependents: the next kernel in the stream may start being scheduled (its CTAs become resident as SMs free up and run
// their prologue: barrier init, TMEM allocation, constant weights streaming into the ring).  wait: blocks until the previous kernel
// has completed and its global writes are visible -- executed by every thread before it touches activations.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits), .y = hi
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));      // (first source -> upper half)
    return r;
}
__device__ __forceinline__ uint32_t pack_op2(float lo, float hi) {          // two GEMM-operand elements, see umma_idesc_op
#ifdef SRK_F16_OPERANDS
    return pack_f16x2(lo, hi);
#else
    return pack_bf16x2(lo, hi);
#endif
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace srk
