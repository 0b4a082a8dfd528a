// Microbenchmark: how fast can all 148 SMs stream the SAME weight slabs from L2 into shared-memory rings with 1-D bulk TMA
// (the weight producer of swin_attn_kernel / swin_mlp_kernel: 288 KB per 128-token tile per SM)?  Unicast per CTA versus
// multicast inside a 2-CTA / 4-CTA cluster (each CTA fetches 1/csz of every slab and multicasts it to all CTAs of the cluster).
// The consumer releases a stage as soon as it is full (no MMA), so the result is the L2 -> SM delivery ceiling.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_stream_bench l2_stream_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* b, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(b)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_acq(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

// CSZ = 1: unicast.  CSZ > 1: cluster of CSZ CTAs, CTA r loads bytes [r * slab / CSZ, (r + 1) * slab / CSZ) of every slab into all CTAs.
// A stage of the ring is refilled once ALL CTAs of the cluster have consumed it (every consumer arrives on every CTA's empty barrier).
template <int CSZ>
__global__ void stream_kernel(const uint8_t* w, uint32_t slab, int nslab_total, int slabs_per_pass, int stages, int consumer_delay, unsigned long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw), sbase = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(sm);
    uint64_t* empty = full + 16;
    const uint32_t ring = sbase + 1024;
    uint32_t rank = 0;
    if (CSZ > 1) rank = cg::this_cluster().block_rank();
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CSZ); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (CSZ > 1) cg::this_cluster().sync(); else __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x == 0) {                 // producer
        uint32_t st = 0, ph = 0;
        for (int s = 0; s < nslab_total; ++s) {
            if (CSZ > 1) mbar_wait_cluster_acq(&empty[st], ph ^ 1); else mbar_wait(&empty[st], ph ^ 1);
            mbar_expect_tx(&full[st], slab);
            const uint8_t* src = w + static_cast<size_t>(s % slabs_per_pass) * slab;
            if (CSZ == 1) bulk_g2s(ring + st * slab, src, slab, &full[st]);
            else {
                const uint32_t part = slab / CSZ;
                bulk_g2s_mc(ring + st * slab + rank * part, src + rank * part, part, &full[st], static_cast<uint16_t>((1u << CSZ) - 1));
            }
            if (++st == static_cast<uint32_t>(stages)) { st = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {         // consumer
        uint32_t st = 0, ph = 0;
        for (int s = 0; s < nslab_total; ++s) {
            mbar_wait(&full[st], ph);
            if (consumer_delay > 0) { const long long c0 = clock64(); while (clock64() - c0 < consumer_delay) {} }
            if (CSZ == 1) mbar_arrive(&empty[st]);
            else for (uint32_t c = 0; c < CSZ; ++c) mbar_arrive_cluster(&empty[st], c);
            if (++st == static_cast<uint32_t>(stages)) { st = 0; ph ^= 1; }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (CSZ > 1) cg::this_cluster().sync();
    if (threadIdx.x == 0) out[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
}

template <int CSZ>
static void run(const uint8_t* w, uint32_t slab, int passes, int slabs_per_pass, int stages, int delay, unsigned long long* dout, int nsm) {
    const int grid = (nsm / CSZ) * CSZ;
    const size_t smem = 2048 + static_cast<size_t>(stages) * slab;
    cudaFuncSetAttribute(stream_kernel<CSZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CSZ; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const int total = passes * slabs_per_pass;
    float best = 1e30f;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchKernelEx(&cfg, stream_kernel<CSZ>, w, slab, total, slabs_per_pass, stages, delay, dout);
        cudaEventRecord(e1);
        if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("csz %d: launch failed: %s\n", CSZ, cudaGetErrorString(cudaGetLastError())); return; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    unsigned long long h[160];
    cudaMemcpy(h, dout, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0, mn = ~0ull; double avg = 0;
    for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; avg += h[i]; }
    avg /= grid;
    const double bytes = static_cast<double>(total) * slab;
    printf("csz %d slab %5u B stages %2d delay %4d: %7.1f us  per-SM delivered %6.1f B/clk (avg), %6.1f (slowest CTA); chip-wide %6.2f TB/s delivered, L2 reads %6.2f TB/s\n",
           CSZ, slab, stages, delay, best * 1e3, bytes / avg, bytes / mx, bytes * grid / (best * 1e-3) / 1e12, bytes * grid / CSZ / (best * 1e-3) / 1e12);
}

int main() {
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const size_t wbytes = 288 * 1024;
    uint8_t* w; cudaMalloc(&w, wbytes); cudaMemset(w, 1, wbytes);
    unsigned long long* dout; cudaMalloc(&dout, 160 * sizeof(unsigned long long));
    printf("SMs %d; every CTA streams the same 288 KB (12 slabs of 24 KB / 24 of 12 KB / 36 of 8 KB) 64 times\n", nsm);
    const int passes = 64;
    for (int delay : {0, 384}) {
        for (uint32_t slab : {24576u, 12288u, 8192u}) {
            const int spp = static_cast<int>(wbytes / slab);
            for (int stages : {3, 4, 6, 8}) {
                if (static_cast<size_t>(stages) * slab > 200 * 1024) continue;
                run<1>(w, slab, passes, spp, stages, delay, dout, nsm);
            }
        }
        for (int stages : {3, 6}) {
            run<2>(w, 24576u, passes, 12, stages, delay, dout, nsm);
            run<4>(w, 24576u, passes, 12, stages, delay, dout, nsm);
        }
    }
    return 0;
}
