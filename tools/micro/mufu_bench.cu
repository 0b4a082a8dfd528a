// Microbenchmark: MUFU.EX2 throughput per SM for f32, f16x2 and bf16x2 operands; LDS.32 / FADD issue rates (sm_100a).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
template <int MODE>
__global__ void k(float* out, int iters) {
    float x0 = threadIdx.x * 1e-3f - 2.f, x1 = x0 - 0.1f, x2 = x0 - 0.2f, x3 = x0 - 0.3f;
    uint32_t h0 = 0xb800b900u + threadIdx.x, h1 = h0 + 7, h2 = h0 + 11, h3 = h0 + 13;
    __shared__ float tab[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = i * 1e-4f;
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
        } else if (MODE == 1) {
            asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
            asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
        } else if (MODE == 2) {
            asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h0)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h1));
            asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h2)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h3));
        } else if (MODE == 3) {      // 4 independent LDS.32 + FADD
            x0 += tab[(idx) & 4095]; x1 += tab[(idx + 33) & 4095]; x2 += tab[(idx + 66) & 4095]; x3 += tab[(idx + 99) & 4095];
            idx += 1;
        } else if (MODE == 4) {      // FADD only
            asm volatile("add.f32 %0, %0, %1;" : "+f"(x0) : "f"(x1)); asm volatile("add.f32 %0, %0, %1;" : "+f"(x1) : "f"(x2));
            asm volatile("add.f32 %0, %0, %1;" : "+f"(x2) : "f"(x3)); asm volatile("add.f32 %0, %0, %1;" : "+f"(x3) : "f"(x0));
        } else if (MODE == 5) {      // tanh f32
            asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x0)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x1));
            asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x2)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x3));
        } else if (MODE == 6) {      // tanh f16x2
            asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h0)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h1));
            asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h2)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h3));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[32] = (float)(t1 - t0);
    out[threadIdx.x & 31] = x0 + x1 + x2 + x3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
}
template <int MODE>
void run(const char* name, int nthreads, int per_instr) {
    float* d; cudaMalloc(&d, 256);
    const int iters = 4096;
    k<MODE><<<148, nthreads>>>(d, iters); cudaDeviceSynchronize();
    k<MODE><<<148, nthreads>>>(d, iters); cudaDeviceSynchronize();
    float h[33]; cudaMemcpy(h, d, 33 * 4, cudaMemcpyDeviceToHost);
    double cyc = h[32];
    double ops = (double)iters * 4 * nthreads * per_instr;
    printf("%-12s threads/SM %4d: %8.0f cycles  -> %6.2f results/clk/SM  (%5.2f cycles per warp-instruction per SMSP)\n", name, nthreads, cyc,
           ops / cyc, cyc / (iters * 4.0 * (nthreads / 128.0)));
    cudaFree(d);
}
int main() {
    for (int nt : {128, 256, 512}) {
        run<0>("ex2.f32", nt, 1); run<1>("ex2.f16x2", nt, 2); run<2>("ex2.bf16x2", nt, 2); run<3>("lds+fadd", nt, 1);
        run<4>("fadd", nt, 1); run<5>("tanh.f32", nt, 1); run<6>("tanh.f16x2", nt, 2);
    }
    return 0;
}
