// The fused Swin kernels with fp16 GEMM operands ("tight" precision mode, include/srk.h: SRK_OPERANDS_F16): swin_kernels.cu compiled
// a second time.  fp16 has TF32's 11-bit significand (bf16: 8); accumulation, residual stream, LayerNorm and softmax are fp32 either way.
#define SRK_F16_OPERANDS 1
#include "swin_kernels.cu"
