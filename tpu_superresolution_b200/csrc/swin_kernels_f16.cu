// The fused Swin kernels with fp16 GEMM operands ("tight" precision mode, include/srk.h: SRK_OPERANDS_F16): swin_kernels.cu compiled
// a second time.  fp16 has TF32's 11-bit significand (bf16: 8); accumulation, residual stream, LayerNorm and softmax are fp32 either way.
#define SRK_F16_OPERANDS 1
#define swin_attn_kernel swin_attn_kernel_f16
#define swin_mlp_kernel swin_mlp_kernel_f16
#define launch_swin_attn launch_swin_attn_f16
#define launch_swin_mlp launch_swin_mlp_f16
#include "swin_kernels.cu"
