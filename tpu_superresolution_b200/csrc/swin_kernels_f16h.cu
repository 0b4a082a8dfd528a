// swin_mlp_kernel with fp16 GEMM operands AND the GELU evaluated on packed halves (include/srk.h: SRK_OPERANDS_F16_HALF_GELU; rowops.cuh:
// gelu_pack2): the default MLP of the drop-in modules.  Against the bf16 variant the hidden activations are 5x closer to the exact GELU
// (fp16's 11-bit significand outweighs the half-precision polynomial) and a launch is 5 % faster (38.5 -> 36.6 us at the BASELINE
// shape: one MUFU and 9 instructions per PAIR of activations instead of 8 + 1 MUFU per element).
#define SRK_F16_OPERANDS 1
#define SRK_HALF_GELU 1
#define SRK_ONLY_MLP 1
#define swin_mlp_kernel swin_mlp_kernel_f16h
#define launch_swin_mlp launch_swin_mlp_f16h
#include "swin_kernels.cu"
