// Internal launch interface between capi.cu and the kernel translation units.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "srk.h"

namespace srk {

struct AttnParams {
    const float* x;
    float* y;
    const uint8_t* wstream;
    const float* vec;
    const float* mask;
    int mode, H, W, shift, nwx, nw_img;
    int total_windows, n_tiles;
    int ld_in, ld_out;
    int apply_ln, add_residual, mask_mode, mask_nw;
    unsigned long long* dbg;   // optional timeline buffer (srk_debug_set_timeline), else nullptr
    int stagger;               // start skew in cycles per (CTA index mod 4)
    // image progress counters (umma.cuh), all optional: after a tile's writes have completed prog_sig[image] += 2 (windows);
    // with wait_target > 0 the kernel waits per tile for prog_wait[image] >= wait_target instead of for the whole previous grid
    int* prog_sig;
    const int* prog_wait;
    int wait_target;
    int use_tmap;              // swin_attn_kernel: the output tensor map (second kernel argument) is valid
    uint32_t div_img_m, div_img_s, div_nwx_m, div_nwx_s;      // multiply-high constants of / nw_img and / nwx (launch_swin_attn sets them)
};

struct MlpParams {
    const float* x;
    float* y;
    const uint8_t* wstream;
    const float* vec;
    int64_t num_tokens;
    int n_tiles;
    int ld_in, ld_out;
    int apply_ln, add_residual;
    unsigned long long* dbg;
    int stagger;
    int* prog_sig;             // += 1 per finished tile of the image (see AttnParams)
    const int* prog_wait;
    int wait_target;
    int tokens_per_image;      // tile -> image (a multiple of 128 when the counters are used)
};

// swin_layer_kernel (swin_kernels.cu): all blocks of one BasicLayer in one persistent launch
constexpr int SRK_LAYER_MAX_BLOCKS_K = 8;
struct LayerBlockW {
    const uint8_t* attn_w;
    const float* attn_vec;
    const uint8_t* mlp_w;
    const float* mlp_vec;
    int shift, pad;
};
struct LayerParams {
    float* y;                  // residual stream (B, H*W, ld) fp32, updated in place
    int ld;
    int B, H, W, nwx, nw_img;
    int T;                     // tiles per half-block = B * H * W / 128
    int tiles_per_image;
    int n_blocks, n_items;     // n_items = n_blocks * 2 * T
    int* progress;             // 2 * B counters, zeroed: [0, B) windows finished by attention, [B, 2B) tiles finished by the MLP
    LayerBlockW blk[SRK_LAYER_MAX_BLOCKS_K];
    unsigned long long* dbg;
};
cudaError_t launch_swin_layer(const LayerParams& p, cudaStream_t stream);

// token_linear_kernel (linear_kernel.cu)
struct LinearParams {
    const float* x;              // SRK_LIN_A_ROWS: fp32 token rows
    int ld_in, apply_ln;
    const uint8_t* a_planes;     // SRK_LIN_A_PLANES: bf16 planes [k-atom][tok][128 B]
    int64_t a_plane_stride;
    int a_mode, k_atoms;         // k_atoms = K / 64: 3 (K = 192) or 6 (K = 384)
    int64_t num_tokens;
    int n_tiles;
    const uint8_t* wstream;      // n_chunks x k_atoms slabs of 192 rows x 128 B
    const float* bias;           // n_chunks x 192
    int n_chunks, act;
    int out_mode;
    uint8_t* out_planes;         // SRK_LIN_OUT_PLANES: 3 planes per chunk
    int64_t out_plane_stride;
    uint32_t plane_phase_mask;   // bit p set: plane p is swizzled with (tok + 4) & 7 instead of tok & 7
    float* y;                    // SRK_LIN_OUT_ROWS
    int ld_out, add_residual;
    unsigned long long* dbg;     // optional timeline buffer (srk_debug_set_timeline)
    // SRK_LIN_OUT_ROWS with ld_out != 180 (a chunk's 180 columns are then 720-byte pieces of wider rows): the output as a 2-D
    // tensor (ld_out x num_tokens fp32), box 180 x 32 -- one tensor-map TMA store per lane quadrant instead of one bulk copy per row
    int use_tmap;
    alignas(64) CUtensorMap tmap_out;
};

// winattn_kernel (winattn_kernel.cu)
struct WinAttnParams {
    const uint8_t *q_planes, *k_planes, *v_planes;   // plane 0 (head pair 0) of each operand
    int64_t plane_stride;        // bytes between planes
    const float* tab;            // [2 * pairs][TABF] strided bias table, * log2(e)
    int c0;                      // table index of (query i, key j) = c0 + SY * (yi - yj) + (xi - xj)
    const uint8_t* zero_page;    // 8 KB: 4 KB of zeros (padded k rows) + 4 KB v padding pattern (ones column, swizzled per row & 7)
    int B, H, W;
    int koff;                    // key-window offset relative to the query window (OCAB: -4)
    int shift_y, shift_x, wrap, mask_shift;
    const float* emask;          // explicit (nW, 256, NK) mask or nullptr
    int emask_nw;
    int n_heads, nwy, nwx, n_items;
    int out_mode;                // 0: bf16 planes (k-atoms of the proj GEMM's A operand), 1: fp32 rows
    uint8_t* o_planes;
    int64_t o_plane_stride;
    float* o_rows;
    int o_ld, o_col0;
    unsigned long long* dbg;     // optional timeline buffer (srk_debug_set_timeline)
    int stagger;                 // cycles group 1 starts behind group 0
};


// Per-device one-time state.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are properties of the CURRENT device
// (the Python binding runs every call under the device of its tensors), so "configured once" must be once per device, not per process.
constexpr int SRK_MAX_DEVICES = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < SRK_MAX_DEVICES) ? dev : 0;
}
inline int device_num_sms() {
    static int sms[SRK_MAX_DEVICES] = {};
    const int dev = current_device();
    if (sms[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = n > 0 ? n : 148;
    }
    return sms[dev];
}
// cudaFuncSetAttribute(kern, MaxDynamicSharedMemorySize, bytes) once per (call site, device); `flags` is the call site's static array
template <typename K>
inline cudaError_t configure_smem_once(bool (&flags)[SRK_MAX_DEVICES], K kern, int bytes) {
    const int dev = current_device();
    if (!flags[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        flags[dev] = true;
    }
    return cudaSuccess;
}

extern unsigned long long* g_timeline;
extern int g_stagger_attn, g_stagger_mlp, g_stagger_winattn, g_pdl;

// Launch with programmatic stream serialisation: the kernel may be scheduled while its predecessor in the stream is still running
// (once every CTA of the predecessor has executed griddepcontrol.launch_dependents, or exited); it orders itself with
// griddepcontrol.wait (umma.cuh: pdl_wait) before it touches anything the predecessor writes.
template <typename P>
inline cudaError_t launch_pdl(void (*kern)(const P), int grid, int threads, size_t smem, cudaStream_t stream, const P& p) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}
template <typename P, typename Q>
inline cudaError_t launch_pdl2(void (*kern)(const P, const Q), int grid, int threads, size_t smem, cudaStream_t stream, const P& p, const Q& q) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p, q);
}
cudaError_t launch_swin_attn(const AttnParams& p, cudaStream_t stream);
cudaError_t launch_swin_mlp(const MlpParams& p, cudaStream_t stream);
cudaError_t launch_swin_attn_f16(const AttnParams& p, cudaStream_t stream);      // swin_kernels_f16.cu: fp16 operand images / weights
cudaError_t launch_swin_mlp_f16(const MlpParams& p, cudaStream_t stream);
cudaError_t launch_swin_mlp_f16h(const MlpParams& p, cudaStream_t stream);       // swin_kernels_f16h.cu: + GELU on packed halves
cudaError_t launch_token_linear(const LinearParams& p, cudaStream_t stream);
cudaError_t launch_winattn(int kind, const WinAttnParams& p, cudaStream_t stream);
int winattn_table_floats(int kind);
cudaError_t launch_layernorm(const float* x, float* y, ::__half* y16, const float* w, const float* b, int64_t num_tokens, int ld_in,
                             int ld_out, cudaStream_t stream);
cudaError_t launch_pixelshuffle_nhwc(const float* x, const float* bias, float* y, int batch, int height, int width, int out_channels,
                                     int r, cudaStream_t stream);
cudaError_t launch_bias_act_add(const float* x, const float* bias, const float* residual, float* y, int64_t pixels, int channels,
                                int act, float slope, cudaStream_t stream);
cudaError_t launch_stitch_accumulate(const float* tiles, float* E, float* Wt, const int32_t* tile_yx, int num_tiles,
                                     int channels, int tile_h, int tile_w, int out_h, int out_w, cudaStream_t stream);
int cab_ws_floats(int batch, int tokens_per_image);
cudaError_t launch_token_mean(const float* x, float* mean, float* sums, int batch, int tokens_per_image, cudaStream_t stream);
cudaError_t launch_token_mean_mlp(const float* x, float* out, float* sums, const float* w1, const float* b1, const float* w2, const float* b2,
                                  int hidden, int batch, int tokens_per_image, cudaStream_t stream);
int channel_gram_ws_floats(int batch, int tokens_per_image);
cudaError_t launch_cab_gate_add(const float* y, const float* y_bias, float* out, float* sums, const float* w1, const float* b1, const float* w2,
                                const float* b2, int hidden, float scale, int batch, int tokens_per_image, cudaStream_t stream);
cudaError_t launch_dwconv3x3_rows(const float* in, int ld_in, int c_in, const float* w, const float* scale, const float* shift,
                                  const float* stats, const float* gamma, const float* beta, const float* gate, int ld_gate, int c_gate,
                                  float* out, int ld_out, int C, int batch, int H, int W, int act_gelu, cudaStream_t stream, uint8_t* out_planes = nullptr, long long plane_stride = 0);
cudaError_t launch_row_stats(const float* in, int ld_in, int c_in, int C, int64_t tokens, float eps, float* stats, cudaStream_t stream);
cudaError_t launch_dat_mix(const float* att, const float* conv, const float* cmap, const float* w1, const float* b1, const float* w2,
                           float b2, int hidden, int mode, float* mix, int64_t tokens, int tokens_per_image, cudaStream_t stream);
cudaError_t launch_channel_gram(const float* qkv, float* gram, float* ws, int batch, int tokens_per_image, cudaStream_t stream);
cudaError_t launch_channel_apply(const float* qkv, const float* attn, float* out, int batch, int tokens_per_image, cudaStream_t stream);
cudaError_t launch_channel_softmax(const float* gram, const float* temperature, float* attn, int batch, cudaStream_t stream);
cudaError_t launch_stitch_normalize(float* E, const float* Wt, int channels, int64_t pixels, cudaStream_t stream);

// conv_kernel.cu
struct ConvArgs {
    const ::__half* in;          // fp16 NHWC (B, H, W, 64 * a_atoms)
    const uint8_t* wstream;      // packing.pack_conv3x3
    const float* bias;           // np floats
    const float* residual;       // fp32 rows like out_f32 (may alias it) or nullptr
    float* out_f32;
    ::__half* out_f16;
    int B, H, W, k_atoms, np, cout, out_mode, ld_out, act;
    float slope;
    int a_atoms;                 // 64-channel atoms of the input image (0: k_atoms); the last k_atoms - a_atoms k-steps re-read its last atoms
};
cudaError_t launch_conv3x3(const ConvArgs& a, cudaStream_t stream);
cudaError_t launch_rows_to_f16(const float* x, int ld_in, int C, ::__half* out, int cp, int64_t pixels, cudaStream_t stream);
cudaError_t launch_rows_to_f16_split(const float* x, int ld_in, int C, ::__half* hi, ::__half* lo, int ld_out, int cp, int64_t pixels,
                                     int act, float slope, int shuffle_h, int shuffle_w, cudaStream_t stream);
cudaError_t launch_image_to_f16_split(const float* x, int64_t sb, int64_t sc, int64_t sy, int64_t sx, int C, int B, int H, int W,
                                      const float* mean3, float range, ::__half* out, cudaStream_t stream);
cudaError_t launch_gather_tiles(const float* slab, int64_t slab_cstride, int slab_w, const int32_t* src_yx, int num_tiles, int channels,
                                int tile_h, int tile_w, float* out, cudaStream_t stream);
cudaError_t launch_stitch_accumulate2(const float* tiles, int64_t sn, int64_t sc, int64_t sy, int64_t sx, float* E, int64_t e_cstride,
                                      const int32_t* dst_yx, int num_tiles, int channels, int tile_h, int tile_w, int out_h, int out_w,
                                      cudaStream_t stream);
cudaError_t launch_stitch_finalize(const float* E, int64_t e_cstride, const float* cnt_y, const float* cnt_x, void* out, int64_t o_cstride,
                                   int dtype, int channels, int out_h, int out_w, cudaStream_t stream);

}  // namespace srk
