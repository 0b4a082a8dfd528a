/* srk.h -- C ABI of libsrk.so, the B200 (sm_100a) fused window-attention kernels.
 *
 * The reference (ViacheslavTimofeev/tpu_superresolution) has no FFI layer: its boundary for this
 * path is the nn.Module surface.  The drop-in modules in
 * tpu_superresolution_b200/swinir.py keep that surface and call the entry points below through
 * ctypes.  Each entry point names the reference code it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; every device pointer is BORROWED for the duration of the
 *    enqueue (PyTorch owns all memory); nothing is allocated or freed on the device;
 *  - functions only enqueue work on `stream` (a cudaStream_t), no implicit synchronisation;
 *  - return 0 on success, non-zero on error (srk_last_error_string() describes it);
 *    unsupported configurations are errors -- there is no CPU or eager fallback;
 *  - activations are fp32 token rows [token][ld] (ld floats per token, ld % 4 == 0, ld >= 180),
 *    i.e. exactly the reference's (B, H*W, C) tensors; GEMM operands are bf16 with fp32
 *    accumulation; weights arrive pre-packed by tpu_superresolution_b200/packing.py.
 */
#ifndef SRK_H_
#define SRK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRK_ABI_VERSION 3

/* Fixed geometry of the SwinIR/HAT/DAT "M" family served by these kernels. */
#define SRK_DIM 180        /* embed_dim                      (finetune_swinir.py:276) */
#define SRK_DIM_PAD 192    /* padded to a multiple of the UMMA K-atom (64) */
#define SRK_HEADS 6        /* num_heads                      (finetune_swinir.py:277) */
#define SRK_HEAD_DIM 30    /* 180 / 6 */
#define SRK_HEAD_PAD 32    /* head_dim padded to the UMMA K step */
#define SRK_WINDOW 8       /* window_size                    (finetune_swinir.py:273) */
#define SRK_HIDDEN 360     /* int(dim * mlp_ratio), mlp_ratio 2 (finetune_swinir.py:278) */
#define SRK_HIDDEN_PAD 384

/* byte sizes of the packed weight streams (see packing.py for the slab order) */
#define SRK_ATTN_WSTREAM_BYTES (3 * 24576 + 9 * 16384 + 3 * 24576)
#define SRK_MLP_WSTREAM_BYTES (9 * 16384 + 6 * 24576)
/* float offsets inside the packed per-block vectors.  LayerNorm's affine, the k bias and the v bias do not appear:
 * packing.py folds gamma/beta into the following GEMM, drops the k bias (softmax-invariant) and moves the v bias
 * into the proj bias (softmax rows sum to one). */
#define SRK_AV_BIAS_Q 384          /* 192: q bias in padded head layout (h*32+d), pre-scaled by head_dim^-0.5 * log2(e) */
#define SRK_AV_BIAS_PROJ 1024      /* 192 */
#define SRK_AV_RPB 1216            /* 6 x 232: relative-position-bias table per head, * log2(e) */
#define SRK_AV_RPB_STRIDE 232
#define SRK_ATTN_VEC_FLOATS (1216 + 6 * 232)
#define SRK_MV_B1 384              /* 384 */
#define SRK_MV_B2 768              /* 192 */
#define SRK_MLP_VEC_FLOATS 960

/* Attention half of a Swin block:  y = x + proj(softmax(q k^T * scale + rpb + mask) v),
 * q,k,v = qkv(LN1(x)) on cyclically shifted 8x8 windows.
 * Replaces network_swinir.py:244-276 (SwinTransformerBlock.forward up to the first residual) and,
 * with mode = SRK_MODE_WINDOWS, network_swinir.py:114-145 (WindowAttention.forward). */
enum { SRK_MODE_IMAGE = 0, SRK_MODE_WINDOWS = 1 };
enum { SRK_MASK_NONE = 0, SRK_MASK_SHIFT = 1, SRK_MASK_EXPLICIT = 2 };
/* GEMM operand type of the fused Swin kernels.  BF16: the default path (gate: max abs <= 2e-3 on [0,1] pixels, measured 1.6 - 4.4e-4).
 * F16: the "tight" mode -- operand images and packed weights in fp16, which has TF32's 11-bit significand (gate <= 2e-4); every
 * GEMM input of these kernels is LayerNorm output, a softmax probability or a projection of those, so fp16's range is safe.
 * The weight stream must have been packed for the same type (packing.py: operands=...). */
enum { SRK_OPERANDS_BF16 = 0, SRK_OPERANDS_F16 = 1,
       SRK_OPERANDS_F16_HALF_GELU = 2 /* srk_swin_mlp_fwd only: fp16 operands (weights packed as for F16) and the GELU evaluated on packed
                                         halves -- faster and still closer to the exact GELU than the bf16 variant; the modules' default MLP */ };

typedef struct SrkSwinAttnDesc {
    int32_t mode;          /* SRK_MODE_IMAGE: x is (batch, height*width, ld); SRK_MODE_WINDOWS: x is (num_windows, 64, ld) */
    int32_t batch;         /* images (mode IMAGE) */
    int32_t height, width; /* feature-map size, multiples of 8 (mode IMAGE) */
    int32_t num_windows;   /* B_ = nW*B (mode WINDOWS) */
    int32_t ld_in, ld_out; /* floats per token row of x and y */
    int32_t shift;         /* 0 or 4: cyclic shift, torch.roll(-shift,-shift) (network_swinir.py:249-250) */
    int32_t apply_ln;      /* 1: LayerNorm(x) first (network_swinir.py:245) */
    int32_t add_residual;  /* 1: y = x + attn (network_swinir.py:276); 0: y = attn */
    int32_t mask_mode;     /* SRK_MASK_SHIFT: closed form of calculate_mask (network_swinir.py:216-237) */
    int32_t mask_nw;       /* SRK_MASK_EXPLICIT: mask is (mask_nw, 64, 64) fp32, window w uses mask[w % mask_nw] */
    int32_t operands;      /* SRK_OPERANDS_BF16 / SRK_OPERANDS_F16 */
} SrkSwinAttnDesc;

int srk_swin_attn_fwd(const SrkSwinAttnDesc* desc, const float* x, float* y, const void* wstream /* SRK_ATTN_WSTREAM_BYTES */,
                      const float* vec /* SRK_ATTN_VEC_FLOATS */, const float* mask /* or NULL */, void* stream);

/* MLP half:  y = x + fc2(gelu(fc1(LN2(x)))), gelu = nn.GELU() (erf form) evaluated by a tanh-form polynomial fit, |error| <= 2.6e-5 before the bf16 rounding (rowops.cuh: gelu_fast).  Replaces network_swinir.py:277 + Mlp.forward :24-30. */
typedef struct SrkMlpDesc {
    int64_t num_tokens;
    int32_t ld_in, ld_out;
    int32_t apply_ln;      /* 1: LayerNorm first */
    int32_t add_residual;  /* 1: y = x + mlp(LN(x)); 0: y = mlp(LN(x)) */
    int32_t operands;      /* SRK_OPERANDS_BF16 / SRK_OPERANDS_F16 */
} SrkMlpDesc;

int srk_swin_mlp_fwd(const SrkMlpDesc* desc, const float* x, float* y, const void* wstream /* SRK_MLP_WSTREAM_BYTES */,
                     const float* vec /* SRK_MLP_VEC_FLOATS */, void* stream);

/* Optional fine-grained ordering of the two halves of consecutive Swin blocks that work in place on one residual stream
 * (a BasicLayer, network_swinir.py:349-416, SRK_MODE_IMAGE).  The kernels are launched with programmatic dependent launch;
 * with a progress array a kernel does not wait for the whole previous grid but, tile by tile, for the image the tile belongs
 * to -- the next kernel's first (cold) tiles then run on the SMs that the previous kernel's uneven last round leaves idle.
 *   progress: int32[2 * batch], zeroed by the caller (stream-ordered) before the first block of the group;
 *             [0, batch): windows finished by srk_swin_attn_fwd_sync launches, [batch, 2 batch): 128-token tiles finished by
 *             srk_swin_mlp_fwd_sync launches, per image (both accumulate over the blocks of the group).
 *   wait_target: value the OTHER half's counter of a tile's image must have reached before the tile is read: block k
 *             (0-based) of the group passes k * tokens_per_image / 128 to the attention half and (k + 1) * windows_per_image
 *             to the MLP half; 0 = order after all preceding work in the stream (first launch of a group, or no counters).
 * sync == NULL (or the plain entry points above): whole-grid ordering. */
typedef struct SrkBlockSync {
    int32_t* progress;
    int32_t batch;
    int32_t tokens_per_image;
    int32_t wait_target;
} SrkBlockSync;
int srk_swin_attn_fwd_sync(const SrkSwinAttnDesc* desc, const float* x, float* y, const void* wstream, const float* vec,
                           const float* mask /* or NULL */, const SrkBlockSync* sync /* or NULL */, void* stream);
int srk_swin_mlp_fwd_sync(const SrkMlpDesc* desc, const float* x, float* y, const void* wstream, const float* vec,
                          const SrkBlockSync* sync /* or NULL */, void* stream);

/* All blocks of one BasicLayer (network_swinir.py:349-416: `depth` SwinTransformerBlocks alternating shift 0 / 4) in ONE persistent
 * launch: y = blocks(y) in place.  The attention and MLP halves of every block are work items walked in global order by one CTA
 * per SM and ordered per image through `progress` (2 * batch int32 counters, zeroed by this call), see swin_layer_kernel.
 * Needs height * width %% 128 == 0 (and multiples of 8); n_blocks <= SRK_LAYER_MAX_BLOCKS; masks in closed form only. */
#define SRK_LAYER_MAX_BLOCKS 8
typedef struct SrkLayerBlock {
    const void* attn_wstream;   /* SRK_ATTN_WSTREAM_BYTES (LayerNorm 1 folded in) */
    const float* attn_vec;      /* SRK_ATTN_VEC_FLOATS */
    const void* mlp_wstream;    /* SRK_MLP_WSTREAM_BYTES (LayerNorm 2 folded in) */
    const float* mlp_vec;       /* SRK_MLP_VEC_FLOATS */
    int32_t shift;              /* 0 or 4 */
    int32_t reserved;
} SrkLayerBlock;
typedef struct SrkLayerDesc {
    int32_t batch, height, width;
    int32_t ld;                 /* floats per token row of y */
    int32_t n_blocks;
} SrkLayerDesc;
int srk_swin_layer_fwd(const SrkLayerDesc* desc, float* y, const SrkLayerBlock* blocks /* host array */, int32_t* progress, void* stream);

/* Token-wise linear layer on tcgen05 for the HAT / DAT paths:  out[tok, :] = act(A[tok, :] W^T + b), N in chunks of 192
 * columns.  Replaces the nn.Linear call sites hat_arch.py:179, :195, :401, :436 and dat_arch.py:371, :435, :483, :526, :79-88.
 *   A  SRK_LIN_A_ROWS  : fp32 token rows [tok][ld_in] (180 valid), optional LayerNorm (affine folded into W, b at pack time)
 *      SRK_LIN_A_PLANES: bf16 planes [k_atoms][num_tokens][64] (128 B per token row, 16-byte chunks permuted by chunk ^ (tok & 7))
 *   out SRK_LIN_OUT_PLANES: bf16 planes [3 * n_chunks][num_tokens][64]; plane p is permuted with (tok + 4) & 7 when bit p of
 *                          plane_phase_mask is set (operand-row phase of the consumer's window images)
 *       SRK_LIN_OUT_ROWS  : fp32 rows [tok][ld_out]; chunk c writes columns [180 c, 180 c + 180) (so a 720-wide hidden layer is
 *                          4 chunks); added into `out` when add_residual. */
enum { SRK_LIN_A_ROWS = 0, SRK_LIN_A_PLANES = 1 };
enum { SRK_LIN_OUT_PLANES = 0, SRK_LIN_OUT_ROWS = 1 };
enum { SRK_LIN_ACT_NONE = 0, SRK_LIN_ACT_GELU = 1 };
#define SRK_LIN_SLAB_BYTES 24576   /* one weight slab: 192 output rows x 64 k, bf16, 128-byte swizzle */
#define SRK_LIN_MAX_CHUNKS 8

typedef struct SrkLinearDesc {
    int64_t num_tokens;
    int32_t a_mode;        /* SRK_LIN_A_* */
    int32_t k_atoms;       /* K / 64: 3 or 6 */
    int32_t ld_in;         /* floats per token row (A_ROWS) */
    int32_t apply_ln;      /* A_ROWS: 1 = LayerNorm first */
    int32_t n_chunks;      /* N / 192 */
    int32_t act;           /* SRK_LIN_ACT_* */
    int32_t out_mode;      /* SRK_LIN_OUT_* */
    int32_t ld_out;        /* floats per output row (OUT_ROWS) */
    int32_t add_residual;  /* OUT_ROWS: 1 = out += result */
    uint32_t plane_phase_mask;
} SrkLinearDesc;

int srk_linear_fwd(const SrkLinearDesc* desc, const void* a, const void* wstream /* n_chunks * k_atoms * SRK_LIN_SLAB_BYTES */,
                   const float* bias /* n_chunks * 192 */, void* out, void* stream);

/* Window attention on q/k/v planes: softmax(q k^T + bias[+ mask]) v for every 256-query window and head.
 *   SRK_WA_HAT_WMSA  hat_arch.py:166-197 inside HAB (:281-301): 16x16 windows, cyclic shift (0 or 8), closed-form shift mask
 *                    (:921-940) or an explicit (nW, 256, 256) mask, relative position bias (961-entry table)
 *   SRK_WA_HAT_OCAB  hat_arch.py:403-432: 16x16 queries against the zero-padded 24x24 overlapping key window (nn.Unfold :378)
 *   SRK_WA_DAT_8x32 / SRK_WA_DAT_32x8  dat_arch.py:193-244: rectangular windows, dynamic position bias, shift (4,16) / (16,4)
 * q/k/v: plane p = heads 2p, 2p+1 (32 padded dims each), [p][batch*height*width][64] bf16, swizzled as written by
 * srk_linear_fwd.  bias_table: [2 * ceil(n_heads / 2)][srk_window_attention_table_floats(kind)] floats, already * log2(e),
 * entry (dy * SY + dx) for query-key offset (dy, dx) (see packing.py).  q must already carry scale * log2(e).
 * The softmax row sum is taken from the P v GEMM: padded dim 30 of every v head must be 1 (packing.py sets that bias).
 * pad_pages: 4096 zero bytes (k rows of OCAB's zero padding) followed by 32 v padding rows of 128 B (zeros except bf16 1.0
 * at dim 30 of both heads, chunk-swizzled with row & 7) -- packing.make_pad_pages().
 * out_mode 0: bf16 planes [ceil(n_heads/2)][tokens][64] swizzled with tok & 7 (A operand of the proj srk_linear_fwd);
 * out_mode 1: fp32 rows, out[tok * out_ld + out_col0 + head * 30 + d]. */
enum { SRK_WA_HAT_WMSA = 0, SRK_WA_HAT_OCAB = 1, SRK_WA_DAT_8x32 = 2, SRK_WA_DAT_32x8 = 3 };

typedef struct SrkWinAttnDesc {
    int32_t kind;
    int32_t batch, height, width;
    int32_t shift_y, shift_x;   /* cyclic shift of the window grid (torch.roll(-shift)) */
    int32_t mask_shift;         /* 1: closed-form shifted-window mask */
    int32_t n_heads;            /* heads in the planes (6; 3 per DAT branch) */
    int32_t emask_nw;           /* explicit mask: number of windows in the mask tensor (0 = none) */
    int32_t out_mode, out_ld, out_col0;
} SrkWinAttnDesc;

int srk_window_attention_fwd(const SrkWinAttnDesc* desc, const void* q_planes, const void* k_planes, const void* v_planes,
                             const float* bias_table, const float* emask, const void* pad_pages /* 8192 bytes, see below */,
                             void* out, void* stream);
int srk_window_attention_table_floats(int32_t kind);

/* Row LayerNorm over the 180 channels of token rows (patch_embed.norm / final norm,
 * network_swinir.py:526-527, :800). */
int srk_layernorm_fwd(const float* x, float* y, const float* w, const float* b, int64_t num_tokens, int32_t ld_in,
                      int32_t ld_out, void* stream);
/* The same with the result (also) written as fp16 NHWC rows of SRK_DIM_PAD channels, zero padded: the input layout of
 * srk_conv3x3_fwd (norm1 in front of HAT's CAB, hat_arch.py:276-278; the final norm in front of conv_after_body,
 * network_swinir.py:800, :829).  y may be NULL (fp16 only). */
int srk_layernorm_f16_fwd(const float* x, float* y, void* y16, const float* w, const float* b, int64_t num_tokens, int32_t ld_in,
                          int32_t ld_out, void* stream);

/* HAT CAB tail: out[b, t, c] += scale * y[b, t, c] * sigmoid(W2 relu(W1 mean_t(y[b, :, c]) + b1) + b2)[c] on channels-last
 * (batch, tokens_per_image, 180) fp32 tensors -- ChannelAttention (hat_arch.py:41-59) fused with `+ conv_x * conv_scale`
 * (hat_arch.py:307).  w1 (hidden, 180), w2 (180, hidden), hidden <= 32; sums_ws: srk_cab_ws_floats() floats of scratch (per-chunk
 * partial sums, reduced in a fixed order: results are run-to-run identical). */
int srk_cab_ws_floats(int32_t batch, int32_t tokens_per_image);
/* mean[b, c] = mean over the tokens of x[b, :, c] for channels-last (batch, tokens_per_image, 180) fp32 rows: the
 * AdaptiveAvgPool2d(1) in front of DAT's channel interaction (dat_arch.py:305-310).  sums_ws: srk_cab_ws_floats() floats
 * (per-chunk partial sums, reduced in a fixed order). */
int srk_token_mean_fwd(const float* x, float* mean, float* sums_ws, int32_t batch, int32_t tokens_per_image, void* stream);
/* ... and the rest of that channel interaction in the same call: out[b] = w2 gelu(w1 mean[b] + b1) + b2 (the two 1x1 convolutions of
 * dat_arch.py:305-310 with the eval BatchNorm folded into w1 / b1, exact GELU).  w1 (hidden, 180), w2 (180, hidden), hidden <= 64. */
int srk_token_mean_mlp_fwd(const float* x, float* out, float* sums_ws, const float* w1, const float* b1, const float* w2, const float* b2,
                           int32_t hidden, int32_t batch, int32_t tokens_per_image, void* stream);
int srk_cab_gate_add(const float* y, const float* y_bias /* bias of the conv that produced y, or NULL: y + y_bias is used */, float* out, float* sums_ws, const float* w1, const float* b1, const float* w2, const float* b2,
                     int32_t hidden, float scale, int32_t batch, int32_t tokens_per_image, void* stream);

/* ---- DAT glue on fp32 channels-last token rows (dat_arch.py) ------------------------------------------------------------
 * Depthwise 3x3 convolution (zero padding) over (batch, height, width) token rows: channel slice [c_in, c_in + channels) of rows
 * of ld_in floats, weights w9c[9][channels] (tap major), out = act((conv) * scale + shift) [* gate slice].  With ln_stats the
 * input rows are LayerNorm-ed first ((x - mean) * rstd * ln_gamma + ln_beta, stats from srk_row_stats_fwd; zero padding applies
 * to the normalised tensor).  Replaces dwconv + BatchNorm + GELU (dat_arch.py:300-304, :418, :508) and SpatialGate (:38-54). */
int srk_dwconv3x3_rows_fwd(const float* in, int32_t ld_in, int32_t c_in, const float* w9c, const float* scale, const float* shift,
                           const float* ln_stats, const float* ln_gamma, const float* ln_beta, const float* gate, int32_t ld_gate,
                           int32_t c_gate, float* out, int32_t ld_out, int32_t channels, int32_t batch, int32_t height, int32_t width,
                           int32_t act_gelu, void* stream);
/* The same with the result as bf16 planes [channel / 64][token][64 channels] (plane_stride bytes apart, 16-byte chunks permuted by
 * chunk ^ (token & 7)): the SRK_LIN_A_PLANES input of srk_linear_fwd -- DAT's SpatialGate output feeds fc2 (K = 360 as six k-atoms)
 * in ONE launch this way.  Channels past `channels` in the last plane are not written: zero them once. */
int srk_dwconv3x3_rows_planes_fwd(const float* in, int32_t ld_in, int32_t c_in, const float* w9c, const float* scale, const float* shift,
                                  const float* ln_stats, const float* ln_gamma, const float* ln_beta, const float* gate, int32_t ld_gate,
                                  int32_t c_gate, void* out_planes, int64_t plane_stride, int32_t channels, int32_t batch, int32_t height,
                                  int32_t width, int32_t act_gelu, void* stream);
/* stats[tok] = (mean, rstd) over the channel slice (nn.LayerNorm statistics, biased variance). */
int srk_row_stats_fwd(const float* in, int32_t ld_in, int32_t c_in, int32_t channels, int64_t tokens, float eps, float* stats, void* stream);
/* Adaptive interaction module (dat_arch.py:420-433 mode 0, :510-523 mode 1) on (tokens, 180) rows: s = w2 . gelu(W1 src + b1) + b2
 * with src = att (mode 0) / conv (mode 1), W1 (hidden <= 16, 180) with its BatchNorm folded; cmap (batch, 180) = channel map
 * before the sigmoid.  mode 0: mix = att * sigmoid(cmap) + sigmoid(s) * conv;  mode 1: mix = att * sigmoid(s) + conv * sigmoid(cmap). */
int srk_dat_mix_fwd(const float* att, const float* conv, const float* cmap, const float* w1, const float* b1, const float* w2, float b2,
                    int32_t hidden, int32_t mode, float* mix, int64_t tokens, int32_t tokens_per_image, void* stream);
/* Channel attention (dat_arch.py:497-505) on qkv rows (batch * tokens_per_image, 540) = q | k | v:
 * gram[b][h] = 900 products sum_n q[n, d1] k[n, d2], then 30 sums of q^2, then 30 sums of k^2 (960 floats per image and head);
 * apply: out[tok, h*30 + d1] = sum_d2 attn[b, h, d1, d2] v[tok, h*30 + d2], attn (batch, 6, 30, 30). */
int srk_dat_channel_gram_fwd(const float* qkv, float* gram, float* ws /* srk_dat_channel_gram_ws_floats() */, int32_t batch,
                             int32_t tokens_per_image, void* stream);
int srk_dat_channel_gram_ws_floats(int32_t batch, int32_t tokens_per_image);
/* attn[b][h][d1][:] = softmax(gram[d1][:] / (max(sqrt(sum q_d1^2), 1e-12) max(sqrt(sum k_d2^2), 1e-12)) * temperature[h]): F.normalize
 * over the tokens, the temperature and the softmax of dat_arch.py:497-503 on the output of srk_dat_channel_gram_fwd; temperature (6). */
int srk_dat_channel_softmax_fwd(const float* gram, const float* temperature, float* attn, int32_t batch, void* stream);
int srk_dat_channel_apply_fwd(const float* qkv, const float* attn, float* out, int32_t batch, int32_t tokens_per_image, void* stream);

/* PixelShuffle(r) on channels-last activations, optional fused LeakyReLU:
 * out[b, h*r+i, w*r+j, c] = in[b, h, w, c*r*r + i*r + j].  Replaces nn.PixelShuffle in Upsample
 * (network_swinir.py:580-588). */
int srk_pixelshuffle_nhwc_fwd(const float* x, float* y, int32_t batch, int32_t height, int32_t width,
                              int32_t out_channels, int32_t r, void* stream);
/* The same with the preceding convolution's bias added on the way (bias[in_channel], may be NULL): the Upsample
 * stage is [conv, PixelShuffle] (network_swinir.py:580-588) and the conv is run without bias. */
int srk_pixelshuffle_nhwc_bias_fwd(const float* x, const float* bias, float* y, int32_t batch, int32_t height, int32_t width,
                                   int32_t out_channels, int32_t r, void* stream);

/* y[p, c] = act(x[p, c] + bias[c]) + residual[p, c] on channels-last activations (p < pixels, c < channels).
 * The tail of every conv of the reference networks: conv bias, then LeakyReLU (conv_before_upsample,
 * network_swinir.py:722-723) or the residual add (RSTB: `self.patch_embed(self.conv(...)) + x`, :501; the long skip,
 * :829).  bias / residual may be NULL; y may alias x or residual. */
#define SRK_ACT_NONE 0
#define SRK_ACT_LEAKY_RELU 1
#define SRK_ACT_GELU 2        /* exact (erf) GELU: nn.GELU() between the CAB convs, hat_arch.py:68 */
int srk_bias_act_add_nhwc(const float* x, const float* bias, const float* residual, float* y, int64_t pixels, int32_t channels,
                          int32_t act, float slope, void* stream);

/* Weighted tile accumulation for the overlapping-tile stitcher (BASELINE.json configs[4]):
 * E[:, y0:y0+th, x0:x0+tw] += tile, Wt[y0:.., x0:..] += 1  (tiles of one launch must not overlap). */
int srk_stitch_accumulate(const float* tiles, float* E, float* Wt, const int32_t* tile_yx, int32_t num_tiles,
                          int32_t channels, int32_t tile_h, int32_t tile_w, int32_t out_h, int32_t out_w, void* stream);
int srk_stitch_normalize(float* E, const float* Wt, int32_t channels, int64_t pixels, void* stream);

/* Tiler, second form (tiling.py; the reference has no tiling -- BASELINE.json configs[4]).
 * srk_gather_tiles: out (num_tiles, channels, tile_h, tile_w) <- windows of the LR band `slab` (channels, slab_rows, slab_w; channel
 *   stride slab_cstride floats) at src_yx[k] = (y0, x0): replaces the per-tile slicing + torch.stack of the host loop.
 * srk_stitch_accumulate_strided: E[c, y0+ty, x0+tx] += tiles[k, c, ty, tx] with explicit element strides of `tiles` (the
 *   models return channels-last memory) and a channel stride for E (a rank's band may live inside the full output image); the
 *   tiles of one call must be pairwise disjoint; rows / columns outside [0, out_h) x [0, out_w) are skipped (seam tiles).
 * srk_stitch_finalize: out = E / (cnt_y[y] * cnt_x[x]) -- the cover count of the E / W rule is separable, so no count plane is
 *   accumulated -- converted to out_dtype (uint8: round(clamp(v, 0, 1) * 255)); out may alias E for SRK_OUT_F32. */
#define SRK_OUT_F32 0
#define SRK_OUT_BF16 1
#define SRK_OUT_U8 2
int srk_gather_tiles(const float* slab, int64_t slab_cstride, int32_t slab_rows, int32_t slab_w, const int32_t* src_yx, int32_t num_tiles,
                     int32_t channels, int32_t tile_h, int32_t tile_w, float* out, void* stream);
int srk_stitch_accumulate_strided(const float* tiles, int64_t stride_n, int64_t stride_c, int64_t stride_y, int64_t stride_x, float* E,
                                  int64_t e_channel_stride, const int32_t* dst_yx, int32_t num_tiles, int32_t channels, int32_t tile_h,
                                  int32_t tile_w, int32_t out_h, int32_t out_w, void* stream);
int srk_stitch_finalize(const float* E, int64_t e_channel_stride, const float* cnt_y, const float* cnt_x, void* out, int64_t out_channel_stride,
                        int32_t out_dtype, int32_t channels, int32_t out_h, int32_t out_w, void* stream);

/* 3x3 convolution, stride 1, zero padding 1, as a tcgen05 implicit GEMM with the conv's tail fused (csrc/conv_kernel.cu).
 * Replaces F.conv2d / nn.Conv2d(…, 3, 1, 1) at network_swinir.py:465 (RSTB), :720 (conv_first), :729 (conv_after_body), :742-745
 * (conv_before_upsample, Upsample, conv_last), hat_arch.py:67-72 (CAB) and the same layers of hat_arch.py / dat_arch.py.
 *   in       fp16 NHWC (batch, height, width, 64 * a_atoms), channels zero-padded (written by srk_rows_to_f16 /
 *            srk_image_to_f16_split or by a previous srk_conv3x3_fwd in an fp16 output mode)
 *   wstream  packing.pack_conv3x3: k_atoms x 3 x 3 slabs of np rows x 128 B;  bias: np floats
 *   out      SRK_CONV_OUT_ROWS_F32:     fp32 (pixels, ld_out), out = act(conv + bias) [+ residual]; residual may alias out
 *            SRK_CONV_OUT_NHWC_F16:     fp16 (pixels, ld_out) with ld_out >= np (the next convolution's input)
 *            SRK_CONV_OUT_SHUFFLE2_F16: np = 256: fp16 (batch, 2 height, 2 width, 64) = nn.PixelShuffle(2) of the conv output
 *                                       (network_swinir.py:584-585; pack_conv3x3(pixel_shuffle=True) permutes the weight rows)
 *            SRK_CONV_OUT_IMAGE:        cout <= 4, np = 16: fp32 (pixels, ld_out = cout)
 * fp16 operands (11-bit significand, as the TF32 library path had), fp32 accumulation. */
#define SRK_CONV_OUT_ROWS_F32 0
#define SRK_CONV_OUT_NHWC_F16 1
#define SRK_CONV_OUT_SHUFFLE2_F16 2
#define SRK_CONV_OUT_IMAGE 3
typedef struct SrkConvDesc {
    int32_t batch, height, width;
    int32_t k_atoms;        /* k-steps of 64 channels: 1..12; normally = a_atoms = padded input channels / 64 */
    int32_t np;             /* padded output channels: multiple of 32 (16 for SRK_CONV_OUT_IMAGE), <= 256 */
    int32_t cout;           /* real output channels (multiple of 4 for SRK_CONV_OUT_ROWS_F32) */
    int32_t out_mode;
    int32_t ld_out;
    int32_t act;            /* SRK_ACT_* */
    float slope;
    int32_t a_atoms;        /* 64-channel atoms of the input image, 1..8 (0: = k_atoms).  With k_atoms > a_atoms the last k_atoms - a_atoms
                             * k-steps re-read the image's last atoms: the tight mode's convolution is ONE launch over the image
                             * [lo(x) | hi(x)] (a_atoms = 2 C/64) against the weights [hi(w) | lo(w) | hi(w)] (k_atoms = 3 C/64):
                             * lo*hi + hi*lo + hi*hi, small terms first, see srk_rows_to_f16_split */
} SrkConvDesc;
int srk_conv3x3_fwd(const SrkConvDesc* desc, const void* in_f16, const void* wstream, const float* bias, const float* residual, void* out,
                    void* stream);
/* fp32 token rows (pixels, ld_in) with `channels` channels -> fp16 NHWC (pixels, cp), cp = 64 * k_atoms, zero padded. */
int srk_rows_to_f16(const float* x, int32_t ld_in, int32_t channels, void* out_f16, int32_t cp, int64_t pixels, void* stream);
/* Tight mode (fp32-class convolutions out of the fp16 tensor-core kernel): fp32 rows -> the fp16 pair hi = fp16(v), lo = fp16(v - hi),
 * v = act(x) (22 bits together), as two NHWC images with row pitch ld_out and cp = 64 * ceil(C / 64) zero-padded channels each.
 * conv(x, w) = lo * hi(w) + hi * lo(w) + hi * hi(w), accumulated in fp32 in that order, is ONE srk_conv3x3_fwd launch over the
 * image [lo | hi] (lo = base, hi = base + cp, ld_out = 2 cp; SrkConvDesc.a_atoms = 2 cp / 64, k_atoms = 3 cp / 64) against weights
 * packed [hi(w) | lo(w) | hi(w)] (packing.pack_conv3x3(split=True)); the dropped lo * lo term is 2^-22 relative.
 * shuffle_h, shuffle_w > 0: x holds the 4 x 64 output channels of a conv + nn.PixelShuffle(2) stage at shuffle_h x shuffle_w pixels
 * per image (channels == 256, cp == 64; network_swinir.py:584-585); the pair is written at the 2x resolution. */
int srk_rows_to_f16_split(const float* x, int32_t ld_in, int32_t channels, void* hi_f16, void* lo_f16, int32_t ld_out, int32_t cp,
                          int64_t pixels, int32_t act, float slope, int32_t shuffle_h, int32_t shuffle_w, void* stream);
/* network input (batch, channels <= 3, height, width) fp32 with element strides (sb, sc, sy, sx) -> fp16 NHWC (pixels, 64) holding
 * [hi(v), v - hi(v), hi(v)] of v = (x - mean[c]) * range (network_swinir.py:803-804): conv_first's input with the fp16 rounding
 * compensated (pack_conv3x3(split_first=True) packs [hi(w), hi(w), w - hi(w)]).  mean3: 3 floats in HOST memory. */
int srk_image_to_f16_split(const float* x, int64_t sb, int64_t sc, int64_t sy, int64_t sx, int32_t channels, int32_t batch, int32_t height,
                           int32_t width, const float* mean3, float range, void* out_f16, void* stream);

int srk_abi_version(void);
const char* srk_last_error_string(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t srk_launch_count(void);
/* debug: device buffer of >= 2048 uint64 receiving CTA 0's clock64() timeline of later launches (NULL = off) */
void srk_debug_set_timeline(void* device_buf);
/* tuning: start skew (cycles per CTA index mod 4) of the attention / MLP kernels, see stagger_start() */
void srk_debug_set_stagger(int attn_cycles, int mlp_cycles);
/* tuning: cycles the second query-half group of srk_window_attention_fwd starts behind the first (default 1500) */
void srk_debug_set_winattn_stagger(int cycles);
/* tuning: 0 disables programmatic dependent launch of the fused Swin kernels (default 1) */
void srk_debug_set_pdl(int enabled);

#ifdef __cplusplus
}
#endif
#endif /* SRK_H_ */
