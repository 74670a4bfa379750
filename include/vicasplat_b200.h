/*
 * vicasplat_b200 -- C-ABI of the B200-native (sm_100a) hot path of WU-CVGL/VicaSplat.
 *
 * Every entry point takes plain device pointers, sizes and a cudaStream_t.  The library never
 * allocates or frees device memory, keeps no mutable global state (tensor maps are built per
 * call on the host stack) and never synchronises the stream.  All functions return 0 on success
 * or a negative VS_ERR_* code; vs_last_error() gives a thread-local message for the last failure.
 *
 * Reference interfaces replaced (paths relative to the VicaSplat repository):
 *   curope.rope_2d                     src/model/encoder/backbone/croco/curope/curope.cpp:49-69
 *                                      src/model/encoder/backbone/croco/curope/kernels.cu:18-108
 *   diff_gaussian_rasterization        call site src/model/decoder/cuda_splatting.py:207-235
 *   DecoderSplattingCUDA.forward       src/model/decoder/decoder_splatting_cuda.py:38-101
 *   VicaSplat.forward (torch ops)      src/model/encoder/vicasplat.py:158-278
 *     nn.Linear / Conv2d               -> vs_gemm (tcgen05 implicit GEMM)
 *     nn.LayerNorm + AdaLN modulate    -> vs_layernorm        backbone_vica.py:268-335
 *     softmax(QK^T)V / SDPA            -> vs_attention        croco/blocks.py:105-109, backbone_vica.py:116-121,188
 *     MyGaussianAdapter.forward        -> vs_gaussian_adapter common/gaussian_adapter.py:167-212
 */
#ifndef VICASPLAT_B200_H_
#define VICASPLAT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* vs_stream_t; /* == cudaStream_t */

enum {
  VS_OK = 0,
  VS_ERR_INVALID = -1,    /* bad argument (mirrors TORCH_CHECK failures of the reference ops) */
  VS_ERR_CUDA = -2,       /* a CUDA runtime / driver call failed */
  VS_ERR_WORKSPACE = -3,  /* caller-provided workspace too small */
  VS_ERR_UNSUPPORTED = -4
};

enum { VS_F32 = 0, VS_BF16 = 1, VS_F16 = 2, VS_F64 = 3 };
enum { VS_ACT_NONE = 0, VS_ACT_GELU = 1, VS_ACT_RELU = 2 };

const char* vs_last_error(void);
int vs_version(void);
/* sizeof() of a parameter struct by its C name (e.g. "vs_gemm_params"), -1 if unknown: lets a
 * foreign-function binding verify its mirror of the struct layout at load time. */
int64_t vs_struct_size(const char* name);
/* number of kernels this library has launched (or recorded into a CUDA graph being captured) in
 * this process so far -- instrumentation for bench.py's gpu_launches. */
int64_t vs_launch_count(void);

/* ------------------------------------------------------------------ RoPE-2D (curope.rope_2d)
 * In-place 2-D rotary embedding on tokens (B, N, H, D), D % 4 == 0, last dim contiguous,
 * stride(2) == D.  positions (B, N, 2) int64 contiguous (y, x).  fwd = +F0 forward, -F0 backward.
 * curope.cpp:49-69 / kernels.cu:18-108.  dtype: VS_F32, VS_F16 or VS_BF16. */
int vs_rope_2d(void* tokens, int dtype, int B, int N, int H, int D, int64_t stride_b,
               int64_t stride_n, const int64_t* positions, float base, float fwd,
               vs_stream_t stream);

/* Fused row-wise rope used by the encoder path: bf16 rows of a packed qkv buffer [rows, ld],
 * q at column q_col, k at column k_col, H heads of 64.  pos (rows, 2) int32: (y, x) >= 0 selects
 * the 2-D image rope (base 100); y < 0 selects the temporal camera rope with frame index
 * t = -1 - y, interleaved pairs, base cam_theta (src/misc/rope_utils.py:133-137,297-305). */
int vs_rope_rows(void* qkv, int64_t ld, int rows, int H, int q_col, int k_col, const int32_t* pos,
                 float base, float cam_theta, vs_stream_t stream);

/* ------------------------------------------------------------------ GEMM / implicit-GEMM conv
 * C[m, n] = epilogue( sum_k A[m, k] * W[n, k] ), bf16 operands, fp32 accumulation in TMEM.
 *   epilogue: v = acc + bias[n]; v = act(v); v *= (1 + gate[m / gate_rows, n]); v += res1 + res2
 * a_mode 0 (rows): A is `a_groups` groups of `a_rows` valid rows (row stride a_row_stride,
 *   group stride a_group_stride, in elements); logical row m = g * a_rows + r.
 * a_mode 1 (conv): A is NHWC bf16 [cn, ch, cw, cin]; stride-1 kh x kw convolution with zero
 *   padding `pad`; logical row m = (n * ch + y) * cw + x; W is [N, kh*kw*cin_pad] with
 *   cin_pad = round_up(cin, 64) (tap-major, channel-minor).  The output map is always ch x cw.
 *   conv_in_h > 0: the input map has conv_in_h rows (a "valid" convolution in y: pad rows are
 *   stored, e.g. conv_in_h = ch + kh - 1 with pad = 0).  conv_stride_{x,y,n} > 0 override the
 *   dense NHWC element strides of the input view -- overlapping views are allowed (stride_x <
 *   cin turns cin into a sliding window over pixels: the 7x7 stem of dpt_gs_head.py:113-118 is run
 *   as kh = 7, kw = 1 over windows of 8 pixels x 8 padded channels).
 * a_mode 2 (tn): both operands are given K-row-wise ("MN-major"): A is (K, a_rows) with row stride
 *   a_row_stride, W is (K, N) with row stride w_row_stride, C[m, n] = sum_k A[k, m] * W[k, n].  This is
 *   the weight-gradient form dW = dY^T X with dY (tokens, N_out) and X (tokens, K_in) used as they
 *   are stored -- no transposed copies (the tensor core reads MN-major shared-memory tiles).
 * a_mode 3 (conv wgrad): the weight gradient of the a_mode 1 convolution, with the pixels as the
 *   contraction dimension and no im2col buffer: A = dY, an NHWC bf16 map [cn, ch, cw, a_rows] (pixel
 *   stride a_row_stride elements; a_rows = Cout); W = X, the convolution's INPUT, described by
 *   cn/ch/cw/cin/kh/kw/pad/conv_in_h/conv_stride_* exactly as a_mode 1 describes its A operand;
 *   C[o, tap * cin_pad + c] = sum over pixels of dY[pixel, o] * X[pixel + tap, c]  (N = kh*kw*cin_pad,
 *   i.e. the packed weight layout of a_mode 1).  Each 64-column block of a W tile is one 4-D TMA box
 *   of X shifted by its filter tap (out-of-bounds = the zero padding).
 * c_accumulate != 0: C (fp32) is ACCUMULATED with vectorised atomic adds (red.global.add.v4.f32)
 *   instead of written: "dW +=" of the weight-gradient GEMMs.  It also enables split-K: the K range is
 *   cut into split_k parts (0 = chosen so that every SM has work; the wgrad GEMMs have few output
 *   tiles and a long K) that run as independent tiles.  Needs a plain epilogue (no bias / act / gate
 *   / residual).
 * mask_mode: 1 = res2 (bf16) is a ReLU mask: v = res2 > 0 ? v : 0, THEN v += res1; 2 = res1 (bf16) is
 *   the mask and nothing is added (backward of a ReLU whose output was kept: dpt_block.py:129-135).
 *   res_up2 != 0: res1 is an NHWC bf16 map at HALF resolution [cn, ch/2, cw/2, N] that is
 *   bilinearly upsampled x2 (align_corners=True, dpt_block.py:214-216) on the fly.
 * Output row mapping: out_row = (m / out_gin) * out_gout + out_off + (m % out_gin).
 * rope_pos != NULL: the rotary embedding of vs_rope_rows is applied to the fp32 accumulator (after
 *   bias) of the q columns [rope_q_col, rope_q_col + 64*rope_heads) and the k columns
 *   [rope_k_col, ...) before the store -- same (y, x) / camera-row convention, positions indexed
 *   by OUTPUT row; both column offsets must be multiples of 64 (replaces the separate in-place
 *   pass over the packed qkv buffer that the reference's cuRoPE2D makes, curope2d.py:12-44).
 */
typedef struct vs_gemm_params {
  const void* A;
  int32_t a_mode;
  int32_t a_rows, a_groups;
  int64_t a_row_stride, a_group_stride;
  int32_t cn, ch, cw, cin, kh, kw, pad;
  int32_t conv_in_h;
  int64_t conv_stride_x, conv_stride_y, conv_stride_n;
  const void* W;
  int64_t w_row_stride;
  int32_t N, K; /* K ignored in conv mode */
  const float* bias;
  int32_t act;
  const float* gate;
  int64_t gate_ld;
  int32_t gate_rows;      /* rows (of the *output* row index) per gate group; 0 = no grouping */
  int32_t first_row_mode; /* for out_row % gate_rows == 0: 0 as others, 1 no gate, 2 skip row */
  const void* res1;
  const void* res2;
  int32_t res_dtype;
  int32_t res_up2;
  int64_t res_ld;
  void* C;
  int32_t c_dtype;
  int64_t ldc;
  void* C2; /* optional bf16 copy of relu(v) with leading dimension ldc2 */
  int64_t ldc2;
  int32_t out_gin, out_gout, out_off; /* out_gin == 0: identity mapping */
  int32_t block_n;                    /* 0 = choose; else 64, 128 or 256 */
  const int32_t* rope_pos;            /* (out_rows, 2) int32 or NULL */
  int32_t rope_q_col, rope_k_col, rope_heads;
  float rope_base, rope_cam_theta;
  int32_t mask_mode;    /* 0 none, 1 res2 masks then res1 is added, 2 res1 masks */
  int32_t c_accumulate; /* C += result (atomic, fp32 only) */
  int32_t split_k;      /* with c_accumulate: number of K ranges (0 = choose) */
  float out_scale;      /* with mask_mode != 0 only; 0 = 1: the accumulator is multiplied by this first (e.g. 1 / keep of a dropout) */
  int32_t operand_dtype; /* 16-bit format of A, W and of 16-bit C / C2 / residual maps: 0 or VS_BF16 = bf16 (speed
                            mode), VS_F16 = fp16 (parity mode: TF32's 10-bit mantissa, same tensor rate; forward
                            a_modes only) */
} vs_gemm_params;

int vs_gemm(const vs_gemm_params* p, vs_stream_t stream);

/* ------------------------------------------------------------------ LayerNorm (+AdaLN modulate)
 * y = LN(x) * w + b, then for rows that are not "first rows": y = y * (1 + scale[f]) + shift[f]
 * with f = row / rows_per_frame.  Rows with row % rows_per_frame == 0 use (w0, b0) instead of
 * (w, b) and are never modulated when w0 != NULL (camera tokens, backbone_vica.py:283-316).
 * x fp32 [rows, C] (ldx); outputs: y_bf16 (nullable), y_f32 (nullable). eps as given (1e-6). */
typedef struct vs_layernorm_params {
  const float* x;
  int64_t ldx;
  int32_t rows, C;
  const float *w, *b, *w0, *b0;
  const float *scale, *shift;
  int64_t mod_ld;
  int32_t rows_per_frame;
  float eps;
  int32_t normalize; /* 0: skip the normalisation (plain convert / modulate) */
  void* y_bf16;
  int64_t ldy_bf16;
  float* y_f32;
  int64_t ldy_f32;
  int32_t y16_dtype; /* format of y_bf16: 0 / VS_BF16 = bf16, VS_F16 = fp16 */
} vs_layernorm_params;
int vs_layernorm(const vs_layernorm_params* p, vs_stream_t stream);

/* ------------------------------------------------------------------ attention (hd = 64)
 * O = softmax(Q K^T * scale) V per (item, head); flash-style, S and O tiles in TMEM (tcgen05).
 * Q/K/V are bf16 row matrices with explicit leading dimensions (so they may alias a packed qkv
 * buffer); head h occupies columns [h*64, h*64+64) of each.  q_rows / kv_rows are the total row
 * counts of the Q and K/V buffers (TMA bounds).  Item i: queries = rows [q_start[i], q_start[i] +
 * q_len[i]) of Q, keys = concatenation of up to two row segments of K/V (kv_start0/len0,
 * kv_start1/len1).  causal_block > 0: a query whose absolute row r has r % causal_block == 0 only
 * sees keys with absolute row index < (r / causal_block + 1) * causal_block (camera-token
 * blocked-causal mask, backbone_vica.py:585-593); other rows see every key of the item.
 * Replaces croco/blocks.py:105-109 and F.scaled_dot_product_attention at
 * backbone_vica.py:116-121,188. */
typedef struct vs_attention_params {
  const void *Q, *K, *V;
  void* O;
  int64_t ldq, ldk, ldv, ldo;
  int32_t q_rows, kv_rows;
  int32_t heads, items;
  const int32_t *q_start, *q_len, *kv_start0, *kv_len0, *kv_start1, *kv_len1; /* device arrays */
  int32_t max_q_len;
  int32_t max_kv_len;   /* upper bound of kv_len0 + kv_len1 over the items (0 = unknown) */
  int32_t causal_block;
  float scale;
  float* lse; /* optional output (q_rows, heads) f32 for vs_attention_backward: log2-domain
                 log-sum-exp of the scaled scores of every query row (+inf for a row without keys) */
  int32_t dtype; /* 0 / VS_BF16: Q, K, V, O (and the probabilities) are bf16; VS_F16: fp16 (parity mode) */
} vs_attention_params;
int vs_attention(const vs_attention_params* p, vs_stream_t stream);

/* ------------------------------------------------------------------ small fused ops
 * half_dtype (where present): the 16-bit format of the bf16-named operands -- 0 / VS_BF16 = bf16,
 * VS_F16 = fp16 (the parity mode of the encoder path, see vs_gemm_params.operand_dtype). */
/* Non-overlapping patch gather: img fp32 NCHW [n,3,h,w] -> bf16 [n*(h/P)*(w/P), 3*P*P]
 * (channel-major, then row, then column: the flattening order of Conv2d weight [E,3,P,P];
 * croco/blocks.py:195-225). */
int vs_patchify(const float* img, void* out, int n, int h, int w, int P, int half_dtype, vs_stream_t stream);
/* kxk stride-s im2col of an NHWC bf16 map (or NCHW fp32 image when src_nchw_f32 != 0) into
 * bf16 [n*ho*wo, kpad], column order (tap-major, channel-minor), zero padded to kpad. */
int vs_im2col(const void* src, int src_nchw_f32, void* out, int n, int h, int w, int c, int k,
              int stride, int pad, int kpad, int half_dtype, vs_stream_t stream);
/* bilinear x2, align_corners=True, NHWC bf16 (heads/dpt_block.py:214-216) */
int vs_upsample2x(const void* src, void* dst, int n, int h, int w, int c, int half_dtype, vs_stream_t stream);
/* dst = bilinear_x2(src) + add (add: full-resolution NHWC bf16 map; the image-feature merge of
 * dpt_gs_head.py:148-150 in its un-fused, training form) */
int vs_upsample2x_add(const void* src, const void* add, void* dst, int n, int h, int w, int c,
                      int half_dtype, vs_stream_t stream);
/* ConvTranspose2d with kernel == stride == k, expressed as GEMM output [n*h*w, k*k*c]
 * (column = (dy*k+dx)*c + co) scattered to NHWC [n, h*k, w*k, c] (bf16 -> bf16). */
int vs_pixel_shuffle(const void* src, void* dst, int n, int h, int w, int c, int k,
                     vs_stream_t stream);
/* intrinsic token: out[f, :] = Linear(9 -> E)(K[f].flatten()), written to x[f*rows_per_frame +
 * row_off] (fp32 residual stream; backbone_vica.py:535-536,455-459). */
int vs_intrinsic_token(const float* K9, const float* w, const float* b, float* x, int frames, int E,
                       int rows_per_frame, int row_off, vs_stream_t stream);
/* camera tokens of the decoder (backbone_vica.py:492-494): row f*rows_per_frame of x (fp32, C) =
 * intr_tok (+ extr_tok when f % T != 0). */
int vs_camera_tokens(const float* intr_tok, const float* extr_tok, float* x, int frames, int T,
                     int C, int rows_per_frame, vs_stream_t stream);
/* fp32 NCHW image [n,3,h,w] -> bf16 NHWC with 8 channels (3 used) and a zero border of `pad`
 * rows above/below and `pad` / (8 - pad) columns left/right: [n, h + 2*pad, w + 8, 8]. */
int vs_image_nhwc8(const float* img, void* out, int n, int h, int w, int pad, int half_dtype, vs_stream_t stream);
/* SiLU on fp32 rows -> bf16 (AdaLNModulation.nonlinear, backbone_vica.py:210-212) */
int vs_silu_bf16(const float* x, int64_t ldx, void* y, int64_t ldy, int rows, int C, int half_dtype,
                 vs_stream_t stream);
/* Camera head tail: cam_feat fp32 [B*T, ld] (rows of camera_dec_norm output, frame 0 unused) ->
 * pred (B, T-1, 8) normalised dual quaternion and c2w (B, T, 4, 4) with identity prepended.
 * ReLU -> Linear(C->8) -> [...,3] += 1 -> / |q_r| -> homogeneous matrix
 * (vicasplat.py:179-199, misc/dq.py:224-262). */
int vs_camera_head(const float* cam_feat, int64_t ld, const float* w, const float* b, int B, int T,
                   int C, float* pred_dq, float* c2w, vs_stream_t stream);
/* pts head tail: feat bf16 [px, Cf] (post-ReLU) -> 1x1 conv (w fp32 [3, Cf], b[3]) -> exp-depth
 * postprocess xyz = x/|x| * expm1(|x|) (heads/postprocess.py:42-61) -> raw[px, raw_ld] cols 0..2 */
int vs_pts_tail(const void* feat, int Cf, const float* w, const float* b, float* raw, int64_t raw_ld,
                int64_t px, int half_dtype, vs_stream_t stream);
/* MyGaussianAdapter.forward (gaussian_adapter.py:167-212).  Input rows (fp32, leading dimension
 * src_ld) hold the head outputs: xyz at columns [center_col, +3), the 8 + 3*d_sh Gaussian parameters
 * (opacity | scale 3 | quaternion xyzw 4 | SH (xyz d_sh)) at [param_col, ...).  The reference's
 * contiguous raw_gaussians layout is (src_ld = 11 + 3*d_sh, center_col = 0, param_col = 3).
 * Outputs (any may be NULL): raw_out (G, 11 + 3*d_sh) in the reference layout, means (G,3),
 * covariances (G,3,3), packed cov6 (G,6) (triu order xx,xy,xz,yy,yz,zz), harmonics (G,3,d_sh)
 * masked, opacities (G), scales (G,3), rotations (G,4). */
int vs_gaussian_adapter(const float* src, int64_t src_ld, int center_col, int param_col, int64_t G,
                        int d_sh, const float* sh_mask, float* raw_out, float* means, float* cov,
                        float* cov6, float* sh, float* opac, float* scales, float* rot,
                        vs_stream_t stream);

/* ------------------------------------------------------------------ Gaussian rasterizer
 * Tile-based EWA splatting of G Gaussians into V views (diff_gaussian_rasterization semantics,
 * call site cuda_splatting.py:207-235; constants SURVEY.md Appendix D).
 * gaussians_shared != 0: one Gaussian set for all V views (decoder_splatting_cuda.py:79-95 repeats
 * the same set per view; demo.py:226-238 passes it un-batched); otherwise per-view sets (V, G, ...).
 * viewmatrix / projmatrix: V x 16 floats, the *transposed* (column-major) 4x4s the reference
 * passes (cuda_splatting.py:192-194).  shs (G, M, 3) or NULL with colors_precomp (G,3).
 * Outputs: color (V,3,H,W), depth (V,1,H,W), alpha (V,1,H,W), radii (V,G) int32,
 * n_touched (V,G) int32.  final_T/n_contrib (V,H,W) are saved for the backward pass. */
typedef struct vs_raster_params {
  int32_t V, G, H, W;
  int32_t gaussians_shared;
  const float* means3D;    /* (G,3) or (V,G,3) */
  const float* cov3D;      /* (G,6) */
  const float* opacities;  /* (G) */
  const float* shs;        /* (G,M,3) */
  int32_t sh_M, sh_degree;
  int32_t sh_stride_coef, sh_stride_chan; /* element strides inside one Gaussian's 3*M block:
                                             (3,1) = reference layout (G,M,3) [default when both 0];
                                             (1,M) = encoder layout (G,3,M), no transpose copy */
  const float* colors_precomp; /* (G,3) or NULL */
  const float* viewmatrix;     /* (V,16) */
  const float* projmatrix;     /* (V,16) */
  const float* campos;         /* (V,3) */
  const float* tanfov;         /* (V,2) host-side values copied by the caller to device: (x,y) */
  const float* bg;             /* (V,3) */
  float scale_modifier;
  float* out_color;
  float* out_depth;
  float* out_alpha;
  int32_t* radii;
  int32_t* n_touched;
  float* final_T;
  int32_t* n_contrib;
  void* workspace;
  int64_t workspace_bytes;
  int64_t max_pairs; /* capacity (in (tile,splat) pairs) the workspace was sized for */
  int64_t* num_pairs_out; /* device int64[2]: [0] pairs actually produced (caller checks it against
                             max_pairs), [1] largest per-tile count (0 on the global-sort path) */
  int32_t max_tile_pairs; /* > 0: upper bound of the pairs of any single tile (e.g. from a previous
                             call on similar data): enables per-tile binning + shared-memory sort
                             (<= 16384); a tile that exceeds it is left unsorted, which the caller
                             detects from num_pairs_out[1].  0: global 64-bit radix sort. */
} vs_raster_params;
int64_t vs_raster_workspace_bytes(int V, int G, int H, int W, int64_t max_pairs);
int vs_raster_forward(const vs_raster_params* p, vs_stream_t stream);

typedef struct vs_raster_bwd_params {
  vs_raster_params fwd;      /* same inputs + workspace still holding the forward state */
  const float* dL_dcolor;    /* (V,3,H,W) */
  const float* dL_ddepth;    /* (V,1,H,W) or NULL */
  const float* dL_dalpha;    /* (V,1,H,W) or NULL */
  float* dL_dmeans3D;        /* (G,3) or (V,G,3), accumulated (caller zeroes) */
  float* dL_dcov3D;          /* (G,6) */
  float* dL_dopacity;        /* (G) */
  float* dL_dshs;            /* (G,M,3) */
  float* dL_dcolors;         /* (G,3) when colors_precomp */
  float* dL_dtau;            /* (V,6): rho (3) then theta (3), accumulated (caller zeroes); or NULL */
  void* bwd_workspace;       /* scratch: V*G*10 floats (per-view screen-space partials) */
  int64_t bwd_workspace_bytes;
} vs_raster_bwd_params;
int vs_raster_backward(const vs_raster_bwd_params* p, vs_stream_t stream);

/* ------------------------------------------------------------------ MSE loss (+ its gradient)
 * LossMse.forward, src/loss/loss_mse.py:23-31: loss = weight * mean((pred - target)^2), and in the
 * same pass dL/dpred = (2 * weight / n) * (pred - target) (nullable): the render is read once
 * instead of three times (sub, square+mean, backward).  Deterministic: per-block partial sums in
 * `workspace`, reduced in a fixed order by the last block to finish.  loss_out: device float[1]. */
int64_t vs_mse_workspace_bytes(void);
int vs_mse_loss(const float* pred, const float* target, int64_t n, float weight, float* loss_out,
                float* grad_out, void* workspace, vs_stream_t stream);

/* ------------------------------------------------------------------ test-time pose update
 * update_pose, src/misc/cam_utils.py:127-148: per camera c2w' = inv(SE3_exp([rho, theta]) * inv(c2w))
 * with the reference's small-angle branches (|theta| < 1e-5, cam_utils.py:74-108); general 4x4
 * inverses like the reference's .inverse().  rho, theta: (n,3); c2w, c2w_out: (n,4,4) row-major
 * (c2w_out may alias c2w).  Replaces a host loop of n SE3_exp calls + two batched inverses. */
int vs_update_pose(const float* rho, const float* theta, const float* c2w, float* c2w_out, int n,
                   vs_stream_t stream);

/* ------------------------------------------------------------------ fused AdamW (+ clip + non-finite scan)
 * One optimizer step over a LIST of fp32 tensors in two launches (the reference steps 847 tensors
 * one by one: torch.optim.AdamW configured at src/model/model_wrapper.py:884-951, betas (0.9, 0.95),
 * weight_decay 0.05, two learning-rate groups; Lightning clips the global gradient norm to
 * gradient_clip_val = 0.5, config/main.yaml:70).
 *   pass 1  global ||g||_2 (deterministic two-stage sum) and a non-finite flag;
 *   pass 2  g' = g * min(1, max_grad_norm / (norm + 1e-6))   (max_grad_norm <= 0: no clipping)
 *           p *= 1 - lr * weight_decay;  m = b1 m + (1 - b1) g';  v = b2 v + (1 - b2) g'^2
 *           p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)        (torch.optim.AdamW)
 *           skip_nonfinite: 0 = gradients used as they are; 1 = the whole step is skipped when a gradient
 *           held an inf / nan; 2 = every non-finite gradient element is replaced as torch.nan_to_num_
 *           does (nan -> 0, +-inf -> +-FLT_MAX) and the step is taken: the reference's
 *           GradientNanCheckCallback (src/main.py:40-45).
 *           step_counter (device int32, nullable): when given, t is read from it instead of `step`; the
 *           first launch advances it, unless the step is being skipped -- so the bias corrections
 *           (and the 'step' a checkpoint records) never run ahead of the moments.
 * Tensors are described by device tables; work is cut into chunks of VS_ADAMW_CHUNK elements:
 * chunk c covers elements [chunk_start[c], chunk_start[c] + VS_ADAMW_CHUNK) of tensor chunk_tensor[c]. */
#define VS_ADAMW_CHUNK 16384
typedef struct vs_adamw_params {
  int32_t n_tensors, n_chunks;
  float* const* params;        /* device array [n_tensors] of device pointers */
  const float* const* grads;
  float* const* exp_avg;
  float* const* exp_avg_sq;
  const int64_t* sizes;        /* device [n_tensors] */
  const float* lrs;            /* device [n_tensors]: learning rate of each tensor's group */
  const int32_t* chunk_tensor; /* device [n_chunks] */
  const int64_t* chunk_start;  /* device [n_chunks] */
  float beta1, beta2, eps, weight_decay;
  int32_t step;                /* t >= 1 */
  float max_grad_norm;
  int32_t skip_nonfinite;
  float* partials;             /* device scratch [n_chunks + 2] floats */
  uint32_t* counter;           /* device uint32, zero before the first call (left at zero) */
  float* grad_norm_out;        /* device float[1]: ||g||_2 before clipping */
  int32_t* found_inf_out;      /* device int32[1] */
  int32_t* step_counter;       /* device int32[1] or NULL (see above) */
} vs_adamw_params;
int vs_adamw_step(const vs_adamw_params* p, vs_stream_t stream);

/* ------------------------------------------------------------------ encoder training path (backward)
 * Hand-written backward pass of the ViT block (croco/blocks.py:58-130: the reference gets it from
 * torch.autograd).  The dense contractions run on vs_gemm:
 *   dgrad  dX[M,K]  = dY[M,N] W[N,K]      -> vs_gemm(A = dY bf16, W = W^T bf16 [K,N])
 *   wgrad  dW[N,K] += dY^T[N,M] X[M,K]    -> vs_gemm(A = dY^T bf16 [N,M], W = X^T bf16 [K,M], C = res1 = dW fp32)
 * and the bandwidth-bound work around them is below. */

/* One pass over a gradient matrix src (rows, cols), f32 or bf16: optionally multiplied elementwise by
 * gelu'(z) (z = the bf16 pre-activation of an MLP's fc1, croco/blocks.py:60,76), it writes any of
 *   copy        (rows, cols) bf16 row-major               -- the dgrad A operand
 *   transposed  (cols, ld_t) bf16, ld_t % 8 == 0 >= rows  -- the wgrad A (or W) operand; pad columns zeroed
 *   colsum      (cols) f32, ACCUMULATED with atomics      -- the bias gradient (caller zeroes)
 * cols % 4 == 0; every row start 16-byte (f32) / 8-byte (bf16) aligned. */
int vs_grad_prep(const void* src, int src_dtype, int64_t ld_src, const void* z, int64_t ld_z, int rows,
                 int cols, void* copy, int64_t ld_copy, void* transposed, int64_t ld_t, float* colsum,
                 vs_stream_t stream);

/* Training-mode GELU (exact erf, nn.GELU default): a = gelu(z), bf16 -> bf16; in training the fc1
 * GEMM stores the pre-activation z (needed by vs_grad_prep) instead of fusing the GELU. */
int vs_gelu_bf16(const void* z, int64_t ld_z, void* a, int64_t ld_a, int rows, int cols,
                 vs_stream_t stream);

/* Backward of y = LayerNorm(x) * gamma + beta (eps as given): dx = dres + dLN/dx (dres nullable: the
 * gradient arriving through the residual connection; dx may alias it), dgamma / dbeta (nullable)
 * ACCUMULATED with atomics (caller zeroes).  x f32, dy f32 or bf16, C = k*128 <= 1024. */
typedef struct vs_layernorm_bwd_params {
  const float* x;
  int64_t ldx;
  const void* dy;
  int32_t dy_dtype;
  int64_t ldy;
  const float* gamma;
  const float* dres;
  int64_t ldres;
  float* dx;
  int64_t lddx;
  float* dgamma;
  float* dbeta;
  int32_t rows, C;
  float eps;
} vs_layernorm_bwd_params;
int vs_layernorm_backward(const vs_layernorm_bwd_params* p, vs_stream_t stream);

/* Backward of vs_attention (same items / segments / mask; `fwd` holds the forward call's arguments,
 * including O and the lse it wrote): dQ, dK, dV (bf16, row matrices like Q / K / V, head h at columns
 * [h*64, h*64+64); they may alias a packed dqkv buffer) from dO.  delta: scratch (q_rows, heads) f32.
 * fwd.max_kv_len must be given.  The key rows of different items must not overlap (dK / dV are
 * written, not accumulated) unless the key-centric dkv_* tables below are given.  Flash-style: scores are recomputed on the tensor cores, no (q, kv)
 * matrix is ever stored.  Replaces autograd through croco/blocks.py:105-109. */
typedef struct vs_attention_bwd_params {
  vs_attention_params fwd;
  const void* dO;
  int64_t lddo;
  void *dQ, *dK, *dV;
  int64_t lddq, lddk, lddv;
  float* delta;
  /* Optional KEY-centric item tables for the dK / dV pass (dkv_items > 0).  Needed when key rows are
   * shared between the forward items (CrossNeighborAttention, backbone_vica.py:173-183: frame j is a
   * neighbour of up to two query frames): item i = keys [dkv_kv_start[i], + dkv_kv_len[i]) and the
   * queries of up to two row segments that attended to them; every dK / dV row is written once.  The
   * statistics (lse, delta) stay indexed by absolute query row.  A key frame that a query frame lists
   * twice (end frames see their single neighbour twice) must appear ONCE in that query's forward item
   * (kv_len1 = 0), as vs_attention's callers already do. */
  int32_t dkv_items, dkv_max_kv_len;
  const int32_t *dkv_kv_start, *dkv_kv_len, *dkv_q_start0, *dkv_q_len0, *dkv_q_start1, *dkv_q_len1;
} vs_attention_bwd_params;
int vs_attention_backward(const vs_attention_bwd_params* p, vs_stream_t stream);

/* Backward of vs_rope_rows / of the rope fused into the qkv epilogue: the inverse rotation applied in
 * place to the q and k columns of the packed bf16 gradient dqkv (cuRoPE2D_func.backward calls
 * rope_2d with fwd = -1 the same way, curope2d.py:24-29). */
int vs_rope_rows_backward(void* dqkv, int64_t ld, int rows, int H, int q_col, int k_col,
                          const int32_t* pos, float base, float cam_theta, vs_stream_t stream);

/* ------------------------------------------------------------------ decoder / head training path (backward)
 * MixDecoderBlock (backbone_vica.py:194-335), DPT heads (heads/dpt_block.py:79-229,264-459,
 * heads/dpt_gs_head.py:98-157) and the per-pixel tails, differentiated by hand; the reference gets these
 * gradients from torch.autograd.  Contractions run on vs_gemm (dgrad: flipped-tap a_mode 1 / rows mode,
 * wgrad: a_mode 2 / 3 with c_accumulate), attention on vs_attention_backward. */

/* Backward of h = LN(x) * gamma + beta, modulated per frame: h = h * (1 + scale_f) + shift_f
 * (vs_layernorm with rows_per_frame; scale == NULL: plain LayerNorm).  dx = dres + dLN/dx for the rows
 * of every frame except (skip_first) its first row, whose gradient is passed through (dx = dres);
 * frame_a[f] += sum_rows dh * xhat, frame_b[f] += sum_rows dh (atomics, caller zeroes): vs_adaln_reduce
 * turns them into d scale / d shift / d gamma / d beta.  x, dx, dres fp32; dh f32 or bf16. */
typedef struct vs_ln_mod_bwd_params {
  const float* x;
  int64_t ldx;
  const void* dh;
  int32_t dh_dtype;
  int64_t lddh;
  const float* gamma;
  const float* scale;
  int64_t mod_ld;
  const float* dres;
  int64_t ldres;
  float* dx;
  int64_t lddx;
  float* frame_a;
  float* frame_b;
  int64_t frame_ld;
  int32_t frames, rows_per_frame, skip_first, C;
  float eps;
} vs_ln_mod_bwd_params;
int vs_layernorm_mod_backward(const vs_ln_mod_bwd_params* p, vs_stream_t stream);
/* dscale[f] = gamma * A_f + beta * B_f, dshift[f] = B_f (written; nullable pair);
 * dgamma += sum_f (1 + scale_f) A_f, dbeta += sum_f (1 + scale_f) B_f (nullable pair; scale NULL = 0). */
int vs_adaln_reduce(const float* frame_a, const float* frame_b, int64_t frame_ld, const float* gamma,
                    const float* beta, const float* scale, int64_t mod_ld, float* dscale, float* dshift,
                    int64_t dmod_ld, float* dgamma, float* dbeta, int frames, int C, vs_stream_t stream);
/* Training-forward form of the gated residual: out[row] = x[row] + (1 + gate[row / rows_per_frame]) *
 * branch[row] (out may alias x; branch bf16 = the projection output the backward pass needs; the
 * inference path fuses this into the GEMM epilogue).  first_row_mode for the first row of each frame:
 * 0 as others, 1 no gate, 2 copied unchanged. */
int vs_gate_residual(const float* x, int64_t ldx, float* out, int64_t ldo, const void* branch, int64_t ldb,
                     const float* gate, int64_t gate_ld, int64_t rows, int C, int rows_per_frame,
                     int first_row_mode, vs_stream_t stream);
/* Backward of the gated residual: dbranch (bf16) = dout * (1 + gate_f) [first rows: mode 1 = dout, mode 2
 * = 0]; dgate[f] += sum_rows dout * branch (nullable); colsum += column sums of dbranch (the bias gradient
 * of the projection; nullable).  Atomic accumulation, caller zeroes.  gate == NULL: plain copy + colsum. */
int vs_gate_backward(const float* dout, int64_t ldd, const void* branch, int64_t ldb, const float* gate,
                     int64_t gate_ld, void* dbranch, int64_t lddb, float* dgate, int64_t dgate_ld,
                     float* colsum, int frames, int rows_per_frame, int C, int first_row_mode,
                     vs_stream_t stream);
/* dx (=|+=) dy * silu'(x) on fp32 rows (AdaLNModulation.nonlinear, backbone_vica.py:210-212) */
int vs_silu_backward(const float* x, int64_t ldx, const float* dy, int64_t lddy, float* dx, int64_t lddx,
                     int rows, int C, int accumulate, vs_stream_t stream);
/* transpose of vs_upsample2x: dy NHWC bf16 [n, 2h, 2w, c] -> dx [n, h, w, c] */
int vs_upsample2x_backward(const void* dy, void* dx, int n, int h, int w, int c, vs_stream_t stream);
/* inverse of vs_pixel_shuffle: NHWC bf16 [n, h*k, w*k, c] -> rows [n*h*w, k*k*c] */
int vs_pixel_unshuffle(const void* src, void* dst, int n, int h, int w, int c, int k, vs_stream_t stream);
/* transpose of vs_im2col (NHWC bf16): dcols [n*ho*wo, kpad] -> dx [n, h, w, c], overlapping taps summed */
int vs_col2im(const void* dcols, void* dx, int n, int h, int w, int c, int k, int stride, int pad, int kpad,
              vs_stream_t stream);
/* dx = dy * (y > 0) on bf16 [rows, C] (dx nullable / may alias dy); colsum += column sums (nullable) */
int vs_relu_backward(const void* dy, int64_t lddy, const void* y, int64_t ldy, void* dx, int64_t lddx,
                     float* colsum, int64_t rows, int C, vs_stream_t stream);
/* nn.Dropout(p), training mode, in place on a bf16 buffer of n elements (dpt_block.py:341): keep with
 * probability 1 - p, scale by 1 / (1 - p); the keep decision is a hash of (seed, element index). */
int vs_dropout_bf16(void* x, int64_t n, float p, uint64_t seed, vs_stream_t stream);
/* Backward of vs_pts_tail: d_xyz fp32 rows (leading dimension d_ld) -> d_feat bf16 [px, Cf] (already
 * masked by feat > 0), dw (3, Cf) and db (3) accumulated. */
int vs_pts_tail_backward(const void* feat, int Cf, const float* w, const float* b, const float* d_xyz,
                         int64_t d_ld, void* d_feat, float* dw, float* db, int64_t px, vs_stream_t stream);
/* Backward of vs_gaussian_adapter: gradients of its outputs (any may be NULL: d_raw (G, 11 + 3 d_sh),
 * d_means (G,3), d_cov (G,3,3), d_cov6 (G,6), d_shs (G,3,d_sh), d_opac (G)) -> d_src rows in the layout of
 * `src` (centre columns and the 8 + 3 d_sh parameter columns are WRITTEN, others untouched). */
int vs_gaussian_adapter_backward(const float* src, int64_t src_ld, int center_col, int param_col, int64_t G,
                                 int d_sh, const float* sh_mask, const float* d_raw, const float* d_means,
                                 const float* d_cov, const float* d_cov6, const float* d_shs,
                                 const float* d_opac, float* d_src, int64_t dsrc_ld, vs_stream_t stream);
/* Backward of vs_camera_head w.r.t. pred_dq: d_feat (B*T, ldd) written (frame 0 rows = 0), dw (8, C) and
 * db (8) accumulated. */
int vs_camera_head_backward(const float* cam_feat, int64_t ld, const float* w, const float* b, int B, int T,
                            int C, const float* d_pred, float* d_feat, int64_t ldd, float* dw, float* db,
                            vs_stream_t stream);

/* ------------------------------------------------------------------ LPIPS consumer (src/loss/loss_lpips.py:27-54)
 * The VGG16 features run on vs_gemm (conv3x3 + bias + ReLU); these are the pieces around it. */
/* 2x2 / stride 2 max pooling on NHWC bf16 maps, and its backward: dx = add (nullable) + dy routed to the first
 * element of each window that equals the pooled value; relu_mask != 0: x is a post-ReLU map and dx is the
 * gradient of its pre-activation (a zero maximum passes nothing). */
int vs_maxpool2(const void* x, void* y, int n, int h, int w, int c, vs_stream_t stream);
int vs_maxpool2_backward(const void* x, const void* y, const void* dy, const void* add, void* dx, int n, int h,
                         int w, int c, int relu_mask, vs_stream_t stream);
/* One LPIPS layer: per_image[i] += mean over the image's pixels of sum_c w[c] (f0/(|f0|+1e-10) - f1/(|f1|+1e-10))^2
 * (atomics, caller zeroes), and -- if df0 != NULL -- the gradient of grad_scale * (that sum over pixels) w.r.t.
 * f0, already masked by f0 > 0 (the features are post-ReLU), bf16.  f0 / f1: bf16 (pixels, C), C in {64, 128, 256,
 * 512}, 16-byte aligned, hw pixels per image (one atomic per 128 pixels when hw % 128 == 0, else one per pixel). */
int vs_lpips_layer(const void* f0, const void* f1, const float* wlin, int64_t pixels, int C, int hw,
                   float grad_scale, float* per_image, void* df0, vs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VICASPLAT_B200_H_ */
