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
} vs_attention_params;
int vs_attention(const vs_attention_params* p, vs_stream_t stream);

/* ------------------------------------------------------------------ small fused ops */
/* Non-overlapping patch gather: img fp32 NCHW [n,3,h,w] -> bf16 [n*(h/P)*(w/P), 3*P*P]
 * (channel-major, then row, then column: the flattening order of Conv2d weight [E,3,P,P];
 * croco/blocks.py:195-225). */
int vs_patchify(const float* img, void* out, int n, int h, int w, int P, vs_stream_t stream);
/* kxk stride-s im2col of an NHWC bf16 map (or NCHW fp32 image when src_nchw_f32 != 0) into
 * bf16 [n*ho*wo, kpad], column order (tap-major, channel-minor), zero padded to kpad. */
int vs_im2col(const void* src, int src_nchw_f32, void* out, int n, int h, int w, int c, int k,
              int stride, int pad, int kpad, vs_stream_t stream);
/* bilinear x2, align_corners=True, NHWC bf16 (heads/dpt_block.py:214-216) */
int vs_upsample2x(const void* src, void* dst, int n, int h, int w, int c, vs_stream_t stream);
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
int vs_image_nhwc8(const float* img, void* out, int n, int h, int w, int pad, vs_stream_t stream);
/* SiLU on fp32 rows -> bf16 (AdaLNModulation.nonlinear, backbone_vica.py:210-212) */
int vs_silu_bf16(const float* x, int64_t ldx, void* y, int64_t ldy, int rows, int C,
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
                int64_t px, vs_stream_t stream);
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
 *           the whole step is skipped when skip_nonfinite != 0 and a gradient held an inf / nan.
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
 * written, not accumulated).  Flash-style: scores are recomputed on the tensor cores, no (q, kv)
 * matrix is ever stored.  Replaces autograd through croco/blocks.py:105-109. */
typedef struct vs_attention_bwd_params {
  vs_attention_params fwd;
  const void* dO;
  int64_t lddo;
  void *dQ, *dK, *dV;
  int64_t lddq, lddk, lddv;
  float* delta;
} vs_attention_bwd_params;
int vs_attention_backward(const vs_attention_bwd_params* p, vs_stream_t stream);

/* Backward of vs_rope_rows / of the rope fused into the qkv epilogue: the inverse rotation applied in
 * place to the q and k columns of the packed bf16 gradient dqkv (cuRoPE2D_func.backward calls
 * rope_2d with fwd = -1 the same way, curope2d.py:24-29). */
int vs_rope_rows_backward(void* dqkv, int64_t ld, int rows, int H, int q_col, int k_col,
                          const int32_t* pos, float base, float cam_theta, vs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VICASPLAT_B200_H_ */
