// Device-side argument block and tile constants of gemm_tc05_kernel, shared by the host entry (gemm.cu)
// and the three kernel translation units (gemm_fwd.cu, gemm_fwd16.cu, gemm_train.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace vs {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row

enum : int {  // vector-access flags, decided on the host from pointer / stride alignment
  VEC_BIAS = 1, VEC_GATE = 2, VEC_RES = 4, VEC_C = 8, VEC_C2 = 16
};

struct GemmDev {
  int mode;
  int tn;   // rows mode with both operands stored [k][mn] (MN-major): the wgrad form
  int wg;   // tn with both operands addressed as NHWC maps (conv wgrad): k-block = 64-pixel box
  int cin_pad;
  int N;
  int num_kb;
  int m_tiles, n_tiles;
  int splits, kb_per_split;   // split-K: unit = (split, n tile, m group); needs the atomic epilogue
  int atomic;                 // C += (red.global.add) instead of C =
  int mask_mode;              // 0 none, 1 res2 masks then + res1, 2 res1 masks (bf16 residual kinds)
  float out_scale;
  int f16;                    // 16-bit operands / outputs / residual maps are fp16 instead of bf16
  // rows mode
  int a_rows, a_groups, tiles_per_group;
  // conv mode
  int cn, ch, cw, bw, bh, bn, tiles_x, tiles_y, cblocks, kw, pad;
  // epilogue
  const float* bias;
  int act;
  const float* gate;
  int gate_ld;
  int gate_rows, first_row_mode;
  const void* res1;
  const void* res2;
  int res_dtype;
  int res_up2;
  int res_ld;
  void* C;
  int c_dtype;
  int ldc;        // leading dimensions fit 31 bits (checked on the host): row * ld is one IMAD.WIDE
  __nv_bfloat16* C2;
  int ldc2;
  int out_gin, out_gout, out_off;
  int vec;
  int fast;      // every present C / C2 / residual pointer allows aligned 4-column segments
  int res_kind;  // 0 none, 1 fp32, 2 bf16, 3 bilinear-x2 bf16
  float up_sx, up_sy;  // res_up2: source step per output pixel (align_corners=True)
  // rotary embedding of the q / k column blocks (nullptr = none); inverse frequencies live in the
  // kernel parameter (constant) bank
  const int* rope_pos;
  int rope_q0, rope_k0, rope_cols;
  float rope_if_img[16], rope_if_cam[32];
};

// one per kernel variant (VS_GEMM_VARIANT): picks the (tile width, cluster size) instantiation and launches
int gemm_launch_fwd(int bn, int cl, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmDev& g, cudaStream_t stream);
int gemm_launch_fwd16(int bn, int cl, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmDev& g, cudaStream_t stream);
int gemm_launch_train(int bn, int cl, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmDev& g, cudaStream_t stream);
int gemm_num_sms();

}  // namespace vs
