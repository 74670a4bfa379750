// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   C[m, n] = epilogue( sum_k A[m, k] * W[n, k] )      bf16 x bf16 -> fp32 (TMEM accumulator)
//
// Persistent, warp-specialised: one CTA per SM loops over 128 x BN output tiles.
//   warp 0       TMA producer: A tile (128 rows x 64 k) and W tile (BN rows x 64 k) per k-block into
//                a STAGES-deep ring (~192 KB in flight) of 128B-swizzled tiles, mbarrier completion
//   warp 1       TMEM allocator + single-thread tcgen05.mma issuer (4 x K=16 UMMAs per k-block);
//                tcgen05.commit frees ring slots and hands a finished accumulator to the epilogue
//   warps 2..    epilogue: thread r owns tile row r (TMEM lane r): tcgen05.ld 32 columns at a time,
//                bias / activation / AdaLN gate / residual(s) in registers, 16-byte row-wise
//                loads and stores (every thread touches whole 32-byte sectors of its own row)
// The accumulator is double buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps
// the main loop of tile i+1.  With CL == 2 the two CTAs of a cluster form a CTA pair
// (tcgen05 cta_group::2) on a 256 x BN tile: each CTA loads its own 128 rows of A and HALF of the
// W tile, the leader issues one M = 256 UMMA that reads both shared memories and writes both
// TMEMs.  A single-CTA 128 x 256 tile ingests 48 KB per 512 MMA cycles per SM -- more than the
// L2 -> SM port sustains (measured ~45 B/clk/SM, tensor pipe 45-55 % busy); the pair needs 32 KB and
// fits two more ring stages.  The A operand is addressed through a TMA tensor map, which is what
// lets the same kernel serve nn.Linear on (strided / grouped) token rows and stride-1 k x k
// convolutions on NHWC maps (one 4-D box per filter tap; out-of-bounds = zero padding).
//
// Replaces in the reference: every nn.Linear in croco/blocks.py:58-130 and backbone_vica.py:57-335
// and every stride-1 nn.Conv2d in heads/dpt_block.py:79-229,264-459, heads/dpt_gs_head.py:98-157.
//
// This file is the host entry (argument checks, tensor maps, tile choice); the kernel is gemm_kernel.cuh.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cuda_bf16.h>

#include "common.h"
#include "gemm_dev.h"
#include "tmap.h"

namespace vs {

int gemm_num_sms() {
  static int n = []() {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

namespace {

inline int num_sms() { return gemm_num_sms(); }

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// cycles of one persistent CTA for tile width bn: waves x k-blocks x max(tensor pipe, L2 feed)
double tile_cost(int bn, int m_tiles, int N, int num_kb) {
  const int n_tiles = ceil_div(N, bn);
  const long long tiles = static_cast<long long>(m_tiles) * n_tiles;
  const int sms = num_sms();
  const double waves = static_cast<double>((tiles + sms - 1) / sms);
  const double active = tiles < sms ? static_cast<double>(tiles) : static_cast<double>(sms);
  const double mma = 2.0 * bn;                                   // 4 UMMAs of 128 x bn x 16
  const double wshare = (bn >= 128 && m_tiles >= 2) ? 0.5 : 1.0;  // a CTA pair splits the W tile
  const double l2 = active * (BM + wshare * bn) * 128.0 / 6300.0;  // chip-wide L2 -> SM feed (B/clk)
  const double per_kb = mma > l2 ? mma : l2;
  return waves * (num_kb * per_kb + 1500.0 /*pipeline fill + epilogue tail*/);
}

}  // namespace
}  // namespace vs

extern "C" int vs_gemm(const vs_gemm_params* p, vs_stream_t stream_) {
  using namespace vs;
  VS_REQUIRE(p != nullptr, "vs_gemm: null params");
  VS_REQUIRE(p->A && p->W, "vs_gemm: A and W must be non-null");
  VS_REQUIRE(p->N > 0, "vs_gemm: N must be positive");
  VS_REQUIRE(al16(p->A) && al16(p->W), "vs_gemm: A and W must be 16-byte aligned");
  VS_REQUIRE(p->w_row_stride % 8 == 0, "vs_gemm: w_row_stride must be a multiple of 8 elements");
  cudaStream_t stream = to_stream(stream_);

  GemmDev g{};
  // a_mode 2 / 3 = rows mode with MN-major operands (g.tn); 3 addresses both through NHWC maps (g.wg)
  g.mode = (p->a_mode == 2 || p->a_mode == 3) ? 0 : p->a_mode;
  g.splits = 1;
  g.atomic = p->c_accumulate != 0;
  g.mask_mode = p->mask_mode;
  g.out_scale = p->out_scale == 0.f ? 1.0f : p->out_scale;
  VS_REQUIRE(p->mask_mode >= 0 && p->mask_mode <= 2, "vs_gemm: mask_mode must be 0, 1 or 2");
  VS_REQUIRE(g.out_scale == 1.0f || p->mask_mode != 0, "vs_gemm: out_scale goes with a mask_mode");
  VS_REQUIRE(p->mask_mode == 0 || (p->res1 && p->res_dtype != VS_F32 && !p->res_up2 &&
                                   (p->mask_mode == 2 ? p->res2 == nullptr : p->res2 != nullptr)),
             "vs_gemm: mask_mode needs bf16 maps (mode 1: res1 + mask res2; mode 2: mask res1 only)");
  VS_REQUIRE(!p->c_accumulate || (p->c_dtype == VS_F32 && p->C && !p->bias && p->act == VS_ACT_NONE &&
                                  !p->gate && !p->res1 && !p->C2 && !p->rope_pos),
             "vs_gemm: c_accumulate needs an fp32 C and a plain epilogue");
  VS_REQUIRE(p->split_k >= 0 && (p->split_k <= 1 || p->c_accumulate),
             "vs_gemm: split_k needs c_accumulate");
  g.N = p->N;
  g.bias = p->bias;
  g.act = p->act;
  g.gate = p->gate;
  g.gate_ld = static_cast<int>(p->gate_ld);
  g.gate_rows = p->gate_rows;
  g.first_row_mode = p->first_row_mode;
  g.res1 = p->res1;
  g.res2 = p->res2;
  g.res_dtype = p->res_dtype;
  g.res_up2 = p->res_up2;
  g.res_ld = static_cast<int>(p->res_ld);
  VS_REQUIRE(!p->res_up2 || (p->a_mode == 1 && p->res1 && p->res_dtype != VS_F32 && !p->res2 &&
                             p->ch % 2 == 0 && p->cw % 2 == 0),
             "vs_gemm: res_up2 needs conv mode, one bf16 residual map and even output sizes");
  g.C = p->C;
  g.c_dtype = p->c_dtype;
  VS_REQUIRE(p->ldc >= 0 && p->ldc < (1ll << 31) && p->ldc2 >= 0 && p->ldc2 < (1ll << 31) &&
                 p->res_ld >= 0 && p->res_ld < (1ll << 31) && p->gate_ld >= 0 && p->gate_ld < (1ll << 31),
             "vs_gemm: leading dimensions must fit 31 bits");
  g.ldc = static_cast<int>(p->ldc);
  g.C2 = static_cast<__nv_bfloat16*>(p->C2);
  g.ldc2 = static_cast<int>(p->ldc2);
  g.out_gin = p->out_gin;
  g.out_gout = p->out_gout;
  g.out_off = p->out_off;
  VS_REQUIRE(p->res2 == nullptr || p->res1 != nullptr, "vs_gemm: res2 requires res1");
  // 16-bit operand format: bf16 (0 / VS_BF16) or fp16 (VS_F16); 16-bit outputs / residual maps share it
  VS_REQUIRE(p->operand_dtype == 0 || p->operand_dtype == VS_BF16 || p->operand_dtype == VS_F16,
             "vs_gemm: operand_dtype must be bf16 or fp16");
  g.f16 = p->operand_dtype == VS_F16;
  const int half_t = g.f16 ? VS_F16 : VS_BF16;
  VS_REQUIRE(p->c_dtype == VS_F32 || p->c_dtype == half_t, "vs_gemm: c_dtype must be f32 or the operand format");
  VS_REQUIRE(p->res_dtype == VS_F32 || p->res_dtype == half_t || p->res1 == nullptr,
             "vs_gemm: res_dtype must be f32 or the operand format");
  VS_REQUIRE(!g.f16 || p->a_mode < 2, "vs_gemm: the fp16 operand format is a forward-path (a_mode 0 / 1) option");
  g.vec = 0;
  if (p->bias && al16(p->bias)) g.vec |= VEC_BIAS;
  if (p->gate && al16(p->gate) && p->gate_ld % 4 == 0) g.vec |= VEC_GATE;
  if (p->res1 && al16(p->res1) && (p->res2 == nullptr || al16(p->res2)) &&
      p->res_ld % (p->res_dtype == VS_F32 ? 4 : 8) == 0)
    g.vec |= VEC_RES;
  if (p->C && al16(p->C) && p->ldc % (p->c_dtype == VS_F32 ? 4 : 8) == 0) g.vec |= VEC_C;
  if (p->C2 && al16(p->C2) && p->ldc2 % 8 == 0) g.vec |= VEC_C2;
  if (p->res_up2) {
    g.up_sy = p->ch > 1 ? static_cast<float>(p->ch / 2 - 1) / static_cast<float>(p->ch - 1) : 0.f;
    g.up_sx = p->cw > 1 ? static_cast<float>(p->cw / 2 - 1) / static_cast<float>(p->cw - 1) : 0.f;
  }
  g.rope_pos = p->rope_pos;
  if (p->rope_pos != nullptr) {
    VS_REQUIRE(p->rope_heads > 0 && p->rope_q_col % 64 == 0 && p->rope_k_col % 64 == 0 &&
                   p->rope_q_col >= 0 && p->rope_k_col >= 0,
               "vs_gemm: rope column offsets must be non-negative multiples of 64");
    VS_REQUIRE(p->rope_base > 0.f && p->rope_cam_theta > 0.f, "vs_gemm: rope bases must be positive");
    VS_REQUIRE(p->act == VS_ACT_NONE && p->gate == nullptr, "vs_gemm: rope goes with a plain epilogue");
    g.rope_q0 = p->rope_q_col;
    g.rope_k0 = p->rope_k_col;
    g.rope_cols = p->rope_heads * 64;
    for (int d = 0; d < 16; ++d)
      g.rope_if_img[d] = static_cast<float>(1.0 / pow(static_cast<double>(p->rope_base), d / 16.0));
    for (int l = 0; l < 32; ++l)
      g.rope_if_cam[l] = static_cast<float>(1.0 / pow(static_cast<double>(p->rope_cam_theta), (2 * l) / 64.0));
  }
  g.res_kind = !p->res1 ? 0 : p->res_up2 ? 3 : p->res_dtype == VS_F32 ? 1 : 2;
  g.fast = p->C && (g.vec & VEC_C) && (!p->C2 || (g.vec & VEC_C2)) && (!p->res1 || (g.vec & VEC_RES));

  CUtensorMap tmA, tmW;
  int m_tiles = 0;
  long long ktot = 0;
  if (p->a_mode == 0) {
    VS_REQUIRE(p->a_rows > 0 && p->a_groups > 0 && p->K > 0, "vs_gemm: empty problem");
    VS_REQUIRE(p->a_row_stride % 8 == 0, "vs_gemm: a_row_stride must be a multiple of 8");
    VS_REQUIRE(p->a_groups == 1 || p->a_group_stride % 8 == 0,
               "vs_gemm: a_group_stride must be a multiple of 8");
    g.a_rows = p->a_rows;
    g.a_groups = p->a_groups;
    g.tiles_per_group = ceil_div(p->a_rows, BM);
    m_tiles = g.tiles_per_group * p->a_groups;
    ktot = p->K;
    const long long gstride =
        p->a_groups > 1 ? p->a_group_stride : static_cast<long long>(p->a_rows) * p->a_row_stride;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(p->K), static_cast<cuuint64_t>(p->a_rows),
                          static_cast<cuuint64_t>(p->a_groups)};
    cuuint64_t str[2] = {static_cast<cuuint64_t>(p->a_row_stride) * 2,
                         static_cast<cuuint64_t>(gstride) * 2};
    cuuint32_t box[3] = {BK, BM, 1};
    int rc = encode_map(&tmA, p->A, 3, dims, str, box);
    if (rc) return rc;
  } else if (p->a_mode == 1) {
    VS_REQUIRE(p->cn > 0 && p->ch > 0 && p->cw > 0 && p->cin > 0, "vs_gemm: empty conv input");
    VS_REQUIRE(p->cin % 8 == 0, "vs_gemm: conv cin must be a multiple of 8");
    VS_REQUIRE(p->kh > 0 && p->kw > 0 && p->pad >= 0, "vs_gemm: bad conv kernel");
    g.cn = p->cn;
    g.ch = p->ch;
    g.cw = p->cw;
    g.kw = p->kw;
    g.pad = p->pad;
    // 128-pixel tile = bw x bh x bn box
    int bw = 1;
    while (bw < 16 && bw < p->cw) bw <<= 1;
    int bh = 1;
    while (bw * bh < BM && bh < p->ch) bh <<= 1;
    int bn = BM / (bw * bh);
    g.bw = bw;
    g.bh = bh;
    g.bn = bn;
    g.tiles_x = ceil_div(p->cw, bw);
    g.tiles_y = ceil_div(p->ch, bh);
    m_tiles = g.tiles_x * g.tiles_y * ceil_div(p->cn, bn);
    g.cblocks = ceil_div(p->cin, BK);
    ktot = static_cast<long long>(p->kh) * p->kw * g.cblocks * BK;
    const long long in_h = p->conv_in_h > 0 ? p->conv_in_h : p->ch;
    const long long sx = p->conv_stride_x > 0 ? p->conv_stride_x : p->cin;
    const long long sy = p->conv_stride_y > 0 ? p->conv_stride_y : sx * p->cw;
    const long long sn = p->conv_stride_n > 0 ? p->conv_stride_n : sy * in_h;
    VS_REQUIRE(sx % 8 == 0 && sy % 8 == 0 && sn % 8 == 0,
               "vs_gemm: conv strides must be multiples of 8 elements");
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(p->cin), static_cast<cuuint64_t>(p->cw),
                          static_cast<cuuint64_t>(in_h), static_cast<cuuint64_t>(p->cn)};
    cuuint64_t str[3] = {static_cast<cuuint64_t>(sx) * 2, static_cast<cuuint64_t>(sy) * 2,
                         static_cast<cuuint64_t>(sn) * 2};
    cuuint32_t box[4] = {BK, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh),
                         static_cast<cuuint32_t>(bn)};
    int rc = encode_map(&tmA, p->A, 4, dims, str, box);
    if (rc) return rc;
  } else if (p->a_mode == 2) {
    // both operands given as [K][mn] row matrices (MN-major): C[m, n] = sum_k A[k, m] * W[k, n]
    VS_REQUIRE(p->a_rows > 0 && p->K > 0, "vs_gemm: empty problem");
    VS_REQUIRE(p->a_groups <= 1, "vs_gemm: a_mode 2 has no row groups");
    VS_REQUIRE(p->a_row_stride % 8 == 0 && p->a_row_stride >= p->a_rows,
               "vs_gemm: a_mode 2: a_row_stride (the stride between k-rows of A) must be a multiple of 8");
    VS_REQUIRE(p->w_row_stride >= p->N, "vs_gemm: a_mode 2: w_row_stride is the stride between k-rows of W");
    VS_REQUIRE(p->rope_pos == nullptr && p->out_gin == 0, "vs_gemm: a_mode 2 goes with a plain output mapping");
    g.tn = 1;
    g.a_rows = p->a_rows;
    g.a_groups = 1;
    g.tiles_per_group = ceil_div(p->a_rows, BM);
    m_tiles = g.tiles_per_group;
    ktot = p->K;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(p->a_rows), static_cast<cuuint64_t>(p->K)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(p->a_row_stride) * 2};
    cuuint32_t box[2] = {64, BK};
    int rc = encode_map(&tmA, p->A, 2, dims, str, box);
    if (rc) return rc;
  } else if (p->a_mode == 3) {
    // conv wgrad: A = dY NHWC [cn, ch, cw, a_rows], W = X (the conv's input view); K = pixels
    VS_REQUIRE(p->a_rows > 0 && p->cn > 0 && p->ch > 0 && p->cw > 0 && p->cin > 0, "vs_gemm: empty conv wgrad");
    VS_REQUIRE(p->cin % 8 == 0 && p->a_row_stride % 8 == 0 && p->a_row_stride >= p->a_rows,
               "vs_gemm: conv wgrad: cin and the dY pixel stride must be multiples of 8");
    VS_REQUIRE(p->kh > 0 && p->kw > 0 && p->pad >= 0, "vs_gemm: bad conv kernel");
    VS_REQUIRE(p->rope_pos == nullptr && p->out_gin == 0, "vs_gemm: a_mode 3 goes with a plain output mapping");
    g.tn = 1;
    g.wg = 1;
    g.a_rows = p->a_rows;
    g.a_groups = 1;
    g.tiles_per_group = ceil_div(p->a_rows, BM);
    m_tiles = g.tiles_per_group;
    g.cn = p->cn;
    g.ch = p->ch;
    g.cw = p->cw;
    g.kw = p->kw;
    g.pad = p->pad;
    g.cin_pad = ceil_div(p->cin, BK) * BK;
    VS_REQUIRE(p->N == p->kh * p->kw * g.cin_pad, "vs_gemm: conv wgrad: N must be kh * kw * round_up(cin, 64)");
    int bw = 1;
    while (bw < 8 && bw < p->cw) bw <<= 1;
    int bh = 1;
    while (bw * bh < BK && bh < p->ch) bh <<= 1;
    const int bn = BK / (bw * bh);
    g.bw = bw;
    g.bh = bh;
    g.bn = bn;
    g.tiles_x = ceil_div(p->cw, bw);
    g.tiles_y = ceil_div(p->ch, bh);
    ktot = static_cast<long long>(g.tiles_x) * g.tiles_y * ceil_div(p->cn, bn) * BK;
    {
      cuuint64_t dims[4] = {static_cast<cuuint64_t>(p->a_rows), static_cast<cuuint64_t>(p->cw),
                            static_cast<cuuint64_t>(p->ch), static_cast<cuuint64_t>(p->cn)};
      const cuuint64_t sx = static_cast<cuuint64_t>(p->a_row_stride);
      cuuint64_t str[3] = {sx * 2, sx * p->cw * 2, sx * p->cw * p->ch * 2};
      cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh),
                           static_cast<cuuint32_t>(bn)};
      int rc = encode_map(&tmA, p->A, 4, dims, str, box);
      if (rc) return rc;
    }
    {
      const long long in_h = p->conv_in_h > 0 ? p->conv_in_h : p->ch;
      const long long sx = p->conv_stride_x > 0 ? p->conv_stride_x : p->cin;
      const long long sy = p->conv_stride_y > 0 ? p->conv_stride_y : sx * p->cw;
      const long long sn = p->conv_stride_n > 0 ? p->conv_stride_n : sy * in_h;
      VS_REQUIRE(sx % 8 == 0 && sy % 8 == 0 && sn % 8 == 0,
                 "vs_gemm: conv strides must be multiples of 8 elements");
      cuuint64_t dims[4] = {static_cast<cuuint64_t>(p->cin), static_cast<cuuint64_t>(p->cw),
                            static_cast<cuuint64_t>(in_h), static_cast<cuuint64_t>(p->cn)};
      cuuint64_t str[3] = {static_cast<cuuint64_t>(sx) * 2, static_cast<cuuint64_t>(sy) * 2,
                           static_cast<cuuint64_t>(sn) * 2};
      cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh),
                           static_cast<cuuint32_t>(bn)};
      int rc = encode_map(&tmW, p->W, 4, dims, str, box);
      if (rc) return rc;
    }
  } else {
    set_error("vs_gemm: unknown a_mode %d", p->a_mode);
    return VS_ERR_INVALID;
  }
  g.num_kb = static_cast<int>((ktot + BK - 1) / BK);
  g.m_tiles = m_tiles;
  VS_REQUIRE(static_cast<long long>(m_tiles) * BM * (p->out_gin > 0 ? 2 : 1) < (1ll << 30),
             "vs_gemm: more than 2^30 output rows");

  int bn = p->block_n;
  if (bn == 0 && p->c_accumulate) {
    // weight-gradient form: split-K keeps every SM busy whatever the number of output tiles, so the widest
    // tile wins (bytes of operand per MMA cycle: 256-wide tiles need 64 B/clk/SM, 128-wide 96 B/clk/SM --
    // ncu on the 3x3 wgrad at 256^2: tensor pipe 25 % active with BN = 128, L2 -> SM feed-bound)
    bn = p->N > 128 ? 256 : (p->N > 64 ? 128 : 64);
  }
  if (bn == 0) {
    double best = 0;
    for (int cand : {256, 128, 64}) {
      if (cand > 64 && cand / 2 >= p->N) continue;  // pure padding
      const double c = tile_cost(cand, m_tiles, p->N, g.num_kb);
      if (bn == 0 || c < best) { bn = cand; best = c; }
    }
  }
  VS_REQUIRE(bn == 64 || bn == 128 || bn == 256, "vs_gemm: block_n must be 64, 128 or 256");
  g.n_tiles = ceil_div(p->N, bn);
  // vertically adjacent tiles are computed by a CTA pair that splits the W tile (see header)
  const int cl = (bn >= 128 && m_tiles >= 2) ? 2 : 1;
  if (g.atomic) {
    // split-K: the wgrad GEMMs have few output tiles and a long K (tokens / pixels)
    const int units = ceil_div(m_tiles, cl) * g.n_tiles;
    const int slots = num_sms() / cl;
    int splits = p->split_k;
    if (splits <= 0) splits = units >= slots ? 1 : (2 * slots) / units;
    splits = std::max(1, std::min(splits, g.num_kb / 4 > 0 ? g.num_kb / 4 : 1));
    g.kb_per_split = ceil_div(g.num_kb, splits);
    g.splits = ceil_div(g.num_kb, g.kb_per_split);
  }
  if (g.wg) {
    // tmW was built with the A map above
  } else if (g.tn) {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(p->N), static_cast<cuuint64_t>(ktot)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(p->w_row_stride) * 2};
    cuuint32_t box[2] = {64, BK};
    int rc = encode_map(&tmW, p->W, 2, dims, str, box);
    if (rc) return rc;
  } else {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(p->N)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(p->w_row_stride) * 2};
    cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(bn / cl)};
    int rc = encode_map(&tmW, p->W, 2, dims, str, box);
    if (rc) return rc;
  }
  // the forward GEMMs run a kernel without the training features (gemm_kernel.cuh: 5 % faster for it)
  const bool train = g.tn || g.wg || g.atomic || g.mask_mode != 0 || g.splits > 1;
  VS_REQUIRE(!(train && g.f16), "vs_gemm: the fp16 operand format is a forward-path option");
  if (train) return vs::gemm_launch_train(bn, cl, tmA, tmW, g, stream);
  return g.f16 ? vs::gemm_launch_fwd16(bn, cl, tmA, tmW, g, stream) : vs::gemm_launch_fwd(bn, cl, tmA, tmW, g, stream);
}
