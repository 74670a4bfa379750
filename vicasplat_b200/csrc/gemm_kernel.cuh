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
// This file is the kernel; it is compiled three times (VS_GEMM_VARIANT 0 / 1 / 2, see gemm_tc05_kernel).
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cuda_bf16.h>

#include "common.h"
#include "gemm_dev.h"
#include "half16.cuh"
#include "ptx.cuh"
#include "tmap.h"

#ifndef VS_GEMM_VARIANT
#error "define VS_GEMM_VARIANT (0 forward bf16, 1 forward fp16, 2 training) before including gemm_kernel.cuh"
#endif

namespace vs {
namespace {

#ifdef VS_EPI_TIMING
__device__ long long g_epi_stamps[16];
#define EPI_STAMP(i) do { if (dbg_on) g_epi_stamps[i] = clock64(); } while (0)
#else
#define EPI_STAMP(i) do { } while (0)
#endif


// exact-erf GELU (nn.GELU default, croco/blocks.py:60) with erf from Abramowitz-Stegun 7.1.26
// (|err| <= 1.5e-7, far below the bf16 rounding of the result): 2 MUFU + ~8 FMA-pipe issue slots per
// value instead of erff's ~30 -- the GELU epilogue otherwise out-lasts the K = 1024 main loop.
// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 on sm_100): one issue slot for two lanes of math.
// The epilogue is issue-bound (8 warps per SM, long dependent chains), not FMA-pipe bound.
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// exact-erf GELU of two values: same Abramowitz-Stegun 7.1.26 polynomial as gelu_erf, rearranged as
//   gelu(x) = max(x, 0) - 0.5 |x| * (p(t) t) * exp(-z^2),  z = |x| / sqrt(2), t = 1 / (1 + 0.3275911 z)
// (erf(z) = 1 - p(t) t exp(-z^2) for z >= 0), evaluated with packed FFMA2 / FMUL2: 20 issue slots per
// pair instead of 2 x 16.
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const uint64_t AX = pk2(fabsf(x0), fabsf(x1));
  const uint64_t Z = mul2(AX, pk2(0.70710678118654752f, 0.70710678118654752f));
  const uint64_t DEN = fma2(pk2(0.3275911f, 0.3275911f), Z, pk2(1.0f, 1.0f));
  const uint64_t M = mul2(mul2(Z, Z), pk2(-1.4426950408889634f, -1.4426950408889634f));
  float d0, d1, m0, m1, t0, t1, e0, e1;
  upk2(DEN, d0, d1);
  upk2(M, m0, m1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(m0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(m1));
  const uint64_t T = pk2(t0, t1);
  uint64_t P = fma2(pk2(1.061405429f, 1.061405429f), T, pk2(-1.453152027f, -1.453152027f));
  P = fma2(P, T, pk2(1.421413741f, 1.421413741f));
  P = fma2(P, T, pk2(-0.284496736f, -0.284496736f));
  P = fma2(P, T, pk2(0.254829592f, 0.254829592f));
  const uint64_t R = mul2(mul2(P, T), pk2(e0, e1));
  const uint64_t H = mul2(AX, pk2(-0.5f, -0.5f));
  const uint64_t O = fma2(H, R, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
  upk2(O, x0, x1);
}

struct TileCoord {
  int n0;
  int grp, r0;          // rows mode
  int x0, y0, img0;     // conv mode
  int kb0, kb1;         // k-block range of this unit (split-K)
};

// work unit u of a cluster -> tile of CTA `rank`: units enumerate (n tile, group of CL m tiles),
// m fastest so that concurrently running clusters share W tiles.  mt may be >= m_tiles for the
// last group when m_tiles is odd: such a tile loads zeros and stores nothing.
template <bool TRAIN>
__device__ __forceinline__ TileCoord tile_coord(const GemmDev& g, int u, int rank, int CL, int BN) {
  TileCoord c{};
  const int m_groups = (g.m_tiles + CL - 1) / CL;
  c.kb0 = 0;
  c.kb1 = g.num_kb;
  if (TRAIN && g.splits > 1) {
    const int per = m_groups * g.n_tiles;
    const int sp = u / per;
    u -= sp * per;
    c.kb0 = sp * g.kb_per_split;
    c.kb1 = min(g.num_kb, c.kb0 + g.kb_per_split);
  }
  const int nt = u / m_groups;
  const int mt = (u - nt * m_groups) * CL + rank;
  c.n0 = nt * BN;
  if (g.mode == 0) {
    c.grp = mt / g.tiles_per_group;
    c.r0 = (mt - c.grp * g.tiles_per_group) * BM;
  } else {
    const int tx = mt % g.tiles_x;
    const int ty = (mt / g.tiles_x) % g.tiles_y;
    const int tn = mt / (g.tiles_x * g.tiles_y);
    c.x0 = tx * g.bw;
    c.y0 = ty * g.bh;
    c.img0 = tn * g.bn;
  }
  return c;
}

__device__ __forceinline__ void load32_f32(const float* p, int nv, bool vec, float (&o)[32]) {
  if (vec && nv == 32) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
      o[4 * i] = t.x; o[4 * i + 1] = t.y; o[4 * i + 2] = t.z; o[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = i < nv ? __ldg(p + i) : 0.f;
  }
}

// ---- 4-column segment helpers of the transposed epilogue (one lane = 4 consecutive columns of a row)
// plain (coherent) loads: the residual stream may be updated in place by this very kernel
__device__ __forceinline__ void add4_res(const void* base, int dtype, long long off, int nv, bool vec,
                                         float4& f, bool f16 = false) {
  if (dtype == VS_F32) {
    const float* p = static_cast<const float*>(base) + off;
    if (vec && nv == 4) {
      const float4 t = *reinterpret_cast<const float4*>(p);
      f.x += t.x; f.y += t.y; f.z += t.z; f.w += t.w;
    } else {
      if (nv > 0) f.x += p[0];
      if (nv > 1) f.y += p[1];
      if (nv > 2) f.z += p[2];
      if (nv > 3) f.w += p[3];
    }
  } else {
    const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(base) + off;
    if (vec && nv == 4) {
      const uint2 t = *reinterpret_cast<const uint2*>(p);
      const float2 a = h2_to_f2(t.x, f16);
      const float2 b = h2_to_f2(t.y, f16);
      f.x += a.x; f.y += a.y; f.z += b.x; f.w += b.y;
    } else {
      const uint16_t* q = reinterpret_cast<const uint16_t*>(p);
      if (nv > 0) f.x += h_to_f(q[0], f16);
      if (nv > 1) f.y += h_to_f(q[1], f16);
      if (nv > 2) f.z += h_to_f(q[2], f16);
      if (nv > 3) f.w += h_to_f(q[3], f16);
    }
  }
}

// += bilinear x2 (align_corners=True) sample of a half-resolution NHWC bf16 map at output pixel
// (x, y) of image im, channels [col, col+4)
__device__ __forceinline__ void add4_res_up2(const __nv_bfloat16* base, long long ld, int im, int x,
                                             int y, int ch, int cw, int col, int nv, bool vec,
                                             float4& f, bool f16 = false) {
  const int h = ch >> 1, w = cw >> 1;
  const float sy = ch > 1 ? static_cast<float>(h - 1) / (ch - 1) : 0.f;
  const float sx = cw > 1 ? static_cast<float>(w - 1) / (cw - 1) : 0.f;
  const float fy = y * sy, fx = x * sx;
  const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float wt[4] = {(1 - ly) * (1 - lx), (1 - ly) * lx, ly * (1 - lx), ly * lx};
  const int ys[4] = {y0, y0, y1, y1}, xs[4] = {x0, x1, x0, x1};
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float4 tmp = make_float4(0.f, 0.f, 0.f, 0.f);
    add4_res(base, VS_BF16, ((static_cast<long long>(im) * h + ys[t]) * w + xs[t]) * ld + col, nv, vec,
             tmp, f16);
    acc.x = fmaf(wt[t], tmp.x, acc.x); acc.y = fmaf(wt[t], tmp.y, acc.y);
    acc.z = fmaf(wt[t], tmp.z, acc.z); acc.w = fmaf(wt[t], tmp.w, acc.w);
  }
  // the stand-alone kernel rounds the upsampled map to bf16 before it is consumed: do the same
  f.x += round_h(acc.x, f16); f.y += round_h(acc.y, f16);
  f.z += round_h(acc.z, f16); f.w += round_h(acc.w, f16);
}

__device__ __forceinline__ void store4_bf16(__nv_bfloat16* p, int nv, bool vec, float4 f, bool relu,
                                            bool f16 = false) {
  if (relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
  if (vec && nv == 4) {
    *reinterpret_cast<uint2*>(p) = make_uint2(f2_to_h2(f.x, f.y, f16), f2_to_h2(f.z, f.w, f16));
  } else {
    uint16_t* q = reinterpret_cast<uint16_t*>(p);
    if (nv > 0) q[0] = f_to_h(f.x, f16);
    if (nv > 1) q[1] = f_to_h(f.y, f16);
    if (nv > 2) q[2] = f_to_h(f.z, f16);
    if (nv > 3) q[3] = f_to_h(f.w, f16);
  }
}

__device__ __forceinline__ void store4_f32(float* p, int nv, bool vec, const float4 f) {
  if (vec && nv == 4) {
    *reinterpret_cast<float4*>(p) = f;
  } else {
    if (nv > 0) p[0] = f.x;
    if (nv > 1) p[1] = f.y;
    if (nv > 2) p[2] = f.z;
    if (nv > 3) p[3] = f.w;
  }
}

__device__ __forceinline__ void add_bf16x4(float4& f, const uint2 t, bool f16 = false) {
  const float2 a = h2_to_f2(t.x, f16);
  const float2 b = h2_to_f2(t.y, f16);
  f.x += a.x; f.y += a.y; f.z += b.x; f.w += b.y;
}
__device__ __forceinline__ void fma_bf16x4(float4& f, const float w, const uint2 t, bool f16 = false) {
  const float2 a = h2_to_f2(t.x, f16);
  const float2 b = h2_to_f2(t.y, f16);
  f.x = fmaf(w, a.x, f.x); f.y = fmaf(w, a.y, f.y); f.z = fmaf(w, b.x, f.z); f.w = fmaf(w, b.y, f.w);
}

// ReLU mask from the kept (post-ReLU, bf16) output: v = m > 0 ? v : 0
__device__ __forceinline__ void mask_bf16x4(float4& f, const uint2 t) {
  // positive <=> sign bit clear and magnitude non-zero
  if (((t.x & 0x8000u) != 0u) || ((t.x & 0x7fffu) == 0u)) f.x = 0.f;
  if (((t.x & 0x80000000u) != 0u) || ((t.x & 0x7fff0000u) == 0u)) f.y = 0.f;
  if (((t.y & 0x8000u) != 0u) || ((t.y & 0x7fffu) == 0u)) f.z = 0.f;
  if (((t.y & 0x80000000u) != 0u) || ((t.y & 0x7fff0000u) == 0u)) f.w = 0.f;
}
__device__ __forceinline__ void red_add_f4(float* p, const float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

struct Up2Taps { int x0, x1, y0, y1; float lx, ly; };
// bilinear x2, align_corners=True: source taps / weights of output pixel (x, y) of a ch x cw map
__device__ __forceinline__ Up2Taps up2_taps(int x, int y, int ch, int cw, float sx, float sy) {
  const int h = ch >> 1, w = cw >> 1;
  const float fy = y * sy, fx = x * sx;
  Up2Taps t;
  t.y0 = static_cast<int>(fy); t.x0 = static_cast<int>(fx);
  t.y1 = min(t.y0 + 1, h - 1); t.x1 = min(t.x0 + 1, w - 1);
  t.ly = fy - t.y0; t.lx = fx - t.x0;
  return t;
}

// Residual segments of one chunk, global -> this warp's 4 KB shared buffer with cp.async (lane l,
// row-group j: 16 bytes at j * 512 + l * 16; kind 1: 4 fp32; kind 2: 4 bf16 of res1 then 4 bf16 of
// res2 or zeros).  NOT through registers: an outstanding LDG is waited for by the next
// tcgen05.wait::ld (same scoreboard -- in-kernel stamps showed ~2 000 clk there per chunk), an
// async copy is not.  Rows without output read row 0 and are never stored.  Plain (coherent)
// global reads: the residual stream may be updated in place by this very kernel.
__device__ __forceinline__ void res_async_issue(const GemmDev& g, int rk, const int (&oj)[8], int col,
                                                uint32_t res_s, int lane) {
  const uint32_t dst = res_s + lane * 16;
  if (rk == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* src = static_cast<const float*>(g.res1) + col + static_cast<long long>(max(oj[j], 0)) * g.res_ld;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + j * 512), "l"(src) : "memory");
    }
  } else if (rk == 2) {
    const __nv_bfloat16* r2 = static_cast<const __nv_bfloat16*>(g.res2 != nullptr ? g.res2 : g.res1);
    const int n2 = g.res2 != nullptr ? 8 : 0;   // src-size 0: the 8 bytes are zero-filled
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long off = static_cast<long long>(max(oj[j], 0)) * g.res_ld;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + j * 512),
                   "l"(static_cast<const __nv_bfloat16*>(g.res1) + col + off) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + j * 512 + 8),
                   "l"(r2 + col + off), "r"(n2) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void res_async_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_shared_u4(uint32_t addr) {
  uint4 t;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(addr));
  return t;
}

__device__ __forceinline__ void st_shared_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 t;
  // volatile (ordered against the st.shared / __syncwarp around it) but NO memory clobber: the global
  // stores that consume these values must stay free to be scheduled behind all eight loads --
  // with the clobber every row was a serial LDS -> STG round trip (~120 clk each, in-kernel stamps)
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(addr));
  return t;
}

// Fast store phase of one 32 x 32 chunk (aligned pointers, full chunk): lane l owns columns
// [col, col+4) of rows 4j + l/8.  RK = residual kind (4 / 5: the bf16 kind with mask_mode 1 / 2),
// CF32 = fp32 output.
// F16 / ATOMIC are compile-time: a run-time test inside the row loops turns every row's LDS -> STG sequence
// into its own branch region (measured: +1.9 ms per 8-scene encoder pass).
template <int RK, bool CF32, bool F16 = false, bool ATOMIC = false>
__device__ __forceinline__ void epi_store(const GemmDev& g, const uint32_t (&lds_base)[2], int lane,
                                          const int (&oj)[8], int col, const uint4 (&rb)[8], int pix_x,
                                          int pix_y, int pix_im, bool all_rows) {
  // per chunk: column-adjusted base pointers; per row one 32 x 32 -> 64 multiply-add each
  float* const cf = static_cast<float*>(g.C) + col;
  __nv_bfloat16* const cb = static_cast<__nv_bfloat16*>(g.C) + col;
  __nv_bfloat16* const c2 = g.C2 != nullptr ? g.C2 + col : nullptr;
#pragma unroll
  for (int jh = 0; jh < 8; jh += 4) {
    uint2 up[RK == 3 ? 4 : 1][4];
    float lxs[4], lys[4];
    if (RK == 3) {   // 16 tap loads in flight per half
      const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(g.res1) + col;
      const int h2 = g.ch >> 1, w2 = g.cw >> 1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = 4 * (jh + j) + (lane >> 3);
        const int px = __shfl_sync(0xffffffffu, pix_x, rr);
        const int py = __shfl_sync(0xffffffffu, pix_y, rr);
        const int pim = __shfl_sync(0xffffffffu, pix_im, rr);
        const Up2Taps tp = up2_taps(px, py, g.ch, g.cw, g.up_sx, g.up_sy);
        lxs[j] = tp.lx; lys[j] = tp.ly;
#pragma unroll
        for (int t = 0; t < 4; ++t) {   // pix_* are clamped to the map: unconditional loads
          const int yy = (t & 2) ? tp.y1 : tp.y0, xx = (t & 1) ? tp.x1 : tp.x0;
          up[RK == 3 ? j : 0][t] =
              *reinterpret_cast<const uint2*>(base + static_cast<long long>((pim * h2 + yy) * w2 + xx) * g.res_ld);
        }
      }
    }
    float4 tv[4];
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      const int j = jh + j4;
      float4 t = ld_shared_f4(lds_base[j & 1] + j * 512);
      if (RK == 1) {
        t.x += __uint_as_float(rb[j].x); t.y += __uint_as_float(rb[j].y);
        t.z += __uint_as_float(rb[j].z); t.w += __uint_as_float(rb[j].w);
      } else if (RK == 2) {
        add_bf16x4(t, make_uint2(rb[j].x, rb[j].y), F16);
        add_bf16x4(t, make_uint2(rb[j].z, rb[j].w), F16);
      } else if (RK == 4) {   // mask_mode 1: res2 masks, then + res1
        t.x *= g.out_scale; t.y *= g.out_scale; t.z *= g.out_scale; t.w *= g.out_scale;
        mask_bf16x4(t, make_uint2(rb[j].z, rb[j].w));
        add_bf16x4(t, make_uint2(rb[j].x, rb[j].y), false);
      } else if (RK == 5) {   // mask_mode 2: res1 masks
        t.x *= g.out_scale; t.y *= g.out_scale; t.z *= g.out_scale; t.w *= g.out_scale;
        mask_bf16x4(t, make_uint2(rb[j].x, rb[j].y));
      } else if (RK == 3) {
        const float lx = lxs[j4], ly = lys[j4];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        fma_bf16x4(acc, (1 - ly) * (1 - lx), up[RK == 3 ? j4 : 0][0], F16);
        fma_bf16x4(acc, (1 - ly) * lx, up[RK == 3 ? j4 : 0][1], F16);
        fma_bf16x4(acc, ly * (1 - lx), up[RK == 3 ? j4 : 0][2], F16);
        fma_bf16x4(acc, ly * lx, up[RK == 3 ? j4 : 0][3], F16);
        // the stand-alone kernel rounds the upsampled map to the 16-bit format before it is consumed
        t.x += round_h(acc.x, F16); t.y += round_h(acc.y, F16);
        t.z += round_h(acc.z, F16); t.w += round_h(acc.w, F16);
      }
      tv[j4] = t;
    }
    // Stores.  A per-row `if (row valid)` compiles to a branch region per row, which serialises
    // LDS -> STG round trips (~120 clk per row, in-kernel stamps); interior tiles (every row of
    // the warp valid, warp-uniform flag) therefore take the branch-free path.
    if (all_rows) {
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const long long o = oj[jh + j4];
        if (CF32) {
          if (ATOMIC) red_add_f4(cf + o * g.ldc, tv[j4]);
          else *reinterpret_cast<float4*>(cf + o * g.ldc) = tv[j4];
        } else {
          store4_bf16(cb + o * g.ldc, 4, true, tv[j4], false, F16);
        }
      }
      if (c2 != nullptr) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          store4_bf16(c2 + static_cast<long long>(oj[jh + j4]) * g.ldc2, 4, true, tv[j4], true, F16);
      }
    } else {
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const int j = jh + j4;
        if (oj[j] >= 0) {
          if (CF32) {
            if (ATOMIC) red_add_f4(cf + static_cast<long long>(oj[j]) * g.ldc, tv[j4]);
            else *reinterpret_cast<float4*>(cf + static_cast<long long>(oj[j]) * g.ldc) = tv[j4];
          } else {
            store4_bf16(cb + static_cast<long long>(oj[j]) * g.ldc, 4, true, tv[j4], false, F16);
          }
          if (c2 != nullptr) store4_bf16(c2 + static_cast<long long>(oj[j]) * g.ldc2, 4, true, tv[j4], true, F16);
        }
      }
    }
  }
}

// V = VS_GEMM_VARIANT of the translation unit: 0 forward / bf16, 1 forward / fp16, 2 training (bf16; adds
// MN-major operands, conv-wgrad addressing, split-K, the atomic and the ReLU-mask epilogues).  They are
// separate kernels because the forward GEMMs pay for every feature compiled into their kernel -- the training
// features cost them 5 % (29.9 -> 31.6 ms per 8-scene encoder pass) although none of their branches is
// ever taken (bisected: each feature alone 0.1 - 0.8 ms, no register or spill difference; code size 2x).
template <int BN, int STAGES, int EPI_WARPS, int CL, int V>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
    gemm_tc05_kernel(const __grid_constant__ CUtensorMap tmA,
                     const __grid_constant__ CUtensorMap tmW, const GemmDev g) {
  constexpr bool TRAIN = V == 2;
  constexpr bool F16 = V == 1;
  constexpr int A_BYTES = BM * 128;
  constexpr int B_BYTES = (BN / CL) * 128;   // a CTA pair splits the W tile
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = 2 * BN;  // two accumulators
  static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "epilogue warps");
  constexpr int NCHUNK = BN / 32;
  constexpr int CH_PER_WARP = NCHUNK / (EPI_WARPS / 4);
  static_assert(CH_PER_WARP >= 1, "too many epilogue warps for this tile width");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  // per-epilogue-warp 32 x 32 fp32 transposition tile (XOR-swizzled 16-byte slots, no padding)
  float4* epi_scratch = reinterpret_cast<float4*>(smem + STAGES * STAGE_BYTES + 256);
  // ... followed by one 4 KB residual staging buffer per epilogue warp (cp.async destination)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CL > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int total_units = ((g.m_tiles + CL - 1) / CL) * g.n_tiles * (TRAIN ? g.splits : 1);
  const int unit0 = blockIdx.x / CL, unit_step = gridDim.x / CL;
  constexpr uint16_t PAIR_MASK = 3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], EPI_WARPS * CL);   // pair: both CTAs' epilogues report to the leader
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CL == 1) {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    } else {
      tmem_alloc_pair(tmem_slot, TMEM_COLS);
      tmem_relinquish_pair();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // peer barriers are initialised before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above overlapped the tail of the previous kernel (PDL); global memory is touched
  // only from here on
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;  // k-block counter across tiles (ring position)
      for (int u = unit0; u < total_units; u += unit_step) {
        const TileCoord tc = tile_coord<TRAIN>(g, u, rank, CL, BN);
        for (int kb = tc.kb0; kb < tc.kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sA = smem + s * STAGE_BYTES;
          uint8_t* sB = sA + A_BYTES;
          if (TRAIN && g.wg) {
            // conv wgrad: the k-block is a 64-pixel box (bw x bh x bn) of both NHWC maps; every
            // 64-column block of the W tile is the input map shifted by that block's filter tap
            const int tx = kb % g.tiles_x;
            const int ty = (kb / g.tiles_x) % g.tiles_y;
            const int x0 = tx * g.bw, y0 = ty * g.bh, img0 = (kb / (g.tiles_x * g.tiles_y)) * g.bn;
            if (CL == 1) mbar_expect_tx(&full[s], STAGE_BYTES);
            else if (rank == 0) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) {
              if (CL == 1) tma_load_4d(sA + i * 8192, &tmA, &full[s], tc.r0 + i * 64, x0, y0, img0);
              else tma_load_4d_pair(sA + i * 8192, &tmA, &full[s], tc.r0 + i * 64, x0, y0, img0);
            }
#pragma unroll
            for (int i = 0; i < (BN / CL) / 64; ++i) {
              const int col = tc.n0 + rank * (BN / CL) + i * 64;
              const int tap = col / g.cin_pad;
              const int c0 = col - tap * g.cin_pad;
              const int dy = tap / g.kw, dx = tap - dy * g.kw;
              if (CL == 1)
                tma_load_4d(sB + i * 8192, &tmW, &full[s], c0, x0 + dx - g.pad, y0 + dy - g.pad, img0);
              else
                tma_load_4d_pair(sB + i * 8192, &tmW, &full[s], c0, x0 + dx - g.pad, y0 + dy - g.pad, img0);
            }
          } else if (TRAIN && g.tn) {
            // MN-major operands: one [64 k-rows][64 mn] box (8 KB) per 64-wide block of the tile
            if (CL == 1) {
              mbar_expect_tx(&full[s], STAGE_BYTES);
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                tma_load_2d(sA + i * 8192, &tmA, &full[s], tc.r0 + i * 64, kb * BK);
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_2d(sB + i * 8192, &tmW, &full[s], tc.n0 + i * 64, kb * BK);
            } else {
              if (rank == 0) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                tma_load_2d_pair(sA + i * 8192, &tmA, &full[s], tc.r0 + i * 64, kb * BK);
#pragma unroll
              for (int i = 0; i < (BN / CL) / 64; ++i)
                tma_load_2d_pair(sB + i * 8192, &tmW, &full[s], tc.n0 + rank * (BN / CL) + i * 64,
                                 kb * BK);
            }
          } else if (CL == 1) {
            mbar_expect_tx(&full[s], STAGE_BYTES);
            if (g.mode == 0) {
              tma_load_3d(sA, &tmA, &full[s], kb * BK, tc.r0, tc.grp);
            } else {
              const int tap = kb / g.cblocks;
              const int c0 = (kb - tap * g.cblocks) * BK;
              const int dy = tap / g.kw, dx = tap - dy * g.kw;
              tma_load_4d(sA, &tmA, &full[s], c0, tc.x0 + dx - g.pad, tc.y0 + dy - g.pad, tc.img0);
            }
            tma_load_2d(sB, &tmW, &full[s], kb * BK, tc.n0);
          } else {
            // both CTAs' bytes are counted on the leader's barrier (its MMA thread waits there)
            if (rank == 0) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);
            if (g.mode == 0) {
              tma_load_3d_pair(sA, &tmA, &full[s], kb * BK, tc.r0, tc.grp);
            } else {
              const int tap = kb / g.cblocks;
              const int c0 = (kb - tap * g.cblocks) * BK;
              const int dy = tap / g.kw, dx = tap - dy * g.kw;
              tma_load_4d_pair(sA, &tmA, &full[s], c0, tc.x0 + dx - g.pad, tc.y0 + dy - g.pad,
                               tc.img0);
            }
            tma_load_2d_pair(sB, &tmW, &full[s], kb * BK, tc.n0 + rank * (BN / CL));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (pair: leader only)
    if (lane == 0 && rank == 0) {
      // a / b format bits [7,10) / [10,13): 1 = bf16, 0 = fp16
      constexpr uint32_t fmt_clear = F16 ? ~((1u << 7) | (1u << 10)) : ~0u;
      const uint32_t idesc = umma_idesc_bf16(BM * CL, BN) & fmt_clear;
      const uint32_t idesc_tn = umma_idesc_bf16(BM * CL, BN, 1, 1) & fmt_clear;
      uint32_t it = 0, ti = 0;
      for (int u = unit0; u < total_units; u += unit_step, ++ti) {
        const uint32_t a = ti & 1;
        int kb_first = 0, kb_last = g.num_kb;
        if (TRAIN && g.splits > 1) {
          const TileCoord tcm = tile_coord<TRAIN>(g, u, 0, CL, BN);
          kb_first = tcm.kb0;
          kb_last = tcm.kb1;
        }
        mbar_wait(&tmem_empty[a], ((ti >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int kb = kb_first; kb < kb_last; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
          if (TRAIN && g.tn) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {   // 16 k-rows of 128 B per step
              const uint64_t da = umma_desc_mn_sw128_lbo(a_addr + k * 2048, 8192);
              const uint64_t db = umma_desc_mn_sw128_lbo(b_addr + k * 2048, 8192);
              if (CL == 1) umma_bf16_ss(d_tmem, da, db, idesc_tn, (kb != kb_first || k != 0) ? 1u : 0u);
              else umma_bf16_ss_pair(d_tmem, da, db, idesc_tn, (kb != kb_first || k != 0) ? 1u : 0u);
            }
          } else {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (CL == 1)
              umma_bf16_ss(d_tmem, umma_desc_k_sw128(a_addr + k * 32),
                           umma_desc_k_sw128(b_addr + k * 32), idesc, (kb != kb_first || k != 0) ? 1u : 0u);
            else
              umma_bf16_ss_pair(d_tmem, umma_desc_k_sw128(a_addr + k * 32),
                                umma_desc_k_sw128(b_addr + k * 32), idesc, (kb != kb_first || k != 0) ? 1u : 0u);
          }
          }
          // slot reusable (in both CTAs of a pair) once these MMAs have read it
          if (CL == 1) umma_commit(&empty[s]);
          else umma_commit_pair(&empty[s], PAIR_MASK);
        }
        if (CL == 1) umma_commit(&tmem_full[a]);
        else umma_commit_pair(&tmem_full[a], PAIR_MASK);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int ew = warp - 2;
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = ew >> 2;          // column half when 8 epilogue warps
    const int r = q * 32 + lane;       // tile row owned by this thread
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t tile_s = smem_u32(epi_scratch + ew * 256);   // 4 KB, 256-byte aligned
    const uint32_t res_s = smem_u32(epi_scratch + (EPI_WARPS + ew) * 256);
    // write side: row = lane, 16-byte slot i ^ (lane & 7)  ->  st_base ^ (i << 4)
    const uint32_t st_base = (tile_s + lane * 128) | ((lane & 7) << 4);
    // read side: row 4 j + (lane >> 3), slot (lane & 7) ^ (row & 7); row & 7 only depends on j & 1
    uint32_t lds_base[2];
#pragma unroll
    for (int pj = 0; pj < 2; ++pj) {
      const int rq = (lane >> 3) + 4 * pj;
      lds_base[pj] = tile_s + (lane >> 3) * 128 + (((lane & 7) ^ rq) << 4);
    }
    uint32_t ti = 0;
    for (int u = unit0; u < total_units; u += unit_step, ++ti) {
      const TileCoord tc = tile_coord<TRAIN>(g, u, rank, CL, BN);
      const uint32_t a = ti & 1;
      long long my_out = -1;
      int my_gate = -1;
      int pix_x = 0, pix_y = 0, pix_im = 0;
      {
        bool valid;
        long long m;
        if (g.mode == 0) {
          valid = (tc.r0 + r) < g.a_rows && tc.grp < g.a_groups;
          m = static_cast<long long>(tc.grp) * g.a_rows + tc.r0 + r;
        } else {
          const int x = tc.x0 + r % g.bw;
          const int y = tc.y0 + (r / g.bw) % g.bh;
          const int im = tc.img0 + r / (g.bw * g.bh);
          valid = x < g.cw && y < g.ch && im < g.cn;
          m = (static_cast<long long>(im) * g.ch + y) * g.cw + x;
          pix_x = min(x, g.cw - 1); pix_y = min(y, g.ch - 1); pix_im = min(im, g.cn - 1);
        }
        if (valid) {
          long long o = m;
          if (g.out_gin > 0) o = (m / g.out_gin) * g.out_gout + g.out_off + (m % g.out_gin);
          my_out = o;
          if (g.gate_rows > 0) {
            const bool first = (o % g.gate_rows) == 0;
            if (first && g.first_row_mode == 2) my_out = -1;
            if (g.gate != nullptr && !(first && g.first_row_mode != 0))
              my_gate = static_cast<int>(o / g.gate_rows);
          } else if (g.gate != nullptr) {
            my_gate = 0;
          }
        }
      }
      int rp_y = 0, rp_x = 0;
      if (g.rope_pos != nullptr && my_out >= 0) {
        const int2 pp = __ldg(reinterpret_cast<const int2*>(g.rope_pos) + my_out);
        rp_y = pp.x; rp_x = pp.y;
      }
      // After the shared-memory exchange lane l holds 4 consecutive columns (cs) of rows 4j + l/8.
      const int cs = lane & 7;
      int oj[8];      // output row (fits 31 bits: checked on the host), -1 = nothing to store
#pragma unroll
      for (int j = 0; j < 8; ++j) oj[j] = static_cast<int>(__shfl_sync(0xffffffffu, my_out, 4 * j + (lane >> 3)));
      const bool all_rows = __all_sync(0xffffffffu, my_out >= 0);   // interior tile: no row predicates
      const int rk = g.res_kind;   // 0 none, 1 f32, 2 bf16 (one or two maps), 3 bilinear-x2 bf16
      // residual segments are fetched ahead of their use (first chunk: before the accumulator is
      // even ready; later chunks: before the TMEM read of that chunk)
      {
        const int nb0 = tc.n0 + half * CH_PER_WARP * 32;
        if (g.fast && (rk == 1 || rk == 2) && nb0 + 32 <= g.N)
          res_async_issue(g, rk, oj, nb0 + 4 * cs, res_s, lane);
      }
#ifdef VS_EPI_TIMING
      const bool dbg_tile = blockIdx.x == 0 && ew == 0 && lane == 0 && ti == 2;
      bool dbg_on = dbg_tile;
#endif
      EPI_STAMP(0);
      mbar_wait(&tmem_full[a], (ti >> 1) & 1);
      tc_fence_after();
      EPI_STAMP(1);
#pragma unroll 1
      for (int cc = 0; cc < CH_PER_WARP; ++cc) {
#ifdef VS_EPI_TIMING
        dbg_on = dbg_tile && cc == 1;
#endif
        EPI_STAMP(2);
        const int c = half * CH_PER_WARP + cc;
        const int nb = tc.n0 + c * 32;
        if (nb >= g.N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(t_lane + a * BN + c * 32, v);
        const int nv = min(32, g.N - nb);
        const bool fast = g.fast && nv == 32;
        const int col = nb + 4 * cs;
        tmem_ld_wait();
        EPI_STAMP(3);
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if (g.bias != nullptr) {
          float b[32];
          load32_f32(g.bias + nb, nv, g.vec & VEC_BIAS, b);
#pragma unroll
          for (int i = 0; i < 32; i += 2) upk2(add2(pk2(f[i], f[i + 1]), pk2(b[i], b[i + 1])), f[i], f[i + 1]);
        }
        if (g.rope_pos != nullptr &&
            (static_cast<unsigned>(nb - g.rope_q0) < static_cast<unsigned>(g.rope_cols) ||
             static_cast<unsigned>(nb - g.rope_k0) < static_cast<unsigned>(g.rope_cols))) {
          // this chunk is one 32-wide half of a head: y block (even) or x block (odd)
          const int hb = (nb >> 5) & 1;
          if (rp_y >= 0) {   // image token: pairs (d, d + 16), angle = pos * base^(-d/16)
            const float pos = static_cast<float>(hb ? rp_x : rp_y);
#pragma unroll
            for (int d = 0; d < 16; ++d) {
              float sn, cs_;
              __sincosf(pos * g.rope_if_img[d], &sn, &cs_);
              const float u = f[d], w = f[d + 16];
              f[d] = u * cs_ - w * sn;
              f[d + 16] = w * cs_ + u * sn;
            }
          } else {           // camera token: interleaved pairs, frame index t = -1 - y
            const float t = static_cast<float>(-1 - rp_y);
#pragma unroll
            for (int l = 0; l < 16; ++l) {
              float sn, cs_;
              __sincosf(t * (hb ? g.rope_if_cam[16 + l] : g.rope_if_cam[l]), &sn, &cs_);
              const float u = f[2 * l], w = f[2 * l + 1];
              f[2 * l] = u * cs_ - w * sn;
              f[2 * l + 1] = w * cs_ + u * sn;
            }
          }
        }
        if (g.act == VS_ACT_GELU) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) gelu_erf2(f[i], f[i + 1]);
        } else if (g.act == VS_ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
        }
        if (my_gate >= 0) {
          float gt[32];
          load32_f32(g.gate + static_cast<long long>(my_gate) * g.gate_ld + nb, nv, g.vec & VEC_GATE, gt);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {   // f * (1 + gate) = f * gate + f
            const uint64_t F = pk2(f[i], f[i + 1]);
            upk2(fma2(F, pk2(gt[i], gt[i + 1]), F), f[i], f[i + 1]);
          }
        }
        // Transpose through shared memory: a thread owns a ROW here (TMEM lane), but row-per-thread
        // global accesses cost one LSU wavefront per lane; afterwards a warp instruction touches
        // 4 rows x 128 B.
#pragma unroll
        for (int i = 0; i < 8; ++i)
          st_shared_f4(st_base ^ (i << 4), f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        EPI_STAMP(4);
        __syncwarp();
        EPI_STAMP(5);
        if (fast) {
          uint4 rb[8];
          if (rk == 1 || rk == 2) {
            res_async_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) rb[j] = ld_shared_u4(res_s + j * 512 + lane * 16);
            // the buffer is in registers: refill it for the next chunk now, a whole store phase
            // plus the next chunk's TMEM read / bias / transposition ahead of its use
            if (cc + 1 < CH_PER_WARP && nb + 64 <= g.N) res_async_issue(g, rk, oj, col + 32, res_s, lane);
          }
          const bool cf32 = g.c_dtype == VS_F32;
#define VS_EPI(RKV, PX, PY, PIM)                                                                    \
  do {                                                                                               \
    if (cf32) epi_store<RKV, true, F16>(g, lds_base, lane, oj, col, rb, PX, PY, PIM, all_rows);      \
    else epi_store<RKV, false, F16>(g, lds_base, lane, oj, col, rb, PX, PY, PIM, all_rows);          \
  } while (0)
          if (TRAIN && g.atomic) {   // weight gradients: fp32, no residual (checked on the host)
            epi_store<0, true, false, true>(g, lds_base, lane, oj, col, rb, 0, 0, 0, all_rows);
          } else {
            // 4 / 5: the bf16 kind with mask_mode 1 / 2 (backward passes)
            switch (TRAIN && rk == 2 && g.mask_mode != 0 ? 3 + g.mask_mode : rk) {
              case 0: VS_EPI(0, 0, 0, 0); break;
              case 1: VS_EPI(1, 0, 0, 0); break;
              case 2: VS_EPI(2, 0, 0, 0); break;
              case 3: VS_EPI(3, pix_x, pix_y, pix_im); break;
              case 4: if constexpr (TRAIN) VS_EPI(4, 0, 0, 0); break;
              default: if constexpr (TRAIN) VS_EPI(5, 0, 0, 0); break;
            }
          }
#undef VS_EPI
        } else {
          const int nvl = min(4, g.N - col);
#pragma unroll 1
          for (int j = 0; j < 8; ++j) {
            const int rr = 4 * j + (lane >> 3);
            float4 t = ld_shared_f4(tile_s + (rr * 8 + (cs ^ (rr & 7))) * 16);
            const long long o = __shfl_sync(0xffffffffu, my_out, rr);
            int px = 0, py = 0, pim = 0;
            if (g.res_up2) {
              px = __shfl_sync(0xffffffffu, pix_x, rr);
              py = __shfl_sync(0xffffffffu, pix_y, rr);
              pim = __shfl_sync(0xffffffffu, pix_im, rr);
            }
            if (o < 0 || nvl <= 0) continue;
            if (g.res1 != nullptr && g.res_up2) {
              add4_res_up2(static_cast<const __nv_bfloat16*>(g.res1), g.res_ld, pim, px, py, g.ch, g.cw,
                           col, nvl, g.vec & VEC_RES, t, F16);
            } else if (TRAIN && g.res1 != nullptr && g.mask_mode != 0) {
              // ReLU mask (bf16 maps): mode 1 masks with res2 and adds res1, mode 2 masks with res1
              const __nv_bfloat16* mk = static_cast<const __nv_bfloat16*>(g.mask_mode == 1 ? g.res2 : g.res1) +
                                        o * g.res_ld + col;
              const uint16_t* mq = reinterpret_cast<const uint16_t*>(mk);
              t.x *= g.out_scale; t.y *= g.out_scale; t.z *= g.out_scale; t.w *= g.out_scale;
              if (nvl > 0 && !(h_to_f(mq[0], F16) > 0.f)) t.x = 0.f;
              if (nvl > 1 && !(h_to_f(mq[1], F16) > 0.f)) t.y = 0.f;
              if (nvl > 2 && !(h_to_f(mq[2], F16) > 0.f)) t.z = 0.f;
              if (nvl > 3 && !(h_to_f(mq[3], F16) > 0.f)) t.w = 0.f;
              if (g.mask_mode == 1) add4_res(g.res1, g.res_dtype, o * g.res_ld + col, nvl, false, t, F16);
            } else if (g.res1 != nullptr) {
              add4_res(g.res1, g.res_dtype, o * g.res_ld + col, nvl, g.vec & VEC_RES, t, F16);
              if (g.res2 != nullptr)
                add4_res(g.res2, g.res_dtype, o * g.res_ld + col, nvl, g.vec & VEC_RES, t, F16);
            }
            if (TRAIN && g.C != nullptr && g.atomic) {
              float* cp = static_cast<float*>(g.C) + o * g.ldc + col;
              if (nvl > 0) atomicAdd(cp + 0, t.x);
              if (nvl > 1) atomicAdd(cp + 1, t.y);
              if (nvl > 2) atomicAdd(cp + 2, t.z);
              if (nvl > 3) atomicAdd(cp + 3, t.w);
            } else if (g.C != nullptr) {
              if (g.c_dtype == VS_F32)
                store4_f32(static_cast<float*>(g.C) + o * g.ldc + col, nvl, g.vec & VEC_C, t);
              else
                store4_bf16(static_cast<__nv_bfloat16*>(g.C) + o * g.ldc + col, nvl, g.vec & VEC_C, t, false, F16);
            }
            if (g.C2 != nullptr) store4_bf16(g.C2 + o * g.ldc2 + col, nvl, g.vec & VEC_C2, t, true, F16);
          }
        }
        EPI_STAMP(6);
        __syncwarp();   // the tile is rewritten by the next chunk
        EPI_STAMP(7);
      }
#ifdef VS_EPI_TIMING
      dbg_on = dbg_tile;
#endif
      EPI_STAMP(8);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL == 1 || rank == 0) mbar_arrive(&tmem_empty[a]);
        else mbar_arrive_cluster(&tmem_empty[a], 0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // no remote arrive / pair MMA may target a CTA that has exited
  if (warp == 1) {
    if (CL == 1) tmem_dealloc(tmem_base, TMEM_COLS);
    else tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

bool pdl_enabled() {   // VS_PDL=0 turns programmatic dependent launch off (A/B measurements)
  static const bool on = []() {
    const char* e = getenv("VS_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

inline int num_sms() { return gemm_num_sms(); }

template <int BN, int STAGES, int EPI_WARPS, int CL, int V>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmDev& g, cudaStream_t stream) {
  constexpr int SMEM = STAGES * (BM * 128 + (BN / CL) * 128) + 1024 /*align*/ + 256 /*barriers*/ +
                       2 * EPI_WARPS * 4096 /*epilogue transposition tiles + residual staging*/;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  auto kernel = gemm_tc05_kernel<BN, STAGES, EPI_WARPS, CL, V>;
    VS_CONFIGURE_PER_DEVICE(
    VS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  );
  const int units = ceil_div(g.m_tiles, CL) * g.n_tiles * g.splits;
  const int slots = num_sms() / CL;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>((units < slots ? units : slots) * CL));
  cfg.blockDim = dim3(64 + 32 * EPI_WARPS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  VS_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmA, tmW, g));
  count_launch();
  return VS_OK;
}

template <int V>
int dispatch(int bn, int cl, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmDev& g, cudaStream_t stream) {
  switch (bn) {
    case 64: return launch<64, 8, 4, 1, V>(tmA, tmW, g, stream);
    case 128:
      return cl == 2 ? launch<128, 8, 4, 2, V>(tmA, tmW, g, stream)
                     : launch<128, 6, 4, 1, V>(tmA, tmW, g, stream);
    default:
      return cl == 2 ? launch<256, 5, 8, 2, V>(tmA, tmW, g, stream)
                     : launch<256, 3, 8, 1, V>(tmA, tmW, g, stream);
  }
}

}  // namespace

#if VS_GEMM_VARIANT == 0
int gemm_launch_fwd(int bn, int cl, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmDev& g, cudaStream_t stream) {
  return dispatch<0>(bn, cl, tmA, tmW, g, stream);
}
#elif VS_GEMM_VARIANT == 1
int gemm_launch_fwd16(int bn, int cl, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmDev& g, cudaStream_t stream) {
  return dispatch<1>(bn, cl, tmA, tmW, g, stream);
}
#else
int gemm_launch_train(int bn, int cl, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmDev& g, cudaStream_t stream) {
  return dispatch<2>(bn, cl, tmA, tmW, g, stream);
}
#endif

}  // namespace vs

#if defined(VS_EPI_TIMING) && VS_GEMM_VARIANT == 0
extern "C" int vs_debug_epi_stamps(long long* out) {
  return cudaMemcpyFromSymbol(out, vs::g_epi_stamps, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}
#endif
