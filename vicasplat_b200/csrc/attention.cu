// Flash-style attention for head_dim 64 on tcgen05 (sm_100a).
//
// One CTA = one (item, head, 128-query tile).  Warp roles (320 threads):
//   warp 0      TMA producer: Q tile once, then K and V tiles (64 keys x 64) through 3-deep rings
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer:
//                 S[128x64] = Q K^T    (both operands K-major, 128B-swizzled, straight from TMA)
//                 O'[128x64] = P V     (P written by the softmax warps as a K-major swizzled tile,
//                                       V consumed as an MN-major operand: no transpose anywhere)
//   warps 2..9  softmax: two threads share query row r (TMEM lane r), each owning half of the key
//               columns and half of the output columns: two passes over S in TMEM (row max, then
//               exp2 / row sum / bf16 P -> shared memory), then the fresh O' tile is folded into the
//               fp32 register accumulator with the running-max correction.
// The exponentials (MUFU) bound this kernel at hd = 64, not the tensor pipe, so S is single
// buffered and two CTAs per SM interleave their softmax and MMA phases.
//
// Replaces: explicit softmax attention croco/blocks.py:105-109; F.scaled_dot_product_attention in
// VideoCameraAttention / CrossNeighborAttention, backbone_vica.py:116-121,188; the blocked-causal
// camera mask backbone_vica.py:585-593.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.h"
#include "half16.cuh"
#include "ptx.cuh"
#include "tmap.h"

namespace vs {
namespace {

constexpr int QT = 128;   // queries per CTA
constexpr int KT = 64;    // keys per step
constexpr int HD = 64;
constexpr int KV_STAGES = 2;                 // K and V tiles in flight per CTA (x 3 CTAs per SM)
constexpr int TILE_BYTES = 128 * 128;        // Q / P tile: 128 rows x 128 B
constexpr int KV_BYTES = KT * 128;           // K / V tile: 64 rows x 128 B
constexpr int SMEM_BYTES = 2 * TILE_BYTES /*Q, P*/ + 2 * KV_STAGES * KV_BYTES + 1024 /*align*/ +
                           256 /*barriers*/;   // 65 KB: three CTAs per SM
constexpr int SOFT_WARPS = 4;                // one thread per query row
constexpr int ATT_THREADS = 64 + SOFT_WARPS * 32;
constexpr int TMEM_COLS = 128;  // S: [0,64)  O: [64,128)  -- 3 CTAs x 128 of the SM's 512 columns

struct AttnDev {
  __nv_bfloat16* O;
  int f16;   // Q / K / V / P / O are fp16 instead of bf16 (half16.cuh)
  long long ldo;
  const int *q_start, *q_len, *kv_start0, *kv_len0, *kv_start1, *kv_len1;
  int causal_block;
  float scale_log2;
  int tail;  // the last `tail` query rows of every item are left to attention_tail_kernel
  float* lse;  // optional (rows, heads): log2-domain log-sum-exp of the scaled scores, for the backward pass
  int heads;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// F16 (the 16-bit format of Q / K / V / P / O) is a compile-time choice: the P conversions sit in the softmax loop
template <bool F16>
__global__ void __launch_bounds__(ATT_THREADS, 3)
    attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnDev a) {
  const int item = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
  const int q_len = max(a.q_len[item] - a.tail, 0);
  pdl_launch_dependents();       // the projection GEMM behind this kernel may start its prologue
  if (qt * QT >= q_len) return;  // uniform for the CTA, before any barrier / allocation
  const int q_row0 = a.q_start[item] + qt * QT;
  const int s0 = a.kv_start0[item], l0 = a.kv_len0[item];
  const int s1 = a.kv_start1 ? a.kv_start1[item] : 0, l1 = a.kv_len1 ? a.kv_len1[item] : 0;
  const int n0 = (l0 + KT - 1) / KT, n1 = (l1 + KT - 1) / KT;
  const int nt = n0 + n1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sP = smem + TILE_BYTES;                       // one P tile (128 rows x 64 keys, bf16)
  uint8_t* sK = smem + 2 * TILE_BYTES;                   // ring of KV_STAGES tiles
  uint8_t* sV = sK + KV_STAGES * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KV_STAGES * KV_BYTES);
  uint64_t *q_full = bars, *s_full = bars + 1, *s_free = bars + 2, *p_full = bars + 3,
           *o_full = bars + 4;
  uint64_t *k_full = bars + 8, *k_empty = k_full + KV_STAGES, *v_full = k_empty + KV_STAGES,
           *v_empty = v_full + KV_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_empty + KV_STAGES);


  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(k_full + i, 1);
      mbar_init(k_empty + i, 1);
      mbar_init(v_full + i, 1);
      mbar_init(v_empty + i, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, SOFT_WARPS * 32);
    mbar_init(p_full, SOFT_WARPS * 32);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S at 0, O at 64

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, head * HD, q_row0);
      for (int j = 0, st = 0, ph = 0; j < nt; ++j) {
        const int row = j < n0 ? s0 + j * KT : s1 + (j - n0) * KT;
        mbar_wait(k_empty + st, ph ^ 1);
        mbar_expect_tx(k_full + st, KV_BYTES);
        tma_load_2d(sK + st * KV_BYTES, &tmK, k_full + st, head * HD, row);
        mbar_wait(v_empty + st, ph ^ 1);
        mbar_expect_tx(v_full + st, KV_BYTES);
        tma_load_2d(sV + st * KV_BYTES, &tmV, v_full + st, head * HD, row);
        if (++st == KV_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // Software pipeline: S(j+1) = Q K(j+1)^T is issued as soon as the softmax warps have copied
    // S(j) into registers (s_free), i.e. BEFORE the probabilities of tile j exist, so the tensor
    // core works on the next scores while the exponentials are computed.  S, P and O are single
    // buffered (128 TMEM columns and 65 KB of shared memory per CTA -> three CTAs per SM, whose
    // phases interleave); every barrier completes once per key tile, phase parity = j & 1.
    if (lane == 0) {
      constexpr uint32_t fmt_clear = F16 ? ~((1u << 7) | (1u << 10)) : ~0u;   // format bits: 1 = bf16, 0 = fp16
      const uint32_t idesc_s = umma_idesc_bf16(QT, KT, 0) & fmt_clear;
      const uint32_t idesc_o = umma_idesc_bf16(QT, HD, 1) & fmt_clear;
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV),
                     p_addr = smem_u32(sP);
      auto issue_s = [&](int j, int st) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base, umma_desc_k_sw128(q_addr + k * 32),
                       umma_desc_k_sw128(k_addr + st * KV_BYTES + k * 32), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(k_empty + st);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      int st_k = 0, ph_k = 0;   // ring position of the next K tile to consume
      if (nt > 0) {
        mbar_wait(k_full + st_k, ph_k);
        tc_fence_after();
        issue_s(0, st_k);
        if (++st_k == KV_STAGES) { st_k = 0; ph_k ^= 1; }
      }
      for (int j = 0, st_v = 0, ph_v = 0; j < nt; ++j) {
        if (j + 1 < nt) {
          mbar_wait(k_full + st_k, ph_k);
          mbar_wait(s_free, j & 1);          // S(j) is in the softmax warps' registers
          tc_fence_after();
          issue_s(j + 1, st_k);
          if (++st_k == KV_STAGES) { st_k = 0; ph_k ^= 1; }
        }
        mbar_wait(v_full + st_v, ph_v);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const int seg_left = j < n0 ? l0 - j * KT : l1 - (j - n0) * KT;
        const int ksteps = (min(seg_left, KT) + 15) >> 4;   // keys beyond the segment: P is not even written
        // O accumulates in TMEM across the key tiles (the softmax warps rescale it in place on the
        // rare steps where the running row maximum moved by more than the stale-max threshold)
#pragma unroll 1
        for (int k = 0; k < ksteps; ++k)
          umma_bf16_ss(tmem_base + KT,
                       umma_desc_k_sw128(p_addr + k * 32),
                       umma_desc_mn_sw128(v_addr + st_v * KV_BYTES + k * 2048), idesc_o,
                       (j | k) != 0 ? 1u : 0u);
        umma_commit(v_empty + st_v);
        umma_commit(o_full);
        if (++st_v == KV_STAGES) { st_v = 0; ph_v ^= 1; }
      }
    }
  } else {
    // 4 softmax warps: thread r owns query row r (TMEM lane r) and all 64 key columns of every S
    // tile, so the row maximum / sum never leave the thread (no shared-memory exchange, no named
    // barrier between warp pairs) and the per-step bookkeeping is paid once per row instead of
    // twice.  The accumulation of O'(j) is deferred until after the probabilities of tile j+1 have
    // been handed to the tensor core, so it never sits between two MMAs.
    const int q = warp & 3;  // TMEM lane quarter of this warp
    const int r = q * 32 + lane;
    const int grow = q_row0 + r;  // absolute query row
    const bool row_valid = (qt * QT + r) < q_len;
    int lim = 0x7fffffff;
    if (a.causal_block > 0 && (grow % a.causal_block) == 0)
      lim = (grow / a.causal_block + 1) * a.causal_block;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // m: the (stale) row maximum the exponentials are taken against, in the log2 domain.  It only
    // follows the true maximum when that has grown by more than STALE (probabilities stay <= 2^8,
    // harmless for the bf16 P tile and the fp32 sums), so O -- which lives in TMEM and is
    // accumulated by the tensor core itself -- needs rescaling on very few steps.
    constexpr float STALE = 8.0f;
    float m = -INFINITY, l = 0.f;
    const int sw = r & 7;

    for (int j = 0; j < nt; ++j) {
      const int row0 = j < n0 ? s0 + j * KT : s1 + (j - n0) * KT;
      const int seg_left = j < n0 ? l0 - j * KT : l1 - (j - n0) * KT;
      const int nvalid = min(min(seg_left, KT), lim - row0);  // keys [0, nvalid) of this tile count
      const int nseg = min(seg_left, KT);
      const bool two = nseg > 32;               // the second 32-column half exists (warp-uniform)
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row maximum, 32 columns at a time (only 32 score registers are ever live: three
      // CTAs per SM leave 96 registers per thread)
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        uint32_t v[32];
        tmem_ld_32x32(t_lane + h * 32, v);
        tmem_ld_wait();
        const int my_valid = nvalid - h * 32;
        if (my_valid >= 32) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx0 = fmaxf(mx0, __uint_as_float(v[i]));
            mx1 = fmaxf(mx1, __uint_as_float(v[i + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(v[i + 2]));
            mx3 = fmaxf(mx3, __uint_as_float(v[i + 3]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < my_valid) mx0 = fmaxf(mx0, __uint_as_float(v[i]));
        }
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * a.scale_log2;
      const bool grow_m = mx > m + STALE || (m == -INFINITY && mx > -INFINITY);
      if (__any_sync(0xffffffffu, grow_m)) {
        // rescale this warp's 32 rows of O (and l) by 2^(m - m_new); rows that keep m use 1
        const float alpha = grow_m ? ex2_approx(m - mx) : 1.0f;   // m == -inf -> 0
        if (grow_m) m = mx;
        l *= alpha;
        if (j > 0) {
          mbar_wait(o_full, (j - 1) & 1);   // PV(j-1) has landed
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32];
            tmem_ld_32x32(t_lane + KT + h * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32(t_lane + KT + h * 32, v);
          }
          tmem_st_wait();
        }
      }
      const float m_use = (m == -INFINITY) ? 0.f : m;
      float rs0 = 0.f, rs1 = 0.f;
      uint8_t* p_row = sP + r * 128;
      bool p_free = j == 0;              // PV(j-1) must have consumed the P tile before it is rewritten
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        uint32_t v[32];                    // pass 2: the same 32 scores again
        tmem_ld_32x32(t_lane + h * 32, v);
        tmem_ld_wait();
        if (h == 1 || !two) {              // S(j) is not needed any more: the tensor core may
          tc_fence_before();               // overwrite it with tile j+1
          mbar_arrive(s_free);
        }
        const int my_valid = nvalid - h * 32;
        uint32_t pk[16];
        if (my_valid >= 32) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2_approx(__uint_as_float(v[i]) * a.scale_log2 - m_use);
            const float p1 = ex2_approx(__uint_as_float(v[i + 1]) * a.scale_log2 - m_use);
            rs0 += p0;
            rs1 += p1;
            pk[i >> 1] = f2_to_h2(p0, p1, F16);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = ex2_approx(__uint_as_float(v[i]) * a.scale_log2 - m_use);
            float p1 = ex2_approx(__uint_as_float(v[i + 1]) * a.scale_log2 - m_use);
            if (i >= my_valid) p0 = 0.f;
            if (i + 1 >= my_valid) p1 = 0.f;
            rs0 += p0;
            rs1 += p1;
            pk[i >> 1] = f2_to_h2(p0, p1, F16);
          }
        }
        if (!p_free) {
          mbar_wait(o_full, (j - 1) & 1);
          p_free = true;
        }
        // 32 columns = 16-byte chunks h*4 .. +3 of the row in the P tile
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = (h * 4 + ch) ^ sw;
          *reinterpret_cast<uint4*>(p_row + chunk * 16) =
              make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(p_full);
      l += rs0 + rs1;
    }
    if (nt > 0) {
      mbar_wait(o_full, (nt - 1) & 1);
      tc_fence_after();
    }
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    if (a.lse != nullptr && row_valid)   // p = exp2(s * scale_log2 - lse) reproduces the probabilities
      a.lse[static_cast<long long>(grow) * a.heads + head] = l > 0.f ? m + log2f(l) : INFINITY;
    __nv_bfloat16* dst = a.O + static_cast<long long>(grow) * a.ldo + head * HD;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t v[32];
      if (nt > 0) {
        tmem_ld_32x32(t_lane + KT + h * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (row_valid) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            w[i] = f2_to_h2(__uint_as_float(v[ch * 8 + 2 * i]) * inv, __uint_as_float(v[ch * 8 + 2 * i + 1]) * inv,
                            F16);
          }
          *reinterpret_cast<uint4*>(dst + h * 32 + ch * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ tail rows
// The per-frame token count of this model is 2 x 128 + 1 (256 patches + the intrinsic token): a
// third 128-row query tile for ONE row would add 50 % more CTAs.  Such 1..4-row remainders are done
// on the CUDA cores instead: one 256-thread block per (item, head, row).
//   phase 1  thread t scores keys t, t+256, ... (one 128-byte K row per load), scores -> smem
//   phase 2  block max / sum of exp2
//   phase 3  thread (part, d) accumulates sum_j p_j V[j][d] over every 4th key, partials -> smem
constexpr int TAIL_MAX_ROWS = 4;
constexpr int TAIL_MAX_KEYS = 2048;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const float o = __shfl_xor_sync(0xffffffffu, v, s);
    v = is_max ? fmaxf(v, o) : v + o;
  }
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

__global__ void __launch_bounds__(256)
    attention_tail_kernel(const __nv_bfloat16* __restrict__ Q, const __nv_bfloat16* __restrict__ K,
                          const __nv_bfloat16* __restrict__ V, long long ldq, long long ldk,
                          long long ldv, int heads, const AttnDev a) {
  __shared__ float s_q[64];
  __shared__ float s_p[TAIL_MAX_KEYS];
  __shared__ float s_red[8];
  __shared__ float s_o[32][64];
  const int t = blockIdx.x % a.tail;
  const int head = (blockIdx.x / a.tail) % heads;
  const int item = blockIdx.x / (a.tail * heads);
  const int rin = a.q_len[item] - a.tail + t;
  if (rin < 0) return;
  const int grow = a.q_start[item] + rin;
  int lim = 0x7fffffff;
  if (a.causal_block > 0 && (grow % a.causal_block) == 0)
    lim = (grow / a.causal_block + 1) * a.causal_block;
  const int s0 = a.kv_start0[item], l0 = a.kv_len0[item];
  const int s1 = a.kv_start1 ? a.kv_start1[item] : 0, l1 = a.kv_len1 ? a.kv_len1[item] : 0;
  const int nk = min(l0 + l1, TAIL_MAX_KEYS);
  if (threadIdx.x < 64)
    s_q[threadIdx.x] = h_to_f(reinterpret_cast<const uint16_t*>(Q)[static_cast<long long>(grow) * ldq + head * HD +
                                                                      threadIdx.x], a.f16) * a.scale_log2;
  __syncthreads();
  // phase 1: 8 lanes per key (one 16-byte load each: a warp instruction touches 4 key rows of 128 B; a thread
  // per key read 32 different lines per instruction and left ONE thread working on the 257-th key)
  float mx = -INFINITY;
  {
    const int d8q = (threadIdx.x & 7) * 8, kpart = threadIdx.x >> 3;
    float q8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q8[i] = s_q[d8q + i];
    for (int j0 = 0; j0 < nk; j0 += 32) {
      const int j = j0 + kpart;
      const int krow = j < l0 ? s0 + j : s1 + (j - l0);
      const bool on = j < nk && krow < lim;
      float s = 0.f;
      if (on) {
        const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(K + static_cast<long long>(krow) * ldk + head * HD + d8q));
        const uint32_t w[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 f = h2_to_f2(w[u], a.f16);
          s = fmaf(q8[2 * u], f.x, s);
          s = fmaf(q8[2 * u + 1], f.y, s);
        }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (!on) s = -INFINITY;
      if (j < nk && (threadIdx.x & 7) == 0) s_p[j] = s;
      mx = fmaxf(mx, s);
    }
  }
  mx = block_reduce(mx, s_red, true);
  const float m_use = mx == -INFINITY ? 0.f : mx;
  float sum = 0.f;
  for (int j = threadIdx.x; j < nk; j += 256) {
    const float p = ex2_approx(s_p[j] - m_use);
    s_p[j] = p;
    sum += p;
  }
  sum = block_reduce(sum, s_red, false);   // also orders the s_p writes before phase 3
  // phase 3: thread = (part of 32, 8 output dims): one 16-byte V load per key, keys part, part+32, ..
  const int d8 = (threadIdx.x & 7) * 8, part = threadIdx.x >> 3;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
  for (int j = part; j < nk; j += 32) {
    const int krow = j < l0 ? s0 + j : s1 + (j - l0);
    const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(V + static_cast<long long>(krow) * ldv + head * HD + d8));
    const uint32_t w[4] = {t4.x, t4.y, t4.z, t4.w};
    const float p = s_p[j];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 f = h2_to_f2(w[u], a.f16);
      acc[2 * u] = fmaf(p, f.x, acc[2 * u]);
      acc[2 * u + 1] = fmaf(p, f.y, acc[2 * u + 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_o[part][d8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x == 0 && a.lse != nullptr)
    a.lse[static_cast<long long>(grow) * a.heads + head] = sum > 0.f ? m_use + log2f(sum) : INFINITY;
  if (threadIdx.x < 64) {
    float o = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) o += s_o[i][threadIdx.x];
    reinterpret_cast<uint16_t*>(a.O)[static_cast<long long>(grow) * a.ldo + head * HD + threadIdx.x] =
        f_to_h(sum > 0.f ? o / sum : 0.f, a.f16);
  }
}

int make_map(CUtensorMap* map, const void* base, int heads, int rows, long long ld, int box_rows) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(heads) * HD, static_cast<cuuint64_t>(rows)};
  cuuint64_t str[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {HD, static_cast<cuuint32_t>(box_rows)};
  return encode_map(map, base, 2, dims, str, box);
}

}  // namespace
}  // namespace vs

extern "C" int vs_attention(const vs_attention_params* p, vs_stream_t stream_) {
  using namespace vs;
  VS_REQUIRE(p != nullptr, "vs_attention: null params");
  VS_REQUIRE(p->Q && p->K && p->V && p->O, "vs_attention: null tensor");
  VS_REQUIRE(p->heads > 0 && p->items >= 0 && p->max_q_len >= 0, "vs_attention: bad sizes");
  VS_REQUIRE(p->q_start && p->q_len && p->kv_start0 && p->kv_len0,
             "vs_attention: item arrays missing");
  VS_REQUIRE((p->kv_start1 == nullptr) == (p->kv_len1 == nullptr),
             "vs_attention: kv_start1 / kv_len1 go together");
  VS_REQUIRE(p->ldq % 8 == 0 && p->ldk % 8 == 0 && p->ldv % 8 == 0 && p->ldo % 8 == 0,
             "vs_attention: leading dimensions must be multiples of 8 elements");
  VS_REQUIRE(((reinterpret_cast<uintptr_t>(p->Q) | reinterpret_cast<uintptr_t>(p->K) |
               reinterpret_cast<uintptr_t>(p->V) | reinterpret_cast<uintptr_t>(p->O)) & 15) == 0,
             "vs_attention: tensors must be 16-byte aligned");
  VS_REQUIRE(p->q_rows > 0 && p->kv_rows > 0, "vs_attention: q_rows / kv_rows must be positive");
  if (p->items == 0 || p->max_q_len == 0) return VS_OK;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_map(&tmQ, p->Q, p->heads, p->q_rows, p->ldq, QT);
  if (rc) return rc;
  rc = make_map(&tmK, p->K, p->heads, p->kv_rows, p->ldk, KT);
  if (rc) return rc;
  rc = make_map(&tmV, p->V, p->heads, p->kv_rows, p->ldv, KT);
  if (rc) return rc;
  AttnDev a{};
  a.O = static_cast<__nv_bfloat16*>(p->O);
  VS_REQUIRE(p->dtype == 0 || p->dtype == VS_BF16 || p->dtype == VS_F16, "vs_attention: dtype must be bf16 or fp16");
  a.f16 = p->dtype == VS_F16;
  a.ldo = p->ldo;
  a.q_start = p->q_start;
  a.q_len = p->q_len;
  a.kv_start0 = p->kv_start0;
  a.kv_len0 = p->kv_len0;
  a.kv_start1 = p->kv_start1;
  a.kv_len1 = p->kv_len1;
  a.causal_block = p->causal_block;
  a.scale_log2 = p->scale * 1.4426950408889634f;
  a.lse = p->lse;
  a.heads = p->heads;
  const int rem = p->max_q_len % QT;
  a.tail = (rem > 0 && rem <= TAIL_MAX_ROWS && p->max_kv_len > 0 && p->max_kv_len <= TAIL_MAX_KEYS)
               ? rem : 0;
  VS_CONFIGURE_PER_DEVICE(
    VS_CUDA(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    VS_CUDA(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  );
  const int main_rows = p->max_q_len - a.tail;
  if (main_rows > 0) {
    dim3 grid(ceil_div(main_rows, QT), p->heads, p->items);
    (a.f16 ? attention_kernel<true> : attention_kernel<false>)<<<grid, ATT_THREADS, SMEM_BYTES, to_stream(stream_)>>>(
        tmQ, tmK, tmV, a);
    VS_LAUNCH_CHECK();
  }
  if (a.tail > 0) {
    attention_tail_kernel<<<static_cast<unsigned>(p->items * p->heads * a.tail), 256, 0,
                            to_stream(stream_)>>>(
        static_cast<const __nv_bfloat16*>(p->Q), static_cast<const __nv_bfloat16*>(p->K),
        static_cast<const __nv_bfloat16*>(p->V), p->ldq, p->ldk, p->ldv, p->heads, a);
    VS_LAUNCH_CHECK();
  }
  return VS_OK;
}
