// Backward pass of vs_attention (head_dim 64) on tcgen05 (sm_100a): the gradients the reference gets
// from torch.autograd through softmax(q k^T / 8) v (croco/blocks.py:105-109).
//
// With P = exp2(S * scale_log2 - lse)  (lse kept by the forward kernel) and D = rowsum(dO . O):
//     dP = dO V^T        dS = P . (dP - D) * scale
//     dQ = dS K          dK = dS^T Q          dV = P^T dO
// Two kernels, each with the structure (and the verified operand layouts) of the forward kernel --
// scores recomputed in both, nothing but O(rows) statistics is kept from the forward pass, no atomics:
//
//   attention_bwd_dq_kernel    one CTA per (item, head, 128 queries), loops over 64-key tiles:
//       S = Q K^T, dP = dO V^T in TMEM (operands K-major as loaded by TMA)  ->  4 warps, one thread per
//       query row, write dS (bf16, K-major swizzled tile)  ->  dQ += dS K in TMEM (K consumed MN-major).
//   attention_bwd_dkv_kernel   one CTA per (item, head, 128 keys), loops over 64-query tiles:
//       S^T = K Q^T, dP^T = V dO^T in TMEM  ->  one thread per KEY row writes P^T and dS^T tiles  ->
//       dV += P^T dO, dK += dS^T Q in TMEM (dO / Q tiles consumed MN-major).
//   attention_bwd_delta_kernel D[row, head] = sum_d dO . O.
//
// dK / dV are written, not accumulated, so the key rows of different items of the dK/dV kernel must not
// overlap.  True for the ViT encoder's per-frame attention and the decoder's video attention with the
// forward tables; the neighbour cross-attention shares key frames between query frames: there the dK/dV
// kernel runs over KEY-centric items (one key frame + the up to two query frames that read it as two
// query segments), so every dK / dV row is still written exactly once -- no atomics, no extra pass.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"
#include "tmap.h"

namespace vs {
namespace {

constexpr int HD = 64;
constexpr int T128 = 128 * 128;   // bytes of a 128-row x 64-column bf16 tile
constexpr int T64 = 64 * 128;     // bytes of a 64-row tile
constexpr int STAGES = 2;
constexpr int BWD_THREADS = 64 + 128;
constexpr int BWD_TMEM_COLS = 256;
// dQ kernel: Q, dO, dS (128-row tiles) + ring of (K, V) 64-row tiles
constexpr int DQ_SMEM = 3 * T128 + STAGES * 2 * T64 + 1024 + 256;
// dK/dV kernel: K, V, P^T, dS^T (128-row tiles) + ring of (Q, dO) 64-row tiles + per-query statistics
constexpr int DKV_SMEM = 4 * T128 + STAGES * 2 * T64 + 1024 + 256 + 2 * 64 * 12;

struct BwdDev {
  const float* lse;    // (rows, heads) log2-domain log-sum-exp from the forward pass
  const float* delta;  // (rows, heads)
  __nv_bfloat16 *dQ, *dK, *dV;
  long long lddq, lddk, lddv;
  const int *q_start, *q_len, *kv_start0, *kv_len0, *kv_start1, *kv_len1;
  // item tables of the dK / dV kernel: up to two key segments AND up to two query segments per item
  // (query-centric items: the forward tables; key-centric items: one key frame + the query frames
  // that read it, so that every dK / dV row is written exactly once)
  const int *dk_start0, *dk_len0, *dk_start1, *dk_len1, *dq_start0, *dq_len0, *dq_start1, *dq_len1;
  int causal_block;
  int heads;
  float scale, scale_log2;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void bar_sync_softmax() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// 32 fp32 values of one TMEM lane -> 32 bf16 at dst (64 B, 16-byte aligned)
__device__ __forceinline__ void store_row32(__nv_bfloat16* dst, const uint32_t (&v)[32], float mul) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      w[i] = pack2(__uint_as_float(v[ch * 8 + 2 * i]) * mul, __uint_as_float(v[ch * 8 + 2 * i + 1]) * mul);
    *reinterpret_cast<uint4*>(dst + ch * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ------------------------------------------------------------------ dQ
__global__ void __launch_bounds__(BWD_THREADS, 2)
    attention_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                            const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                            const BwdDev a) {
  constexpr int QT = 128, KT = 64;
  const int item = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
  const int q_len = a.q_len[item];
  if (qt * QT >= q_len) return;
  const int q_row0 = a.q_start[item] + qt * QT;
  const int s0 = a.kv_start0[item], l0 = a.kv_len0[item];
  const int s1 = a.kv_start1 ? a.kv_start1[item] : 0, l1 = a.kv_len1 ? a.kv_len1[item] : 0;
  const int n0 = (l0 + KT - 1) / KT, n1 = (l1 + KT - 1) / KT;
  const int nt = n0 + n1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + T128;
  uint8_t* sdS = smem + 2 * T128;
  uint8_t* sK = smem + 3 * T128;            // ring: STAGES x K tile
  uint8_t* sV = sK + STAGES * T64;          // ring: STAGES x V tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * T64);
  uint64_t *q_full = bars, *s_full = bars + 1, *s_free = bars + 2, *ds_full = bars + 3,
           *ds_free = bars + 4;
  uint64_t *kv_full = bars + 8, *kv_empty = kv_full + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(kv_full + i, 1);
      mbar_init(kv_empty + i, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(ds_full, 128);
    mbar_init(ds_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BWD_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;   // columns: S [0,64)  dP [64,128)  dQ [128,192)

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * T128);
      tma_load_2d(sQ, &tmQ, q_full, head * HD, q_row0);
      tma_load_2d(sdO, &tmdO, q_full, head * HD, q_row0);
      for (int j = 0, st = 0, ph = 0; j < nt; ++j) {
        const int row = j < n0 ? s0 + j * KT : s1 + (j - n0) * KT;
        mbar_wait(kv_empty + st, ph ^ 1);
        mbar_expect_tx(kv_full + st, 2 * T64);
        tma_load_2d(sK + st * T64, &tmK, kv_full + st, head * HD, row);
        tma_load_2d(sV + st * T64, &tmV, kv_full + st, head * HD, row);
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(QT, KT, 0);
      constexpr uint32_t idesc_q = umma_idesc_bf16(QT, HD, 1);
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sdO), ds_addr = smem_u32(sdS),
                     k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      mbar_wait(q_full, 0);
      for (int j = 0, st = 0, ph = 0; j < nt; ++j) {
        mbar_wait(kv_full + st, ph);
        if (j > 0) mbar_wait(s_free, (j - 1) & 1);   // S / dP of tile j-1 are in registers
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base, umma_desc_k_sw128(q_addr + k * 32),
                       umma_desc_k_sw128(k_addr + st * T64 + k * 32), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base + 64, umma_desc_k_sw128(do_addr + k * 32),
                       umma_desc_k_sw128(v_addr + st * T64 + k * 32), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(ds_full, j & 1);                   // dS(j) is in shared memory
        tc_fence_after();
        const int seg_left = j < n0 ? l0 - j * KT : l1 - (j - n0) * KT;
        const int ksteps = (min(seg_left, KT) + 15) >> 4;
#pragma unroll 1
        for (int k = 0; k < ksteps; ++k)
          umma_bf16_ss(tmem_base + 128, umma_desc_k_sw128(ds_addr + k * 32),
                       umma_desc_mn_sw128(k_addr + st * T64 + k * 2048), idesc_q,
                       (j | k) != 0 ? 1u : 0u);
        umma_commit(kv_empty + st);
        umma_commit(ds_free);
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const int grow = q_row0 + r;
    const bool row_valid = (qt * QT + r) < q_len;
    int lim = 0x7fffffff;
    if (a.causal_block > 0 && (grow % a.causal_block) == 0)
      lim = (grow / a.causal_block + 1) * a.causal_block;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float L = row_valid ? a.lse[static_cast<long long>(grow) * a.heads + head] : INFINITY;
    const float Dr = row_valid ? a.delta[static_cast<long long>(grow) * a.heads + head] : 0.f;
    const int sw = r & 7;
    uint8_t* ds_row = sdS + r * 128;
    // a warp whose 32 query rows all lie beyond the item (the 257-th token leaves 127 such rows in the
    // third tile) only keeps the barriers going: its dS rows feed nothing but its own unused dQ rows
    const bool warp_active = __any_sync(0xffffffffu, row_valid);

    for (int j = 0; j < nt; ++j) {
      if (!warp_active) {
        mbar_wait(s_full, j & 1);
        mbar_arrive(s_free);
        mbar_arrive(ds_full);
        continue;
      }
      const int row0 = j < n0 ? s0 + j * KT : s1 + (j - n0) * KT;
      const int seg_left = j < n0 ? l0 - j * KT : l1 - (j - n0) * KT;
      const int nvalid = row_valid ? min(min(seg_left, KT), lim - row0) : 0;
      const int nseg = min(seg_left, KT);
      const bool two = nseg > 32;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      bool ds_writable = j == 0;   // dQ MMA of tile j-1 must have consumed the dS tile
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        uint32_t v[32], w[32];
        tmem_ld_32x32(t_lane + h * 32, v);
        tmem_ld_32x32(t_lane + 64 + h * 32, w);
        tmem_ld_wait();
        if (h == 1 || !two) {
          tc_fence_before();
          mbar_arrive(s_free);
        }
        const int my_valid = nvalid - h * 32;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ex2_approx(__uint_as_float(v[i]) * a.scale_log2 - L);
          const float p1 = ex2_approx(__uint_as_float(v[i + 1]) * a.scale_log2 - L);
          float d0 = p0 * (__uint_as_float(w[i]) - Dr) * a.scale;
          float d1 = p1 * (__uint_as_float(w[i + 1]) - Dr) * a.scale;
          if (i >= my_valid) d0 = 0.f;
          if (i + 1 >= my_valid) d1 = 0.f;
          pk[i >> 1] = pack2(d0, d1);
        }
        if (!ds_writable) {
          mbar_wait(ds_free, (j - 1) & 1);
          ds_writable = true;
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = (h * 4 + ch) ^ sw;
          *reinterpret_cast<uint4*>(ds_row + chunk * 16) =
              make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(ds_full);
    }
    if (nt > 0) {
      mbar_wait(ds_free, (nt - 1) & 1);   // the last dQ MMA has landed
      tc_fence_after();
    }
    __nv_bfloat16* dst = a.dQ + static_cast<long long>(grow) * a.lddq + head * HD;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t v[32];
      if (nt > 0) {
        tmem_ld_32x32(t_lane + 128 + h * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (row_valid) store_row32(dst + h * 32, v, 1.0f);
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BWD_TMEM_COLS);
}

// ------------------------------------------------------------------ dK, dV
__global__ void __launch_bounds__(BWD_THREADS, 2)
    attention_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                             const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                             const BwdDev a) {
  constexpr int KT = 128, QT = 64;
  const int item = blockIdx.z, head = blockIdx.y, kt = blockIdx.x;
  const int l0 = a.dk_len0[item], l1 = a.dk_len1 ? a.dk_len1[item] : 0;
  const int n0 = (l0 + KT - 1) / KT, n1 = (l1 + KT - 1) / KT;
  if (kt >= n0 + n1) return;
  const bool seg1 = kt >= n0;
  const int k_row0 = seg1 ? a.dk_start1[item] + (kt - n0) * KT : a.dk_start0[item] + kt * KT;
  const int k_valid = min(KT, seg1 ? l1 - (kt - n0) * KT : l0 - kt * KT);
  const int qs0 = a.dq_start0[item], ql0 = a.dq_len0[item];
  const int qs1 = a.dq_start1 ? a.dq_start1[item] : 0, ql1 = a.dq_len1 ? a.dq_len1[item] : 0;
  const int nq0 = (ql0 + QT - 1) / QT;
  const int nq = nq0 + (ql1 + QT - 1) / QT;
  // 64-query tile j: first row, and how many of its rows belong to the segment
  auto tile_row0 = [&](int j) { return j < nq0 ? qs0 + j * QT : qs1 + (j - nq0) * QT; };
  auto tile_rows = [&](int j) { return j < nq0 ? ql0 - j * QT : ql1 - (j - nq0) * QT; };

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + T128;
  uint8_t* sP = smem + 2 * T128;            // P^T tile: 128 key rows x 64 queries
  uint8_t* sdS = smem + 3 * T128;           // dS^T tile
  uint8_t* sQ = smem + 4 * T128;            // ring: STAGES x Q tile
  uint8_t* sdO = sQ + STAGES * T64;         // ring: STAGES x dO tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdO + STAGES * T64);
  uint64_t *k_full = bars, *s_full = bars + 1, *s_free = bars + 2, *p_full = bars + 3,
           *p_free = bars + 4;
  uint64_t *qd_full = bars + 8, *qd_empty = qd_full + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qd_empty + STAGES);
  float* sL = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2][64]
  float* sD = sL + 2 * 64;                                                         // [2][64]
  int* sLim = reinterpret_cast<int*>(sD + 2 * 64);                                 // [2][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(k_full, 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(qd_full + i, 1);
      mbar_init(qd_empty + i, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(p_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BWD_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;   // columns: S^T [0,64)  dP^T [64,128)  dV [128,192)  dK [192,256)

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(k_full, 2 * T128);
      tma_load_2d(sK, &tmK, k_full, head * HD, k_row0);
      tma_load_2d(sV, &tmV, k_full, head * HD, k_row0);
      for (int j = 0, st = 0, ph = 0; j < nq; ++j) {
        mbar_wait(qd_empty + st, ph ^ 1);
        mbar_expect_tx(qd_full + st, 2 * T64);
        tma_load_2d(sQ + st * T64, &tmQ, qd_full + st, head * HD, tile_row0(j));
        tma_load_2d(sdO + st * T64, &tmdO, qd_full + st, head * HD, tile_row0(j));
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(KT, QT, 0);
      constexpr uint32_t idesc_g = umma_idesc_bf16(KT, HD, 1);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP),
                     ds_addr = smem_u32(sdS), q_addr = smem_u32(sQ), do_addr = smem_u32(sdO);
      mbar_wait(k_full, 0);
      for (int j = 0, st = 0, ph = 0; j < nq; ++j) {
        mbar_wait(qd_full + st, ph);
        if (j > 0) mbar_wait(s_free, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base, umma_desc_k_sw128(k_addr + k * 32),
                       umma_desc_k_sw128(q_addr + st * T64 + k * 32), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_base + 64, umma_desc_k_sw128(v_addr + k * 32),
                       umma_desc_k_sw128(do_addr + st * T64 + k * 32), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(p_full, j & 1);                    // P^T(j) and dS^T(j) are in shared memory
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < QT / 16; ++k)
          umma_bf16_ss(tmem_base + 128, umma_desc_k_sw128(p_addr + k * 32),
                       umma_desc_mn_sw128(do_addr + st * T64 + k * 2048), idesc_g,
                       (j | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < QT / 16; ++k)
          umma_bf16_ss(tmem_base + 192, umma_desc_k_sw128(ds_addr + k * 32),
                       umma_desc_mn_sw128(q_addr + st * T64 + k * 2048), idesc_g,
                       (j | k) != 0 ? 1u : 0u);
        umma_commit(qd_empty + st);
        umma_commit(p_free);
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;       // key row of the tile = TMEM lane
    const int t = threadIdx.x - 64;    // 0..127 among the softmax threads
    const int krow = k_row0 + r;       // absolute key row
    const bool key_valid = r < k_valid;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int sw = r & 7;
    uint8_t* p_row = sP + r * 128;
    uint8_t* ds_row = sdS + r * 128;
    const bool warp_active = __any_sync(0xffffffffu, key_valid);   // see the dQ kernel

    for (int j = 0; j < nq; ++j) {
      // per-query statistics of this 64-query tile -> shared memory (double buffered by tile parity)
      {
        const int b = (j & 1) * 64;
        const int qi = t & 63;
        const int qrow = tile_row0(j) + qi;
        const bool qv = qi < tile_rows(j);
        if (t < 64) {
          sL[b + qi] = qv ? a.lse[static_cast<long long>(qrow) * a.heads + head] : INFINITY;
          int lim = qv ? 0x7fffffff : 0;   // queries beyond the item see no key
          if (qv && a.causal_block > 0 && (qrow % a.causal_block) == 0)
            lim = (qrow / a.causal_block + 1) * a.causal_block;
          sLim[b + qi] = lim;
        } else {
          sD[b + qi] = qv ? a.delta[static_cast<long long>(qrow) * a.heads + head] : 0.f;
        }
      }
      bar_sync_softmax();
      if (!warp_active) {
        mbar_wait(s_full, j & 1);
        mbar_arrive(s_free);
        mbar_arrive(p_full);
        continue;
      }
      const float* Lb = sL + (j & 1) * 64;
      const float* Db = sD + (j & 1) * 64;
      const int* limb = sLim + (j & 1) * 64;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      bool writable = j == 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32], w[32];
        tmem_ld_32x32(t_lane + h * 32, v);
        tmem_ld_32x32(t_lane + 64 + h * 32, w);
        tmem_ld_wait();
        if (h == 1) {
          tc_fence_before();
          mbar_arrive(s_free);
        }
        uint32_t pp[16], pd[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int c0 = h * 32 + i;
          const bool ok0 = key_valid && krow < limb[c0];
          const bool ok1 = key_valid && krow < limb[c0 + 1];
          float p0 = ex2_approx(__uint_as_float(v[i]) * a.scale_log2 - Lb[c0]);
          float p1 = ex2_approx(__uint_as_float(v[i + 1]) * a.scale_log2 - Lb[c0 + 1]);
          float d0 = p0 * (__uint_as_float(w[i]) - Db[c0]) * a.scale;
          float d1 = p1 * (__uint_as_float(w[i + 1]) - Db[c0 + 1]) * a.scale;
          if (!ok0) { p0 = 0.f; d0 = 0.f; }
          if (!ok1) { p1 = 0.f; d1 = 0.f; }
          pp[i >> 1] = pack2(p0, p1);
          pd[i >> 1] = pack2(d0, d1);
        }
        if (!writable) {
          mbar_wait(p_free, (j - 1) & 1);   // the dV / dK MMAs of tile j-1 have consumed both tiles
          writable = true;
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = (h * 4 + ch) ^ sw;
          *reinterpret_cast<uint4*>(p_row + chunk * 16) =
              make_uint4(pp[ch * 4], pp[ch * 4 + 1], pp[ch * 4 + 2], pp[ch * 4 + 3]);
          *reinterpret_cast<uint4*>(ds_row + chunk * 16) =
              make_uint4(pd[ch * 4], pd[ch * 4 + 1], pd[ch * 4 + 2], pd[ch * 4 + 3]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    if (nq > 0) {
      mbar_wait(p_free, (nq - 1) & 1);
      tc_fence_after();
    }
    __nv_bfloat16* dv = a.dV + static_cast<long long>(krow) * a.lddv + head * HD;
    __nv_bfloat16* dk = a.dK + static_cast<long long>(krow) * a.lddk + head * HD;
#pragma unroll
    for (int h = 0; h < 4; ++h) {   // 0,1: dV halves   2,3: dK halves
      uint32_t v[32];
      if (nq > 0) {
        tmem_ld_32x32(t_lane + 128 + h * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (key_valid) store_row32((h < 2 ? dv : dk) + (h & 1) * 32, v, 1.0f);
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BWD_TMEM_COLS);
}

// ------------------------------------------------------------------ D = rowsum(dO . O)
__global__ void attention_bwd_delta_kernel(const __nv_bfloat16* __restrict__ O, long long ldo,
                                           const __nv_bfloat16* __restrict__ dO, long long lddo,
                                           float* __restrict__ delta, int rows, int heads) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * heads) return;
  const int row = static_cast<int>(i / heads), head = static_cast<int>(i - static_cast<long long>(row) * heads);
  const uint4* o4 = reinterpret_cast<const uint4*>(O + static_cast<long long>(row) * ldo + head * HD);
  const uint4* d4 = reinterpret_cast<const uint4*>(dO + static_cast<long long>(row) * lddo + head * HD);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint4 x = __ldg(o4 + k), y = __ldg(d4 + k);
    const uint32_t xa[4] = {x.x, x.y, x.z, x.w}, ya[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xa[u]));
      const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ya[u]));
      s = fmaf(f.x, g.x, s);
      s = fmaf(f.y, g.y, s);
    }
  }
  delta[i] = s;
}

int make_map(CUtensorMap* map, const void* base, int heads, int rows, long long ld, int box_rows) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(heads) * HD, static_cast<cuuint64_t>(rows)};
  cuuint64_t str[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {HD, static_cast<cuuint32_t>(box_rows)};
  return encode_map(map, base, 2, dims, str, box);
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace vs

extern "C" int vs_attention_backward(const vs_attention_bwd_params* p, vs_stream_t stream_) {
  using namespace vs;
  VS_REQUIRE(p != nullptr, "vs_attention_backward: null params");
  const vs_attention_params& f = p->fwd;
  VS_REQUIRE(f.Q && f.K && f.V && f.O && f.lse && p->dO && p->dQ && p->dK && p->dV && p->delta,
             "vs_attention_backward: null tensor");
  VS_REQUIRE(f.heads > 0 && f.items >= 0 && f.max_q_len >= 0 && f.max_kv_len > 0,
             "vs_attention_backward: bad sizes (max_kv_len is required)");
  VS_REQUIRE(f.q_start && f.q_len && f.kv_start0 && f.kv_len0, "vs_attention_backward: item arrays missing");
  VS_REQUIRE((f.kv_start1 == nullptr) == (f.kv_len1 == nullptr),
             "vs_attention_backward: kv_start1 / kv_len1 go together");
  VS_REQUIRE(f.ldq % 8 == 0 && f.ldk % 8 == 0 && f.ldv % 8 == 0 && f.ldo % 8 == 0 && p->lddo % 8 == 0 &&
                 p->lddq % 8 == 0 && p->lddk % 8 == 0 && p->lddv % 8 == 0,
             "vs_attention_backward: leading dimensions must be multiples of 8 elements");
  VS_REQUIRE(al16(f.Q) && al16(f.K) && al16(f.V) && al16(f.O) && al16(p->dO) && al16(p->dQ) &&
                 al16(p->dK) && al16(p->dV),
             "vs_attention_backward: tensors must be 16-byte aligned");
  VS_REQUIRE(f.q_rows > 0 && f.kv_rows > 0, "vs_attention_backward: q_rows / kv_rows must be positive");
  if (f.items == 0 || f.max_q_len == 0) return VS_OK;
  cudaStream_t s = to_stream(stream_);

  {
    const long long n = static_cast<long long>(f.q_rows) * f.heads;
    attention_bwd_delta_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(f.O), f.ldo, static_cast<const __nv_bfloat16*>(p->dO), p->lddo,
        p->delta, f.q_rows, f.heads);
    VS_LAUNCH_CHECK();
  }

  BwdDev a{};
  a.lse = f.lse;
  a.delta = p->delta;
  a.dQ = static_cast<__nv_bfloat16*>(p->dQ);
  a.dK = static_cast<__nv_bfloat16*>(p->dK);
  a.dV = static_cast<__nv_bfloat16*>(p->dV);
  a.lddq = p->lddq;
  a.lddk = p->lddk;
  a.lddv = p->lddv;
  a.q_start = f.q_start;
  a.q_len = f.q_len;
  a.kv_start0 = f.kv_start0;
  a.kv_len0 = f.kv_len0;
  a.kv_start1 = f.kv_start1;
  a.kv_len1 = f.kv_len1;
  const bool key_centric = p->dkv_items > 0;
  if (key_centric) {
    VS_REQUIRE(p->dkv_kv_start && p->dkv_kv_len && p->dkv_q_start0 && p->dkv_q_len0 && p->dkv_max_kv_len > 0,
               "vs_attention_backward: key-centric dK/dV tables incomplete");
    VS_REQUIRE((p->dkv_q_start1 == nullptr) == (p->dkv_q_len1 == nullptr),
               "vs_attention_backward: dkv_q_start1 / dkv_q_len1 go together");
    a.dk_start0 = p->dkv_kv_start;
    a.dk_len0 = p->dkv_kv_len;
    a.dk_start1 = nullptr;
    a.dk_len1 = nullptr;
    a.dq_start0 = p->dkv_q_start0;
    a.dq_len0 = p->dkv_q_len0;
    a.dq_start1 = p->dkv_q_start1;
    a.dq_len1 = p->dkv_q_len1;
  } else {
    a.dk_start0 = f.kv_start0;
    a.dk_len0 = f.kv_len0;
    a.dk_start1 = f.kv_start1;
    a.dk_len1 = f.kv_len1;
    a.dq_start0 = f.q_start;
    a.dq_len0 = f.q_len;
    a.dq_start1 = nullptr;
    a.dq_len1 = nullptr;
  }
  a.causal_block = f.causal_block;
  a.heads = f.heads;
  a.scale = f.scale;
  a.scale_log2 = f.scale * 1.4426950408889634f;

    VS_CONFIGURE_PER_DEVICE(
    VS_CUDA(cudaFuncSetAttribute(attention_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 DQ_SMEM));
    VS_CUDA(cudaFuncSetAttribute(attention_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 DKV_SMEM));
  );
  CUtensorMap tmQ, tmdO, tmK, tmV;
  int rc;
  // dQ: 128-row Q / dO boxes, 64-row K / V boxes
  if ((rc = make_map(&tmQ, f.Q, f.heads, f.q_rows, f.ldq, 128))) return rc;
  if ((rc = make_map(&tmdO, p->dO, f.heads, f.q_rows, p->lddo, 128))) return rc;
  if ((rc = make_map(&tmK, f.K, f.heads, f.kv_rows, f.ldk, 64))) return rc;
  if ((rc = make_map(&tmV, f.V, f.heads, f.kv_rows, f.ldv, 64))) return rc;
  {
    dim3 grid(ceil_div(f.max_q_len, 128), f.heads, f.items);
    attention_bwd_dq_kernel<<<grid, BWD_THREADS, DQ_SMEM, s>>>(tmQ, tmdO, tmK, tmV, a);
    VS_LAUNCH_CHECK();
  }
  // dK / dV: 64-row Q / dO boxes, 128-row K / V boxes
  if ((rc = make_map(&tmQ, f.Q, f.heads, f.q_rows, f.ldq, 64))) return rc;
  if ((rc = make_map(&tmdO, p->dO, f.heads, f.q_rows, p->lddo, 64))) return rc;
  if ((rc = make_map(&tmK, f.K, f.heads, f.kv_rows, f.ldk, 128))) return rc;
  if ((rc = make_map(&tmV, f.V, f.heads, f.kv_rows, f.ldv, 128))) return rc;
  {
    // upper bound of the 128-key tiles of an item: each of the two segments may end in a partial tile
    const int tiles = key_centric ? ceil_div(p->dkv_max_kv_len, 128)
                                  : ceil_div(f.max_kv_len, 128) + (f.kv_start1 ? 1 : 0);
    dim3 grid(tiles, f.heads, key_centric ? p->dkv_items : f.items);
    attention_bwd_dkv_kernel<<<grid, BWD_THREADS, DKV_SMEM, s>>>(tmQ, tmdO, tmK, tmV, a);
    VS_LAUNCH_CHECK();
  }
  return VS_OK;
}
