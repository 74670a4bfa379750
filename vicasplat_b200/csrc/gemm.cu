// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   C[m, n] = epilogue( sum_k A[m, k] * W[n, k] )      bf16 x bf16 -> fp32 (TMEM accumulator)
//
// One CTA computes one 128 x BN output tile.  Warp roles (192 threads):
//   warp 0      TMA producer: A tile (128 rows x 64 k) and W tile (BN rows x 64 k) per k-block into
//               a STAGES-deep ring of 128B-swizzled shared-memory tiles, completion on mbarriers
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (4 x K=16 UMMAs per k-block),
//               tcgen05.commit releases ring slots and finally signals the epilogue
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> shared-memory transpose
//               -> coalesced bias / activation / AdaLN gate / residual / store
// The A operand is addressed through a TMA tensor map, which is what makes the same kernel serve
// nn.Linear on (possibly strided / grouped) token rows and stride-1 kxk convolutions on NHWC maps
// (one 4-D box per filter tap; out-of-bounds coordinates are zero-filled by TMA = zero padding).
//
// Replaces in the reference: every nn.Linear in croco/blocks.py:58-130 and backbone_vica.py:57-335
// and every stride-1 nn.Conv2d in heads/dpt_block.py:79-229,264-459, heads/dpt_gs_head.py:98-157.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"
#include "tmap.h"

namespace vs {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row

struct GemmDev {
  int mode;
  int N;
  int num_kb;
  // rows mode
  int a_rows, tiles_per_group;
  // conv mode
  int cn, ch, cw, bw, bh, bn, tiles_x, tiles_y, cblocks, kw, pad;
  // epilogue
  const float* bias;
  int act;
  const float* gate;
  long long gate_ld;
  int gate_rows, first_row_mode;
  const void* res1;
  const void* res2;
  int res_dtype;
  long long res_ld;
  void* C;
  int c_dtype;
  long long ldc;
  __nv_bfloat16* C2;
  long long ldc2;
  int out_gin, out_gout, out_off;
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192)
    gemm_tc05_kernel(const __grid_constant__ CUtensorMap tmA,
                     const __grid_constant__ CUtensorMap tmW, const GemmDev g) {
  constexpr int A_BYTES = BM * 128;
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static_assert(4 * 32 * 33 * 4 <= STAGE_BYTES, "epilogue scratch must fit in stage 0");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int mt = blockIdx.y;

  // ---- tile coordinates
  int grp = 0, r0 = 0;           // rows mode
  int x0 = 0, y0 = 0, img0 = 0;  // conv mode
  if (g.mode == 0) {
    grp = mt / g.tiles_per_group;
    r0 = (mt % g.tiles_per_group) * BM;
  } else {
    const int tx = mt % g.tiles_x;
    const int ty = (mt / g.tiles_x) % g.tiles_y;
    const int tn = mt / (g.tiles_x * g.tiles_y);
    x0 = tx * g.bw;
    y0 = ty * g.bh;
    img0 = tn * g.bn;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tmem_full, 1);
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

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < g.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sA = smem + s * STAGE_BYTES;
        uint8_t* sB = sA + A_BYTES;
        mbar_expect_tx(&full[s], STAGE_BYTES);
        if (g.mode == 0) {
          tma_load_3d(sA, &tmA, &full[s], kb * BK, r0, grp);
        } else {
          const int tap = kb / g.cblocks;
          const int c0 = (kb - tap * g.cblocks) * BK;
          const int dy = tap / g.kw, dx = tap - dy * g.kw;
          tma_load_4d(sA, &tmA, &full[s], c0, x0 + dx - g.pad, y0 + dy - g.pad, img0);
        }
        tma_load_2d(sB, &tmW, &full[s], kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      for (int kb = 0; kb < g.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma_bf16_ss(tmem_base, umma_desc_k_sw128(a_addr + k * 32),
                       umma_desc_k_sw128(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);  // slot reusable once these MMAs have read it
      }
      umma_commit(tmem_full);
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    float* scratch = reinterpret_cast<float*>(smem) + q * (32 * 33);
    // row bookkeeping for tile row (q*32 + lane); broadcast by shuffle in the store loop
    long long my_out = -1;
    int my_gate = -1;
    {
      const int r = q * 32 + lane;
      bool valid;
      long long m;
      if (g.mode == 0) {
        valid = (r0 + r) < g.a_rows;
        m = static_cast<long long>(grp) * g.a_rows + r0 + r;
      } else {
        const int x = x0 + r % g.bw;
        const int y = y0 + (r / g.bw) % g.bh;
        const int im = img0 + r / (g.bw * g.bh);
        valid = x < g.cw && y < g.ch && im < g.cn;
        m = (static_cast<long long>(im) * g.ch + y) * g.cw + x;
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
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      if (n0 + c * 32 >= g.N) break;
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int n = n0 + c * 32 + lane;
      const bool nvalid = n < g.N;
      const float bias_v = (g.bias != nullptr && nvalid) ? __ldg(g.bias + n) : 0.0f;
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        const long long orow = __shfl_sync(0xffffffffu, my_out, rr);
        const int grow = __shfl_sync(0xffffffffu, my_gate, rr);
        if (orow < 0 || !nvalid) continue;
        float val = scratch[rr * 33 + lane] + bias_v;
        if (g.act == VS_ACT_GELU) val = gelu_erf(val);
        else if (g.act == VS_ACT_RELU) val = fmaxf(val, 0.0f);
        if (grow >= 0) val *= 1.0f + __ldg(g.gate + static_cast<long long>(grow) * g.gate_ld + n);
        if (g.res1 != nullptr) {
          const long long ro = orow * g.res_ld + n;
          if (g.res_dtype == VS_F32) {
            val += static_cast<const float*>(g.res1)[ro];
            if (g.res2 != nullptr) val += static_cast<const float*>(g.res2)[ro];
          } else {
            val += __bfloat162float(static_cast<const __nv_bfloat16*>(g.res1)[ro]);
            if (g.res2 != nullptr)
              val += __bfloat162float(static_cast<const __nv_bfloat16*>(g.res2)[ro]);
          }
        }
        if (g.C != nullptr) {
          if (g.c_dtype == VS_F32) static_cast<float*>(g.C)[orow * g.ldc + n] = val;
          else static_cast<__nv_bfloat16*>(g.C)[orow * g.ldc + n] = __float2bfloat16(val);
        }
        if (g.C2 != nullptr) g.C2[orow * g.ldc2 + n] = __float2bfloat16(fmaxf(val, 0.0f));
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ host side
template <int BN, int STAGES>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmDev& g, int m_tiles,
           cudaStream_t stream) {
  constexpr int SMEM = STAGES * (BM * 128 + BN * 128) + 1024 /*align*/ + 256 /*barriers*/;
  static bool configured = false;  // attribute is per-function, set once per process
  if (!configured) {
    VS_CUDA(cudaFuncSetAttribute(gemm_tc05_kernel<BN, STAGES>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  dim3 grid(ceil_div(g.N, BN), m_tiles);
  gemm_tc05_kernel<BN, STAGES><<<grid, 192, SMEM, stream>>>(tmA, tmW, g);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace
}  // namespace vs

extern "C" int vs_gemm(const vs_gemm_params* p, vs_stream_t stream_) {
  using namespace vs;
  VS_REQUIRE(p != nullptr, "vs_gemm: null params");
  VS_REQUIRE(p->A && p->W, "vs_gemm: A and W must be non-null");
  VS_REQUIRE(p->N > 0, "vs_gemm: N must be positive");
  VS_REQUIRE((reinterpret_cast<uintptr_t>(p->A) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(p->W) & 15) == 0,
             "vs_gemm: A and W must be 16-byte aligned");
  VS_REQUIRE(p->w_row_stride % 8 == 0, "vs_gemm: w_row_stride must be a multiple of 8 elements");
  cudaStream_t stream = to_stream(stream_);

  GemmDev g{};
  g.mode = p->a_mode;
  g.N = p->N;
  g.bias = p->bias;
  g.act = p->act;
  g.gate = p->gate;
  g.gate_ld = p->gate_ld;
  g.gate_rows = p->gate_rows;
  g.first_row_mode = p->first_row_mode;
  g.res1 = p->res1;
  g.res2 = p->res2;
  g.res_dtype = p->res_dtype;
  g.res_ld = p->res_ld;
  g.C = p->C;
  g.c_dtype = p->c_dtype;
  g.ldc = p->ldc;
  g.C2 = static_cast<__nv_bfloat16*>(p->C2);
  g.ldc2 = p->ldc2;
  g.out_gin = p->out_gin;
  g.out_gout = p->out_gout;
  g.out_off = p->out_off;
  VS_REQUIRE(p->res2 == nullptr || p->res1 != nullptr, "vs_gemm: res2 requires res1");
  VS_REQUIRE(p->c_dtype == VS_F32 || p->c_dtype == VS_BF16, "vs_gemm: c_dtype must be f32/bf16");
  VS_REQUIRE(p->res_dtype == VS_F32 || p->res_dtype == VS_BF16,
             "vs_gemm: res_dtype must be f32/bf16");

  CUtensorMap tmA, tmW;
  int m_tiles = 0;
  long long ktot = 0;
  if (p->a_mode == 0) {
    VS_REQUIRE(p->a_rows > 0 && p->a_groups > 0 && p->K > 0, "vs_gemm: empty problem");
    VS_REQUIRE(p->a_row_stride % 8 == 0, "vs_gemm: a_row_stride must be a multiple of 8");
    VS_REQUIRE(p->a_groups == 1 || p->a_group_stride % 8 == 0,
               "vs_gemm: a_group_stride must be a multiple of 8");
    g.a_rows = p->a_rows;
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
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(p->cin), static_cast<cuuint64_t>(p->cw),
                          static_cast<cuuint64_t>(p->ch), static_cast<cuuint64_t>(p->cn)};
    cuuint64_t str[3] = {static_cast<cuuint64_t>(p->cin) * 2,
                         static_cast<cuuint64_t>(p->cin) * p->cw * 2,
                         static_cast<cuuint64_t>(p->cin) * p->cw * p->ch * 2};
    cuuint32_t box[4] = {BK, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh),
                         static_cast<cuuint32_t>(bn)};
    int rc = encode_map(&tmA, p->A, 4, dims, str, box);
    if (rc) return rc;
  } else {
    set_error("vs_gemm: unknown a_mode %d", p->a_mode);
    return VS_ERR_INVALID;
  }
  g.num_kb = static_cast<int>((ktot + BK - 1) / BK);

  // tile width: fill the 148 SMs first, then prefer wide tiles (fewer A re-reads)
  int bn = p->block_n;
  if (bn == 0) {
    const long long t256 = static_cast<long long>(m_tiles) * ceil_div(p->N, 256);
    const long long t128 = static_cast<long long>(m_tiles) * ceil_div(p->N, 128);
    if (p->N >= 256 && t256 >= 2 * 148) bn = 256;       // 1 CTA/SM, two full waves
    else if (p->N > 64 && t128 >= 148) bn = 128;        // 2 CTAs/SM
    else bn = 64;
  }
  VS_REQUIRE(bn == 64 || bn == 128 || bn == 256, "vs_gemm: block_n must be 64, 128 or 256");
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(p->N)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(p->w_row_stride) * 2};
    cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(bn)};
    int rc = encode_map(&tmW, p->W, 2, dims, str, box);
    if (rc) return rc;
  }
  switch (bn) {
    case 64: return launch<64, 4>(tmA, tmW, g, m_tiles, stream);
    case 128: return launch<128, 3>(tmA, tmW, g, m_tiles, stream);
    default: return launch<256, 4>(tmA, tmW, g, m_tiles, stream);
  }
}
