// Bandwidth-bound kernels of the encoder's training (backward) path: the ViT block of SURVEY.md §8 E2
// (croco/blocks.py:58-130) differentiated by hand.  The dense contractions of the backward pass
// (dgrad = dY W, wgrad = dY^T X) run on the same tcgen05 GEMM as the forward pass (vs_gemm with
// transposed operand copies); what is left around them is HBM-bound and lives here:
//
//   grad_prep_kernel      one pass over a gradient matrix that emits everything the two GEMMs and the
//                         bias need: bf16 row-major copy (dgrad A operand), bf16 TRANSPOSED copy (wgrad A
//                         operand), column sums (bias gradient), optionally multiplied by gelu'(z) first
//   layernorm_backward    dx (+ residual-stream gradient), d gamma, d beta; statistics recomputed
//   gelu_kernel           training-mode GELU (the pre-activation has to be kept for the backward pass)
#include <cuda_bf16.h>

#include <algorithm>

#include "common.h"

namespace vs {
namespace {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ float4 load4(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ float4 load4(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// exact-erf GELU (nn.GELU default, croco/blocks.py:60) and its derivative, erf from Abramowitz-Stegun
// 7.1.26 (|err| <= 1.5e-7, as in the GEMM's fused GELU epilogue): with z = |x| / sqrt(2),
//   E = exp(-z^2),  R = p(t) t E,  t = 1 / (1 + 0.3275911 z)   =>   erf(z) = 1 - R
// gelu and gelu' share E (the Gaussian factor of the derivative): 2 MUFU + ~12 FMA-pipe slots per value.
__device__ __forceinline__ void erf_terms(float x, float& R, float& E) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  E = e;
  R = p * t * e;
}
__device__ __forceinline__ float gelu_f(float x) {
  float R, E;
  erf_terms(x, R, E);
  return fmaf(-0.5f * fabsf(x), R, fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float R, E;
  erf_terms(x, R, E);
  const float cdf = x >= 0.f ? fmaf(-0.5f, R, 1.0f) : 0.5f * R;   // Phi(x)
  return fmaf(x * 0.3989422804014327f, E, cdf);                    // + x phi(x)
}

// ------------------------------------------------------------------ grad_prep
// 64 x 64 tile per CTA, 256 threads.  Thread (cg = t % 16, rp = t / 16) owns columns 4cg..4cg+3 of the
// row pairs (2(rp + 16 i), +1), i = 0, 1: 16-byte (fp32) / 8-byte (bf16) loads, 128 B per row segment.
// The transposed tile goes through 8 KB of shared memory laid out [col][32 words] (one word = the
// bf16 pair of two adjacent rows) with the word index XOR-ed by 2 * (col / 4): the 4-byte writes of a
// warp hit 32 different banks and the 16-byte reads of 8 consecutive lanes 8 different bank groups.
template <typename T>
__global__ void __launch_bounds__(256)
grad_prep_kernel(const T* __restrict__ src, long long ld_src, const bf16* __restrict__ z,
                 long long ld_z, int rows, int cols, bf16* __restrict__ copy, long long ld_copy,
                 bf16* __restrict__ tr, long long ld_tr, float* __restrict__ colsum) {
  __shared__ uint32_t tile[64 * 32];
  __shared__ float csum[16][64];
  const int t = threadIdx.x;
  const int cg = t & 15, rp = t >> 4;
  const int row0 = blockIdx.x * 64, col0 = blockIdx.y * 64;   // rows on grid.x: up to 2^31 blocks
  const int c = col0 + 4 * cg;
  const bool col_ok = c < cols;  // cols % 4 == 0
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int w = rp + 16 * i;  // word (row pair) index inside the tile
    float4 v[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = row0 + 2 * w + j;
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col_ok && r < rows) {
        v[j] = load4(src + static_cast<long long>(r) * ld_src + c);
        if (z != nullptr) {
          const float4 zz = load4(z + static_cast<long long>(r) * ld_z + c);
          v[j].x *= gelu_grad_f(zz.x); v[j].y *= gelu_grad_f(zz.y);
          v[j].z *= gelu_grad_f(zz.z); v[j].w *= gelu_grad_f(zz.w);
        }
        if (copy != nullptr) {
          uint2 pk;
          pk.x = pack2(v[j].x, v[j].y);
          pk.y = pack2(v[j].z, v[j].w);
          *reinterpret_cast<uint2*>(copy + static_cast<long long>(r) * ld_copy + c) = pk;
        }
      }
      s0 += v[j].x; s1 += v[j].y; s2 += v[j].z; s3 += v[j].w;
    }
    const int pw = w ^ (2 * cg);  // swizzled word index (same for the thread's 4 columns)
    tile[(4 * cg + 0) * 32 + pw] = pack2(v[0].x, v[1].x);
    tile[(4 * cg + 1) * 32 + pw] = pack2(v[0].y, v[1].y);
    tile[(4 * cg + 2) * 32 + pw] = pack2(v[0].z, v[1].z);
    tile[(4 * cg + 3) * 32 + pw] = pack2(v[0].w, v[1].w);
  }
  if (colsum != nullptr) {
    csum[rp][4 * cg + 0] = s0; csum[rp][4 * cg + 1] = s1;
    csum[rp][4 * cg + 2] = s2; csum[rp][4 * cg + 3] = s3;
  }
  __syncthreads();
  if (colsum != nullptr && t < 64 && col0 + t < cols) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += csum[i][t];
    atomicAdd(colsum + col0 + t, s);
  }
  if (tr != nullptr) {
    // transposed rows: column cc of the source = row cc of `tr`, 64 source rows = 128 contiguous bytes
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int lc = (t >> 3) + 32 * j;  // local column
      const int k = t & 7;               // 16-byte chunk = rows 8k .. 8k+7 = words 4k .. 4k+3
      const int m = (lc >> 2) & 15;      // swizzle = XOR by 2m: group k ^ (m >> 1), halves swapped if m odd
      uint4 q = *reinterpret_cast<const uint4*>(&tile[lc * 32 + 4 * (k ^ (m >> 1))]);
      if (m & 1) {
        uint32_t a = q.x, b = q.y;
        q.x = q.z; q.y = q.w; q.z = a; q.w = b;
      }
      const int cc = col0 + lc;
      const long long r8 = row0 + 8 * k;
      if (cc < cols && r8 < ld_tr)   // pad columns [rows, ld_tr) receive the zeros of the guarded loads
        *reinterpret_cast<uint4*>(tr + static_cast<long long>(cc) * ld_tr + r8) = q;
    }
  }
}

// ------------------------------------------------------------------ GELU (training forward)
__global__ void gelu_kernel(const bf16* __restrict__ z, long long ld_z, bf16* __restrict__ a,
                            long long ld_a, int rows, int cols) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int per_row = cols >> 3;
  if (i >= static_cast<long long>(rows) * per_row) return;
  const int r = static_cast<int>(i / per_row), cb = static_cast<int>(i - static_cast<long long>(r) * per_row) * 8;
  const uint4 u = *reinterpret_cast<const uint4*>(z + static_cast<long long>(r) * ld_z + cb);
  const uint32_t in[4] = {u.x, u.y, u.z, u.w};
  uint32_t out[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&in[k]));
    out[k] = pack2(gelu_f(f.x), gelu_f(f.y));
  }
  *reinterpret_cast<uint4*>(a + static_cast<long long>(r) * ld_a + cb) =
      make_uint4(out[0], out[1], out[2], out[3]);
}

// ------------------------------------------------------------------ LayerNorm backward
// y = (x - mean) * rstd * gamma + beta  (nn.LayerNorm, eps 1e-6: croco/blocks.py:88-96).  One warp per
// row, the row in registers (C = k * 128 <= 1024), statistics recomputed from x:
//   g = dy * gamma;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat))  (+ dres, the gradient that
// reaches x through the residual connection).
// d gamma / d beta: every warp keeps a PRIVATE fp32 accumulator row pair in shared memory (8 warps x
// 2 x C floats = 64 KB at C = 1024) and adds its row's (dy * xhat, dy) with conflict-free float4
// read-modify-writes -- no barrier and no atomics inside the row loop, 128 registers so that two CTAs
// fit on an SM; the 8 accumulators are folded once at the end, one global atomicAdd per column and CTA.
template <typename T>
__global__ void __launch_bounds__(256, 2)
layernorm_backward_kernel(const float* __restrict__ x, long long ldx, const T* __restrict__ dy,
                          long long ldy, const float* __restrict__ gamma,
                          const float* dres, long long ldres, float* dx,  // may alias
                          long long lddx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                          int rows, int C, float eps) {
  extern __shared__ float4 stage[];  // [8 warps][2][C / 4]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nv = C / 128;
  const int c4 = C / 4;
  const bool params = dgamma != nullptr;
  float4* mine_g = stage + (2 * warp) * c4;
  float4* mine_b = mine_g + c4;
  if (params) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = lane; i < c4; i += 32) { mine_g[i] = zero; mine_b[i] = zero; }
    __syncwarp();
  }
  const float inv_c = 1.0f / C;
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const float* xr = x + static_cast<long long>(row) * ldx;
    const T* dyr = dy + static_cast<long long>(row) * ldy;
    float4 v[8], g[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        v[i] = load4(xr + (i * 32 + lane) * 4);
        g[i] = load4(dyr + (i * 32 + lane) * 4);
        s += v[i].x + v[i].y + v[i].z + v[i].w;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * inv_c + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        const int idx = i * 32 + lane;
        const float4 w = load4(gamma + idx * 4);
        v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
        if (params) {
          float4 a = mine_g[idx], b = mine_b[idx];
          a.x += g[i].x * v[i].x; a.y += g[i].y * v[i].y; a.z += g[i].z * v[i].z; a.w += g[i].w * v[i].w;
          b.x += g[i].x; b.y += g[i].y; b.z += g[i].z; b.w += g[i].w;
          mine_g[idx] = a;
          mine_b[idx] = b;
        }
        g[i].x *= w.x; g[i].y *= w.y; g[i].z *= w.z; g[i].w *= w.w;
        sg += g[i].x + g[i].y + g[i].z + g[i].w;
        sgx += g[i].x * v[i].x + g[i].y * v[i].y + g[i].z * v[i].z + g[i].w * v[i].w;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
    }
    const float mg = sg * inv_c, mgx = sgx * inv_c;
    float* dxr = dx + static_cast<long long>(row) * lddx;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        float4 o;
        o.x = rstd * (g[i].x - mg - v[i].x * mgx);
        o.y = rstd * (g[i].y - mg - v[i].y * mgx);
        o.z = rstd * (g[i].z - mg - v[i].z * mgx);
        o.w = rstd * (g[i].w - mg - v[i].w * mgx);
        if (dres != nullptr) {
          const float4 r = load4(dres + static_cast<long long>(row) * ldres + (i * 32 + lane) * 4);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        *reinterpret_cast<float4*>(dxr + (i * 32 + lane) * 4) = o;
      }
  }
  if (params) {
    __syncthreads();
    if (threadIdx.x < c4) {
      float4 accg = make_float4(0.f, 0.f, 0.f, 0.f), accb = accg;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float4 a = stage[(2 * w) * c4 + threadIdx.x], b = stage[(2 * w + 1) * c4 + threadIdx.x];
        accg.x += a.x; accg.y += a.y; accg.z += a.z; accg.w += a.w;
        accb.x += b.x; accb.y += b.y; accb.z += b.z; accb.w += b.w;
      }
      float* dg = dgamma + threadIdx.x * 4;
      float* db = dbeta + threadIdx.x * 4;
      atomicAdd(dg + 0, accg.x); atomicAdd(dg + 1, accg.y); atomicAdd(dg + 2, accg.z); atomicAdd(dg + 3, accg.w);
      atomicAdd(db + 0, accb.x); atomicAdd(db + 1, accb.y); atomicAdd(db + 2, accb.z); atomicAdd(db + 3, accb.w);
    }
  }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace vs

using namespace vs;

extern "C" int vs_grad_prep(const void* src, int src_dtype, int64_t ld_src, const void* z,
                            int64_t ld_z, int rows, int cols, void* copy, int64_t ld_copy,
                            void* transposed, int64_t ld_t, float* colsum, vs_stream_t stream) {
  VS_REQUIRE(src != nullptr, "grad_prep: null source");
  VS_REQUIRE(src_dtype == VS_F32 || src_dtype == VS_BF16, "grad_prep: source must be f32 or bf16");
  VS_REQUIRE(rows >= 0 && cols >= 0 && cols % 4 == 0, "grad_prep: cols must be a multiple of 4");
  if (rows == 0 || cols == 0) return VS_OK;
  VS_REQUIRE(ld_src % 4 == 0 && al16(src), "grad_prep: source rows must be 16-byte aligned");
  VS_REQUIRE(z == nullptr || (ld_z % 4 == 0 && (reinterpret_cast<uintptr_t>(z) & 7) == 0),
             "grad_prep: z rows must be 8-byte aligned");
  VS_REQUIRE(copy == nullptr || (ld_copy % 4 == 0 && (reinterpret_cast<uintptr_t>(copy) & 7) == 0),
             "grad_prep: copy rows must be 8-byte aligned");
  VS_REQUIRE(transposed == nullptr || (ld_t % 8 == 0 && ld_t >= rows && al16(transposed)),
             "grad_prep: transposed leading dimension must be a multiple of 8 and >= rows");
  VS_REQUIRE(src_dtype == VS_F32 || ld_src % 4 == 0, "grad_prep: bf16 source rows must be 8-byte aligned");
  const dim3 grid(ceil_div(rows, 64), ceil_div(cols, 64));
  VS_REQUIRE(grid.y < 65536, "grad_prep: more than 2^22 columns");
  cudaStream_t s = to_stream(stream);
  if (src_dtype == VS_F32)
    grad_prep_kernel<float><<<grid, 256, 0, s>>>(
        static_cast<const float*>(src), ld_src, static_cast<const bf16*>(z), ld_z, rows, cols,
        static_cast<bf16*>(copy), ld_copy, static_cast<bf16*>(transposed), ld_t, colsum);
  else
    grad_prep_kernel<bf16><<<grid, 256, 0, s>>>(
        static_cast<const bf16*>(src), ld_src, static_cast<const bf16*>(z), ld_z, rows, cols,
        static_cast<bf16*>(copy), ld_copy, static_cast<bf16*>(transposed), ld_t, colsum);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_gelu_bf16(const void* z, int64_t ld_z, void* a, int64_t ld_a, int rows, int cols,
                            vs_stream_t stream) {
  VS_REQUIRE(z && a, "gelu: null tensor");
  VS_REQUIRE(cols % 8 == 0 && ld_z % 8 == 0 && ld_a % 8 == 0 && al16(z) && al16(a),
             "gelu: rows must be 16-byte aligned");
  if (rows <= 0 || cols <= 0) return VS_OK;
  const long long n = static_cast<long long>(rows) * (cols / 8);
  gelu_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, to_stream(stream)>>>(
      static_cast<const bf16*>(z), ld_z, static_cast<bf16*>(a), ld_a, rows, cols);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_layernorm_backward(const vs_layernorm_bwd_params* p, vs_stream_t stream) {
  VS_REQUIRE(p && p->x && p->dy && p->gamma && p->dx, "layernorm_backward: null tensor");
  VS_REQUIRE(p->C % 128 == 0 && p->C <= 1024 && p->C > 0,
             "layernorm_backward: C must be k*128 <= 1024");
  VS_REQUIRE(p->dy_dtype == VS_F32 || p->dy_dtype == VS_BF16, "layernorm_backward: dy must be f32/bf16");
  VS_REQUIRE(p->ldx % 4 == 0 && p->ldy % 4 == 0 && p->lddx % 4 == 0 && p->ldres % 4 == 0,
             "layernorm_backward: leading dimensions must be multiples of 4");
  VS_REQUIRE((p->dgamma == nullptr) == (p->dbeta == nullptr),
             "layernorm_backward: dgamma / dbeta go together");
  if (p->rows <= 0) return VS_OK;
  int dev = 0, sms = 148;
  VS_CUDA(cudaGetDevice(&dev));
  VS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(ceil_div(p->rows, 8), 2 * sms);
  const size_t smem = 2 * 8 * sizeof(float) * p->C;   // <= 64 KB
    VS_CONFIGURE_PER_DEVICE(
    VS_CUDA(cudaFuncSetAttribute(layernorm_backward_kernel<float>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    VS_CUDA(cudaFuncSetAttribute(layernorm_backward_kernel<__nv_bfloat16>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  );
  cudaStream_t s = to_stream(stream);
  if (p->dy_dtype == VS_F32)
    layernorm_backward_kernel<float><<<grid, 256, smem, s>>>(
        p->x, p->ldx, static_cast<const float*>(p->dy), p->ldy, p->gamma, p->dres, p->ldres, p->dx,
        p->lddx, p->dgamma, p->dbeta, p->rows, p->C, p->eps);
  else
    layernorm_backward_kernel<bf16><<<grid, 256, smem, s>>>(
        p->x, p->ldx, static_cast<const bf16*>(p->dy), p->ldy, p->gamma, p->dres, p->ldres, p->dx,
        p->lddx, p->dgamma, p->dbeta, p->rows, p->C, p->eps);
  VS_LAUNCH_CHECK();
  return VS_OK;
}
