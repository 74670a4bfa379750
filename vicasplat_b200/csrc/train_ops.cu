// Bandwidth-bound kernels of the training (backward) path of the MixDecoder blocks, the DPT heads and the
// per-pixel tails (the reference gets all of these from torch.autograd):
//
//   decoder block  (backbone_vica.py:194-335)
//     layernorm_mod_backward   LayerNorm + AdaLN modulate backward with per-FRAME reductions
//     adaln_reduce             per-frame sums -> d scale / d shift / d gamma / d beta
//     gate_residual            x += (1 + gate_f) * branch                       (training forward)
//     gate_backward            d branch = dout (1 + gate_f), d gate_f = sum_rows dout . branch, bias colsum
//     silu_backward            AdaLNModulation's SiLU
//   DPT heads      (heads/dpt_block.py:79-229,264-459; heads/dpt_gs_head.py:98-157)
//     upsample2x_backward      transpose of bilinear x2 (align_corners=True) as a gather
//     pixel_unshuffle          ConvTranspose(k == stride) backward: NHWC map -> GEMM rows
//     col2im                   strided-conv dgrad: d cols -> d map (gather over the taps)
//     relu_backward            dx = dy . (y > 0) (+ column sums = bias gradient)
//   tails          (heads/postprocess.py:42-61, common/gaussian_adapter.py:167-212, vicasplat.py:179-199)
//     pts_tail_backward, gaussian_adapter_backward, camera_head_backward  (arithmetic: tail_math.h)
#include <cuda_bf16.h>

#include <algorithm>

#include "common.h"
#include "tail_math.h"

namespace vs {
namespace {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint32_t pk2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void st4_bf16(bf16* p, const float4 v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pk2(v.x, v.y), pk2(v.z, v.w));
}
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool al8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }
inline unsigned blocks_for(long long n, int per) { return static_cast<unsigned>((n + per - 1) / per); }

// ------------------------------------------------------------------ LayerNorm + modulate backward
// h = (xhat * gamma + beta) * (1 + sc_f) + sh_f   (rows of frame f; backbone_vica.py:268-278)
//   g = dh * (1 + sc_f) * gamma;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat))  (+ dres)
// and, per frame,  A_f = sum_rows dh * xhat,  B_f = sum_rows dh  -- everything the parameters need:
//   d sh_f = B_f,  d sc_f = gamma A_f + beta B_f,  d gamma = sum_f (1 + sc_f) A_f,  d beta = sum_f (1 + sc_f) B_f
// (adaln_reduce_kernel).  One CTA per (row chunk, frame): the frame index is CTA-uniform, every warp
// keeps a private fp32 accumulator pair in shared memory (as layernorm_backward_kernel does), the CTA
// folds them once and adds them to the frame's global rows.
template <typename T>
__global__ void __launch_bounds__(256, 2)
layernorm_mod_backward_kernel(const float* __restrict__ x, long long ldx, const T* __restrict__ dh,
                              long long lddh, const float* __restrict__ gamma,
                              const float* __restrict__ scale, long long mod_ld, const float* dres,
                              long long ldres, float* dx, long long lddx, float* __restrict__ frame_a,
                              float* __restrict__ frame_b, long long frame_ld, int rows_per_frame,
                              int skip_first, int rows_per_chunk, int C, float eps) {
  extern __shared__ float4 stage[];  // [8 warps][2][C / 4]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = C / 128, c4 = C / 4;
  const int frame = blockIdx.y;
  float4* mine_a = stage + (2 * warp) * c4;
  float4* mine_b = mine_a + c4;
  {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = lane; i < c4; i += 32) { mine_a[i] = zero; mine_b[i] = zero; }
    __syncwarp();
  }
  const long long row_base = static_cast<long long>(frame) * rows_per_frame;
  if (skip_first && blockIdx.x == 0 && dres != dx && warp == 0) {   // pass the first row's gradient through
    for (int i = lane; i < c4; i += 32) {
      const float4 r = dres != nullptr ? ld4(dres + row_base * ldres + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(dx + row_base * lddx + i * 4) = r;
    }
  }
  const int lo = skip_first + blockIdx.x * rows_per_chunk;
  const int hi = min(rows_per_frame, lo + rows_per_chunk);
  const float inv_c = 1.0f / C;
  const float* scf = scale != nullptr ? scale + static_cast<long long>(frame) * mod_ld : nullptr;
  for (int rr = lo + warp; rr < hi; rr += 8) {
    const long long row = row_base + rr;
    const float* xr = x + row * ldx;
    const T* dr = dh + row * lddh;
    float4 v[8], g[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        v[i] = ld4(xr + (i * 32 + lane) * 4);
        g[i] = ld4(dr + (i * 32 + lane) * 4);
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
        float4 w = ld4(gamma + idx * 4);
        if (scf != nullptr) {
          const float4 sc = ld4(scf + idx * 4);
          w.x *= 1.f + sc.x; w.y *= 1.f + sc.y; w.z *= 1.f + sc.z; w.w *= 1.f + sc.w;
        }
        v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
        float4 a = mine_a[idx], b = mine_b[idx];
        a.x += g[i].x * v[i].x; a.y += g[i].y * v[i].y; a.z += g[i].z * v[i].z; a.w += g[i].w * v[i].w;
        b.x += g[i].x; b.y += g[i].y; b.z += g[i].z; b.w += g[i].w;
        mine_a[idx] = a;
        mine_b[idx] = b;
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
    float* dxr = dx + row * lddx;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        float4 o;
        o.x = rstd * (g[i].x - mg - v[i].x * mgx);
        o.y = rstd * (g[i].y - mg - v[i].y * mgx);
        o.z = rstd * (g[i].z - mg - v[i].z * mgx);
        o.w = rstd * (g[i].w - mg - v[i].w * mgx);
        if (dres != nullptr) {
          const float4 r = ld4(dres + row * ldres + (i * 32 + lane) * 4);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        *reinterpret_cast<float4*>(dxr + (i * 32 + lane) * 4) = o;
      }
  }
  __syncthreads();
  if (threadIdx.x < c4) {
    float4 acca = make_float4(0.f, 0.f, 0.f, 0.f), accb = acca;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float4 a = stage[(2 * w) * c4 + threadIdx.x], b = stage[(2 * w + 1) * c4 + threadIdx.x];
      acca.x += a.x; acca.y += a.y; acca.z += a.z; acca.w += a.w;
      accb.x += b.x; accb.y += b.y; accb.z += b.z; accb.w += b.w;
    }
    float* fa = frame_a + static_cast<long long>(frame) * frame_ld + threadIdx.x * 4;
    float* fb = frame_b + static_cast<long long>(frame) * frame_ld + threadIdx.x * 4;
    atomicAdd(fa + 0, acca.x); atomicAdd(fa + 1, acca.y); atomicAdd(fa + 2, acca.z); atomicAdd(fa + 3, acca.w);
    atomicAdd(fb + 0, accb.x); atomicAdd(fb + 1, accb.y); atomicAdd(fb + 2, accb.z); atomicAdd(fb + 3, accb.w);
  }
}

// one thread per column: d sc_f / d sh_f written, d gamma / d beta accumulated
__global__ void adaln_reduce_kernel(const float* __restrict__ fa, const float* __restrict__ fb,
                                    long long fld, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ scale,
                                    long long mod_ld, float* __restrict__ dscale, float* __restrict__ dshift,
                                    long long dmod_ld, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                    int frames, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float gm = gamma[c], bt = beta[c];
  float ag = 0.f, ab = 0.f;
  for (int f = 0; f < frames; ++f) {
    const float a = fa[f * fld + c], b = fb[f * fld + c];
    const float sc = scale != nullptr ? scale[f * mod_ld + c] : 0.f;
    if (dscale != nullptr) {
      dscale[f * dmod_ld + c] = gm * a + bt * b;
      dshift[f * dmod_ld + c] = b;
    }
    ag += (1.f + sc) * a;
    ab += (1.f + sc) * b;
  }
  if (dgamma != nullptr) {
    dgamma[c] += ag;
    dbeta[c] += ab;
  }
}

// ------------------------------------------------------------------ gate: forward residual + backward
// first_row_mode (row 0 of each frame = the camera token): 0 as the others, 1 no gate, 2 row not touched
__global__ void gate_residual_kernel(const float* x, long long ldx, float* out, long long ldo,
                                     const bf16* __restrict__ br, long long ldb,
                                     const float* __restrict__ gate, long long gate_ld, long long rows,
                                     int C, int rows_per_frame, int first_row_mode) {
  const int c4 = C / 4;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * c4) return;
  const long long row = i / c4;
  const int c = static_cast<int>(i - row * c4) * 4;
  long long frame = 0;
  bool first = false;
  if (rows_per_frame > 0) {
    frame = row / rows_per_frame;
    first = row - frame * rows_per_frame == 0;
  }
  float4 v = ld4(x + row * ldx + c);
  if (!(first && first_row_mode == 2)) {
    const float4 b = ld4(br + row * ldb + c);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gate != nullptr && !(first && first_row_mode == 1)) g = ld4(gate + frame * gate_ld + c);
    v.x = fmaf(b.x, 1.f + g.x, v.x); v.y = fmaf(b.y, 1.f + g.y, v.y);
    v.z = fmaf(b.z, 1.f + g.z, v.z); v.w = fmaf(b.w, 1.f + g.w, v.w);
  } else if (out == x) {
    return;
  }
  *reinterpret_cast<float4*>(out + row * ldo + c) = v;
}

// grid (chunks, frames), thread = 4 columns, loops over the chunk's rows of ONE frame
__global__ void __launch_bounds__(256)
gate_backward_kernel(const float* __restrict__ dout, long long ldd, const bf16* __restrict__ br,
                     long long ldb, const float* __restrict__ gate, long long gate_ld,
                     bf16* __restrict__ dbr, long long lddb, float* __restrict__ dgate,
                     long long dgate_ld, float* __restrict__ colsum, int rows_per_frame,
                     int rows_per_chunk, int C, int first_row_mode) {
  const int c = threadIdx.x * 4;
  if (c >= C) return;
  const int frame = blockIdx.y;
  const long long row_base = static_cast<long long>(frame) * rows_per_frame;
  const int lo = blockIdx.x * rows_per_chunk, hi = min(rows_per_frame, lo + rows_per_chunk);
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gate != nullptr) g = ld4(gate + static_cast<long long>(frame) * gate_ld + c);
  const float4 g1 = make_float4(1.f + g.x, 1.f + g.y, 1.f + g.z, 1.f + g.w);
  float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), cs = dg;
#pragma unroll 4
  for (int rr = lo; rr < hi; ++rr) {
    const long long row = row_base + rr;
    float4 d = ld4(dout + row * ldd + c);
    if (rr == 0 && first_row_mode == 2) {
      d = make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (!(rr == 0 && first_row_mode == 1)) {
      if (dgate != nullptr) {
        const float4 b = ld4(br + row * ldb + c);
        dg.x = fmaf(d.x, b.x, dg.x); dg.y = fmaf(d.y, b.y, dg.y);
        dg.z = fmaf(d.z, b.z, dg.z); dg.w = fmaf(d.w, b.w, dg.w);
      }
      d.x *= g1.x; d.y *= g1.y; d.z *= g1.z; d.w *= g1.w;
    }
    cs.x += d.x; cs.y += d.y; cs.z += d.z; cs.w += d.w;
    st4_bf16(dbr + row * lddb + c, d);
  }
  if (dgate != nullptr) {
    float* p = dgate + static_cast<long long>(frame) * dgate_ld + c;
    atomicAdd(p + 0, dg.x); atomicAdd(p + 1, dg.y); atomicAdd(p + 2, dg.z); atomicAdd(p + 3, dg.w);
  }
  if (colsum != nullptr) {
    atomicAdd(colsum + c + 0, cs.x); atomicAdd(colsum + c + 1, cs.y);
    atomicAdd(colsum + c + 2, cs.z); atomicAdd(colsum + c + 3, cs.w);
  }
}

__global__ void silu_backward_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ dy,
                                     long long lddy, float* __restrict__ dx, long long lddx, int rows,
                                     int C, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const int r = i / C, c = i - r * C;
  const float v = x[r * ldx + c];
  const float s = 1.0f / (1.0f + expf(-v));
  const float d = dy[r * lddy + c] * s * (1.0f + v * (1.0f - s));
  float* o = dx + r * lddx + c;
  *o = accumulate ? *o + d : d;
}

// ------------------------------------------------------------------ DPT operators
// bilinear x2 (align_corners=True) transposed: input pixel (y, x) gathers from the output pixels whose
// taps include it.  The taps / weights are recomputed with EXACTLY the float arithmetic of
// upsample2x_kernel, so forward and backward use the same interpolation matrix.
__device__ __forceinline__ float up2_weight(int o, int i, int n_in, float step) {
  const float f = o * step;
  const int i0 = static_cast<int>(f);
  const int i1 = min(i0 + 1, n_in - 1);
  const float l = f - i0;
  return (i0 == i ? 1.f - l : 0.f) + (i1 == i ? l : 0.f);
}

__global__ void upsample2x_backward_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int n,
                                           int h, int w, int c) {
  const unsigned c8 = c / 8;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= static_cast<unsigned>(w) * c8) return;
  const int x = idx / c8, cc = idx - x * c8;
  const int y = blockIdx.y, im = blockIdx.z;
  const int ho = 2 * h, wo = 2 * w;
  const float sy = ho > 1 ? static_cast<float>(h - 1) / (ho - 1) : 0.f;
  const float sx = wo > 1 ? static_cast<float>(w - 1) / (wo - 1) : 0.f;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const bf16* base = dy + static_cast<size_t>(im) * ho * wo * c + cc * 8;
  float wx[6];
#pragma unroll
  for (int u = 0; u < 6; ++u) {
    const int ox = 2 * x - 2 + u;
    wx[u] = (ox >= 0 && ox < wo) ? up2_weight(ox, x, w, sx) : 0.f;
  }
  for (int v = 0; v < 6; ++v) {
    const int oy = 2 * y - 2 + v;
    if (oy < 0 || oy >= ho) continue;
    const float wy = up2_weight(oy, y, h, sy);
    if (wy == 0.f) continue;
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      if (wx[u] == 0.f) continue;
      const int ox = 2 * x - 2 + u;
      const float wt = wy * wx[u];
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(base + (static_cast<size_t>(oy) * wo + ox) * c));
      const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&tw[j]));
        acc[2 * j] = fmaf(wt, f.x, acc[2 * j]);
        acc[2 * j + 1] = fmaf(wt, f.y, acc[2 * j + 1]);
      }
    }
  }
  *reinterpret_cast<uint4*>(dx + ((static_cast<size_t>(im) * h + y) * w + x) * c + cc * 8) =
      make_uint4(pk2(acc[0], acc[1]), pk2(acc[2], acc[3]), pk2(acc[4], acc[5]), pk2(acc[6], acc[7]));
}

// inverse of pixel_shuffle_kernel: NHWC [n, h*k, w*k, c] -> rows [n*h*w, k*k*c] (column (dy*k+dx)*c + co)
__global__ void pixel_unshuffle_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int n, int h,
                                       int w, int c, int k) {
  const int c8 = c / 8;
  const int ho = h * k, wo = w * k;
  const long long total = static_cast<long long>(n) * ho * wo * c8;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cc = static_cast<int>(i % c8);
  long long r = i / c8;
  const int xo = static_cast<int>(r % wo); r /= wo;
  const int yo = static_cast<int>(r % ho);
  const int im = static_cast<int>(r / ho);
  const int y = yo / k, dy = yo - y * k, x = xo / k, dx = xo - x * k;
  const long long drow = (static_cast<long long>(im) * h + y) * w + x;
  const uint4 v = *reinterpret_cast<const uint4*>(src + ((static_cast<long long>(im) * ho + yo) * wo + xo) * c + cc * 8);
  *reinterpret_cast<uint4*>(dst + drow * (static_cast<long long>(k) * k * c) + (dy * k + dx) * c + cc * 8) = v;
}

// dgrad of a k x k stride-s convolution run as im2col + GEMM: d map[im, y, x, :] = sum over the taps
// (dy, dx) with (y + pad - dy) % s == 0, (x + pad - dx) % s == 0 of d cols[(im, yo, xo), tap * c + :]
__global__ void col2im_kernel(const bf16* __restrict__ dcols, bf16* __restrict__ dx, int n, int h, int w,
                              int c, int k, int stride, int pad, int kpad, int ho, int wo) {
  const unsigned c8 = c / 8;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= static_cast<unsigned>(w) * c8) return;
  const int x = idx / c8, cc = idx - x * c8;
  const int y = blockIdx.y, im = blockIdx.z;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int dy = 0; dy < k; ++dy) {
    const int ty = y + pad - dy;
    if (ty < 0 || ty % stride != 0 || ty / stride >= ho) continue;
    for (int dxx = 0; dxx < k; ++dxx) {
      const int tx = x + pad - dxx;
      if (tx < 0 || tx % stride != 0 || tx / stride >= wo) continue;
      const size_t row = (static_cast<size_t>(im) * ho + ty / stride) * wo + tx / stride;
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(dcols + row * kpad + (dy * k + dxx) * c + cc * 8));
      const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&tw[j]));
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
  }
  *reinterpret_cast<uint4*>(dx + ((static_cast<size_t>(im) * h + y) * w + x) * c + cc * 8) =
      make_uint4(pk2(acc[0], acc[1]), pk2(acc[2], acc[3]), pk2(acc[4], acc[5]), pk2(acc[6], acc[7]));
}

// dx = dy . (y > 0) on bf16 [rows, C] maps (dx may alias dy), column sums of the result accumulated.
// grid-stride over row chunks: a CTA owns 64 rows per step, thread = 8 columns of a row slice.
__global__ void __launch_bounds__(256)
relu_backward_kernel(const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ y, long long ldy,
                     bf16* dx, long long lddx, float* __restrict__ colsum, long long rows, int C) {
  const int c8 = C / 8;                 // column groups; 256 threads cover tpr = 256 / c8 rows at once (c8 <= 256)
  const int cg = threadIdx.x % c8, rs = threadIdx.x / c8;
  const int tpr = 256 / c8;
  float cs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cs[j] = 0.f;
  if (rs < tpr) {
    for (long long r = static_cast<long long>(blockIdx.x) * tpr + rs; r < rows; r += static_cast<long long>(gridDim.x) * tpr) {
      const uint4 d = *reinterpret_cast<const uint4*>(dy + r * lddy + cg * 8);
      const uint4 m = __ldg(reinterpret_cast<const uint4*>(y + r * ldy + cg * 8));
      const uint32_t dw[4] = {d.x, d.y, d.z, d.w}, mw[4] = {m.x, m.y, m.z, m.w};
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dw[j]));
        const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&mw[j]));
        if (!(g.x > 0.f)) f.x = 0.f;
        if (!(g.y > 0.f)) f.y = 0.f;
        cs[2 * j] += f.x;
        cs[2 * j + 1] += f.y;
        ow[j] = pk2(f.x, f.y);
      }
      if (dx != nullptr) *reinterpret_cast<uint4*>(dx + r * lddx + cg * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
  if (colsum != nullptr) {
    __shared__ float part[256][9];
#pragma unroll
    for (int j = 0; j < 8; ++j) part[threadIdx.x][j] = cs[j];
    __syncthreads();
    if (threadIdx.x < c8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s = 0.f;
        for (int r2 = 0; r2 < tpr; ++r2) s += part[r2 * c8 + threadIdx.x][j];
        atomicAdd(colsum + threadIdx.x * 8 + j, s);
      }
    }
  }
}

// nn.Dropout(p) of the gs head in training mode (dpt_block.py:341), in place on the bf16 post-ReLU map: an
// element is kept with probability 1 - p and scaled by 1 / (1 - p).  Counter-based: the decision is a hash
// of (seed, element index), so no generator state lives on the device and the mask never has to be stored
// -- the backward pass reads it off the kept map (dropped or inactive = 0), vs_gemm's ReLU-mask epilogue
// with out_scale = 1 / (1 - p).
__device__ __forceinline__ uint32_t mix32(uint64_t v) {
  v ^= v >> 33; v *= 0xff51afd7ed558ccdull;
  v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ull;
  v ^= v >> 33;
  return static_cast<uint32_t>(v);
}
__global__ void dropout_kernel(bf16* __restrict__ x, long long n8, uint32_t threshold, float scale,
                               uint64_t seed) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  uint4 v = *reinterpret_cast<const uint4*>(x + i * 8);
  uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    const uint32_t r0 = mix32(seed ^ (static_cast<uint64_t>(i * 8 + 2 * j) * 0x9e3779b97f4a7c15ull));
    const uint32_t r1 = mix32(seed ^ (static_cast<uint64_t>(i * 8 + 2 * j + 1) * 0x9e3779b97f4a7c15ull));
    f.x = r0 < threshold ? 0.f : f.x * scale;
    f.y = r1 < threshold ? 0.f : f.y * scale;
    w[j] = pk2(f.x, f.y);
  }
  *reinterpret_cast<uint4*>(x + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------ tails
// pts head tail backward: a = W feat + b (3-vector, recomputed), xyz = a / |a| * expm1(|a|);
// d a from d xyz (tail_math.h); d feat = (W^T d a) . (feat > 0)  [feat is the post-ReLU head.2 output];
// dW += d a (x) feat, db += d a.  Persistent blocks of Cf threads: per 128-pixel chunk every thread
// first handles one PIXEL (d a into shared memory, its d feat row), then one CHANNEL (its 3 dW sums).
__global__ void __launch_bounds__(128)
pts_tail_backward_kernel(const bf16* __restrict__ feat, int Cf, const float* __restrict__ w,
                         const float* __restrict__ b, const float* __restrict__ dxyz, long long d_ld,
                         bf16* __restrict__ dfeat, float* __restrict__ dw, float* __restrict__ db,
                         long long px) {
  extern __shared__ float sm[];   // w [3][Cf] | da [128][3]
  float* s_w = sm;
  float* s_da = sm + 3 * Cf;
  for (int i = threadIdx.x; i < 3 * Cf; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  float acc[3] = {0.f, 0.f, 0.f}, accb = 0.f;
  const long long chunks = (px + 127) / 128;
  for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
    const long long pix = ch * 128 + threadIdx.x;
    float da[3] = {0.f, 0.f, 0.f};
    if (pix < px) {
      const uint4* f = reinterpret_cast<const uint4*>(feat + pix * Cf);
      float a[3] = {b[0], b[1], b[2]};
      for (int c8 = 0; c8 < Cf / 8; ++c8) {
        const uint4 t = __ldg(f + c8);
        const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[j]));
          const int c = c8 * 8 + 2 * j;
#pragma unroll
          for (int k = 0; k < 3; ++k) a[k] = fmaf(v.x, s_w[k * Cf + c], fmaf(v.y, s_w[k * Cf + c + 1], a[k]));
        }
      }
      const float g[3] = {dxyz[pix * d_ld], dxyz[pix * d_ld + 1], dxyz[pix * d_ld + 2]};
      exp_postprocess_backward_one(a, g, da);
      uint4* o = reinterpret_cast<uint4*>(dfeat + pix * Cf);
      for (int c8 = 0; c8 < Cf / 8; ++c8) {
        const uint4 t = __ldg(f + c8);
        const uint32_t u[4] = {t.x, t.y, t.z, t.w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[j]));
          const int c = c8 * 8 + 2 * j;
          const float d0 = da[0] * s_w[c] + da[1] * s_w[Cf + c] + da[2] * s_w[2 * Cf + c];
          const float d1 = da[0] * s_w[c + 1] + da[1] * s_w[Cf + c + 1] + da[2] * s_w[2 * Cf + c + 1];
          ow[j] = pk2(v.x > 0.f ? d0 : 0.f, v.y > 0.f ? d1 : 0.f);
        }
        o[c8] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
    s_da[threadIdx.x * 3 + 0] = da[0]; s_da[threadIdx.x * 3 + 1] = da[1]; s_da[threadIdx.x * 3 + 2] = da[2];
    __syncthreads();
    // thread = channel(s): dW[k][c] += sum_p da[p][k] * feat[p][c]
    const int np = static_cast<int>(min(static_cast<long long>(128), px - ch * 128));
    for (int c = threadIdx.x; c < Cf; c += blockDim.x) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
      const bf16* fc = feat + ch * 128 * Cf + c;
      for (int p = 0; p < np; ++p) {
        const float v = __bfloat162float(fc[static_cast<long long>(p) * Cf]);
        s0 = fmaf(s_da[p * 3], v, s0); s1 = fmaf(s_da[p * 3 + 1], v, s1); s2 = fmaf(s_da[p * 3 + 2], v, s2);
      }
      if (c == threadIdx.x) { acc[0] += s0; acc[1] += s1; acc[2] += s2; }
      else { atomicAdd(dw + c, s0); atomicAdd(dw + Cf + c, s1); atomicAdd(dw + 2 * Cf + c, s2); }
    }
    if (threadIdx.x < 3) {
      float s = 0.f;
      for (int p = 0; p < np; ++p) s += s_da[p * 3 + threadIdx.x];
      accb += s;
    }
    __syncthreads();
  }
  if (threadIdx.x < Cf) {
    atomicAdd(dw + threadIdx.x, acc[0]); atomicAdd(dw + Cf + threadIdx.x, acc[1]);
    atomicAdd(dw + 2 * Cf + threadIdx.x, acc[2]);
  }
  if (threadIdx.x < 3) atomicAdd(db + threadIdx.x, accb);
}

// Gaussian adapter backward, one block per 64 Gaussians: the 11 leading scalars per Gaussian by one
// thread each (tail_math.h), the SH block as a flat, coalesced (Gaussian, coefficient) sweep.
constexpr int AB_G = 64, AB_THREADS = 256;
__global__ void __launch_bounds__(AB_THREADS)
adapter_backward_kernel(const float* __restrict__ src, long long src_ld, int center_col, int param_col,
                        long long G, int d_sh, const float* __restrict__ mask,
                        const float* __restrict__ d_raw, const float* __restrict__ d_means,
                        const float* __restrict__ d_cov, const float* __restrict__ d_cov6,
                        const float* __restrict__ d_shs, const float* __restrict__ d_opac,
                        float* __restrict__ d_src, long long dsrc_ld) {
  const long long g0 = static_cast<long long>(blockIdx.x) * AB_G;
  const int cnt = static_cast<int>(min(static_cast<long long>(AB_G), G - g0));
  const int raw_w = 11 + 3 * d_sh;
  if (threadIdx.x < cnt) {
    const long long g = g0 + threadIdx.x;
    const float* s = src + g * src_ld;
    float raw[11];
#pragma unroll
    for (int i = 0; i < 3; ++i) raw[i] = s[center_col + i];
#pragma unroll
    for (int i = 0; i < 8; ++i) raw[3 + i] = s[param_col + i];
    float dm[3] = {0.f, 0.f, 0.f}, G9[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dop = 0.f;
    if (d_means != nullptr) { dm[0] = d_means[g * 3]; dm[1] = d_means[g * 3 + 1]; dm[2] = d_means[g * 3 + 2]; }
    if (d_cov != nullptr)
#pragma unroll
      for (int i = 0; i < 9; ++i) G9[i] = d_cov[g * 9 + i];
    if (d_cov6 != nullptr) {
      const float* c6 = d_cov6 + g * 6;
      G9[0] += c6[0]; G9[1] += c6[1]; G9[2] += c6[2]; G9[4] += c6[3]; G9[5] += c6[4]; G9[8] += c6[5];
    }
    if (d_opac != nullptr) dop = d_opac[g];
    float dr[11];
    adapter_backward_general(raw, dm, G9, dop, dr);
    if (d_raw != nullptr)
#pragma unroll
      for (int i = 0; i < 11; ++i) dr[i] += d_raw[g * raw_w + i];
    float* o = d_src + g * dsrc_ld;
#pragma unroll
    for (int i = 0; i < 3; ++i) o[center_col + i] = dr[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[param_col + i] = dr[3 + i];
  }
  const int per = 3 * d_sh;
  for (int i = threadIdx.x; i < cnt * per; i += AB_THREADS) {
    const int g = i / per, j = i - g * per;
    float v = 0.f;
    if (d_shs != nullptr) v = d_shs[(g0 + g) * per + j] * __ldg(mask + (j % d_sh));
    if (d_raw != nullptr) v += d_raw[(g0 + g) * raw_w + 11 + j];
    d_src[(g0 + g) * dsrc_ld + param_col + 8 + j] = v;
  }
}

// camera head backward, one block (256 threads) per (b, t)
__global__ void camera_head_backward_kernel(const float* __restrict__ feat, long long ld,
                                            const float* __restrict__ w, const float* __restrict__ bias,
                                            int T, int C, const float* __restrict__ d_pred,
                                            float* __restrict__ d_feat, long long ldd,
                                            float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float o[8], dv[8];
  const int bt = blockIdx.x;
  const int b = bt / T, t = bt - b * T;
  float* df = d_feat + static_cast<long long>(bt) * ldd;
  if (t == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) df[c] = 0.f;
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* f = feat + static_cast<long long>(bt) * ld;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) acc += fmaxf(f[c], 0.f) * w[warp * C + c];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) o[warp] = acc + bias[warp];
  __syncthreads();
  if (threadIdx.x == 0) {
    float v[8], g[8], d[8];
    for (int i = 0; i < 8; ++i) { v[i] = o[i]; g[i] = d_pred[(static_cast<long long>(b) * (T - 1) + (t - 1)) * 8 + i]; }
    v[3] += 1.0f;
    dq_normalise_backward_one(v, g, d);
    for (int i = 0; i < 8; ++i) { dv[i] = d[i]; atomicAdd(db + i, d[i]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float fv = f[c];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s = fmaf(dv[j], w[j * C + c], s);
      if (fv > 0.f) atomicAdd(dw + j * C + c, dv[j] * fv);
    }
    df[c] = fv > 0.f ? s : 0.f;
  }
}

}  // namespace
}  // namespace vs

using namespace vs;

extern "C" int vs_layernorm_mod_backward(const vs_ln_mod_bwd_params* p, vs_stream_t stream) {
  VS_REQUIRE(p && p->x && p->dh && p->gamma && p->dx && p->frame_a && p->frame_b,
             "layernorm_mod_backward: null tensor");
  VS_REQUIRE(p->C % 128 == 0 && p->C <= 1024 && p->C > 0, "layernorm_mod_backward: C must be k*128 <= 1024");
  VS_REQUIRE(p->dh_dtype == VS_F32 || p->dh_dtype == VS_BF16, "layernorm_mod_backward: dh must be f32/bf16");
  VS_REQUIRE(p->ldx % 4 == 0 && p->lddh % 4 == 0 && p->lddx % 4 == 0 && p->ldres % 4 == 0 &&
                 p->mod_ld % 4 == 0 && p->frame_ld % 4 == 0,
             "layernorm_mod_backward: leading dimensions must be multiples of 4");
  VS_REQUIRE(p->frames >= 0 && p->rows_per_frame > 0 && (p->skip_first == 0 || p->skip_first == 1),
             "layernorm_mod_backward: bad frame layout");
  if (p->frames == 0) return VS_OK;
  const int body = p->rows_per_frame - p->skip_first;
  // enough CTAs for two per SM, at least 8 rows (one per warp) each
  int chunks = std::max(1, std::min(ceil_div(body, 8), ceil_div(2 * 148, p->frames)));
  const int per = ceil_div(std::max(body, 1), chunks);
  chunks = std::max(1, ceil_div(body, per));
  const size_t smem = 2 * 8 * sizeof(float) * p->C;
  VS_CONFIGURE_PER_DEVICE(
    VS_CUDA(cudaFuncSetAttribute(layernorm_mod_backward_kernel<float>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    VS_CUDA(cudaFuncSetAttribute(layernorm_mod_backward_kernel<__nv_bfloat16>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  );
  VS_REQUIRE(p->frames < 65536, "layernorm_mod_backward: too many frames");
  const dim3 grid(chunks, p->frames);
  cudaStream_t s = to_stream(stream);
  if (p->dh_dtype == VS_F32)
    layernorm_mod_backward_kernel<float><<<grid, 256, smem, s>>>(
        p->x, p->ldx, static_cast<const float*>(p->dh), p->lddh, p->gamma, p->scale, p->mod_ld, p->dres,
        p->ldres, p->dx, p->lddx, p->frame_a, p->frame_b, p->frame_ld, p->rows_per_frame, p->skip_first, per,
        p->C, p->eps);
  else
    layernorm_mod_backward_kernel<bf16><<<grid, 256, smem, s>>>(
        p->x, p->ldx, static_cast<const bf16*>(p->dh), p->lddh, p->gamma, p->scale, p->mod_ld, p->dres,
        p->ldres, p->dx, p->lddx, p->frame_a, p->frame_b, p->frame_ld, p->rows_per_frame, p->skip_first, per,
        p->C, p->eps);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_adaln_reduce(const float* frame_a, const float* frame_b, int64_t frame_ld,
                               const float* gamma, const float* beta, const float* scale, int64_t mod_ld,
                               float* dscale, float* dshift, int64_t dmod_ld, float* dgamma, float* dbeta,
                               int frames, int C, vs_stream_t stream) {
  VS_REQUIRE(frame_a && frame_b && gamma && beta, "adaln_reduce: null tensor");
  VS_REQUIRE((dscale == nullptr) == (dshift == nullptr) && (dgamma == nullptr) == (dbeta == nullptr),
             "adaln_reduce: dscale / dshift and dgamma / dbeta go together");
  if (frames <= 0 || C <= 0) return VS_OK;
  adaln_reduce_kernel<<<blocks_for(C, 128), 128, 0, to_stream(stream)>>>(
      frame_a, frame_b, frame_ld, gamma, beta, scale, mod_ld, dscale, dshift, dmod_ld, dgamma, dbeta, frames, C);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_gate_residual(const float* x, int64_t ldx, float* out, int64_t ldo, const void* branch,
                                int64_t ldb, const float* gate, int64_t gate_ld, int64_t rows, int C,
                                int rows_per_frame, int first_row_mode, vs_stream_t stream) {
  VS_REQUIRE(x && out && branch, "gate_residual: null tensor");
  VS_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && ldb % 4 == 0 && gate_ld % 4 == 0 && al16(x) &&
                 al16(out) && al8(branch) &&
                 (gate == nullptr || al16(gate)),
             "gate_residual: rows must be 16-byte (fp32) / 8-byte (bf16) aligned");
  VS_REQUIRE(gate == nullptr || rows_per_frame > 0, "gate_residual: a gate needs rows_per_frame");
  if (rows <= 0) return VS_OK;
  gate_residual_kernel<<<blocks_for(rows * (C / 4), 256), 256, 0, to_stream(stream)>>>(
      x, ldx, out, ldo, static_cast<const bf16*>(branch), ldb, gate, gate_ld, rows, C, rows_per_frame,
      first_row_mode);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_gate_backward(const float* dout, int64_t ldd, const void* branch, int64_t ldb,
                                const float* gate, int64_t gate_ld, void* dbranch, int64_t lddb,
                                float* dgate, int64_t dgate_ld, float* colsum, int frames,
                                int rows_per_frame, int C, int first_row_mode, vs_stream_t stream) {
  VS_REQUIRE(dout && dbranch, "gate_backward: null tensor");
  VS_REQUIRE(dgate == nullptr || (branch != nullptr && gate != nullptr), "gate_backward: dgate needs branch and gate");
  VS_REQUIRE(C % 4 == 0 && C <= 1024 && ldd % 4 == 0 && ldb % 4 == 0 && gate_ld % 4 == 0 && lddb % 4 == 0 &&
                 dgate_ld % 4 == 0 && al16(dout) && al8(dbranch) && (branch == nullptr || al8(branch)) &&
                 (gate == nullptr || al16(gate)),
             "gate_backward: C <= 1024, rows 16-byte (fp32) / 8-byte (bf16) aligned");
  VS_REQUIRE(rows_per_frame > 0 && frames < 65536, "gate_backward: bad frame layout");
  if (frames <= 0) return VS_OK;
  int chunks = std::max(1, std::min(ceil_div(rows_per_frame, 16), ceil_div(4 * 148, frames)));
  const int per = ceil_div(rows_per_frame, chunks);
  chunks = ceil_div(rows_per_frame, per);
  gate_backward_kernel<<<dim3(chunks, frames), 256, 0, to_stream(stream)>>>(
      dout, ldd, static_cast<const bf16*>(branch), ldb, gate, gate_ld, static_cast<bf16*>(dbranch), lddb,
      dgate, dgate_ld, colsum, rows_per_frame, per, C, first_row_mode);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_silu_backward(const float* x, int64_t ldx, const float* dy, int64_t lddy, float* dx,
                                int64_t lddx, int rows, int C, int accumulate, vs_stream_t stream) {
  VS_REQUIRE(x && dy && dx, "silu_backward: null tensor");
  if (rows <= 0 || C <= 0) return VS_OK;
  silu_backward_kernel<<<blocks_for(static_cast<long long>(rows) * C, 256), 256, 0, to_stream(stream)>>>(
      x, ldx, dy, lddy, dx, lddx, rows, C, accumulate);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_upsample2x_backward(const void* dy, void* dx, int n, int h, int w, int c,
                                      vs_stream_t stream) {
  VS_REQUIRE(dy && dx && c % 8 == 0 && al16(dy) && al16(dx), "upsample2x_backward: NHWC bf16, c % 8 == 0");
  if (n <= 0 || h <= 0 || w <= 0) return VS_OK;
  VS_REQUIRE(h < 65536 && n < 65536, "upsample2x_backward: map too large");
  const dim3 grid(blocks_for(static_cast<long long>(w) * (c / 8), 256), h, n);
  upsample2x_backward_kernel<<<grid, 256, 0, to_stream(stream)>>>(
      static_cast<const bf16*>(dy), static_cast<bf16*>(dx), n, h, w, c);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_pixel_unshuffle(const void* src, void* dst, int n, int h, int w, int c, int k,
                                  vs_stream_t stream) {
  VS_REQUIRE(src && dst && c % 8 == 0 && k > 0 && al16(src) && al16(dst), "pixel_unshuffle: bf16, c % 8 == 0");
  const long long total = static_cast<long long>(n) * h * k * w * k * (c / 8);
  if (total <= 0) return VS_OK;
  pixel_unshuffle_kernel<<<blocks_for(total, 256), 256, 0, to_stream(stream)>>>(
      static_cast<const bf16*>(src), static_cast<bf16*>(dst), n, h, w, c, k);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_col2im(const void* dcols, void* dx, int n, int h, int w, int c, int k, int stride, int pad,
                         int kpad, vs_stream_t stream) {
  VS_REQUIRE(dcols && dx && c % 8 == 0 && kpad % 8 == 0 && kpad >= k * k * c && stride > 0 && al16(dcols) && al16(dx),
             "col2im: bf16, c % 8 == 0, kpad >= k*k*c");
  if (n <= 0 || h <= 0 || w <= 0) return VS_OK;
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  const dim3 grid(blocks_for(static_cast<long long>(w) * (c / 8), 256), h, n);
  col2im_kernel<<<grid, 256, 0, to_stream(stream)>>>(static_cast<const bf16*>(dcols), static_cast<bf16*>(dx),
                                                      n, h, w, c, k, stride, pad, kpad, ho, wo);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_relu_backward(const void* dy, int64_t lddy, const void* y, int64_t ldy, void* dx,
                                int64_t lddx, float* colsum, int64_t rows, int C, vs_stream_t stream) {
  VS_REQUIRE(dy && y, "relu_backward: null tensor");
  VS_REQUIRE(C % 8 == 0 && C <= 2048 && lddy % 8 == 0 && ldy % 8 == 0 && lddx % 8 == 0 && al16(dy) && al16(y) &&
                 (dx == nullptr || al16(dx)),
             "relu_backward: bf16 rows, C % 8 == 0 <= 2048, 16-byte aligned");
  if (rows <= 0) return VS_OK;
  const int tpr = 256 / (C / 8);
  const long long want = (rows + tpr - 1) / tpr;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(want, 148 * 8));
  relu_backward_kernel<<<grid, 256, 0, to_stream(stream)>>>(
      static_cast<const bf16*>(dy), lddy, static_cast<const bf16*>(y), ldy, static_cast<bf16*>(dx), lddx,
      colsum, rows, C);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_dropout_bf16(void* x, int64_t n, float p, uint64_t seed, vs_stream_t stream) {
  VS_REQUIRE(x != nullptr && al16(x) && n % 8 == 0, "dropout: bf16 buffer, 16-byte aligned, n % 8 == 0");
  VS_REQUIRE(p >= 0.f && p < 1.f, "dropout: p must be in [0, 1)");
  if (n <= 0 || p == 0.f) return VS_OK;
  const uint32_t threshold = static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0);
  dropout_kernel<<<blocks_for(n / 8, 256), 256, 0, to_stream(stream)>>>(static_cast<bf16*>(x), n / 8, threshold,
                                                                          1.0f / (1.0f - p), seed);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_pts_tail_backward(const void* feat, int Cf, const float* w, const float* b,
                                    const float* d_xyz, int64_t d_ld, void* d_feat, float* dw, float* db,
                                    int64_t px, vs_stream_t stream) {
  VS_REQUIRE(feat && w && b && d_xyz && d_feat && dw && db, "pts_tail_backward: null tensor");
  VS_REQUIRE(Cf % 8 == 0 && Cf > 0 && Cf <= 1024 && al16(feat) && al16(d_feat), "pts_tail_backward: Cf % 8 == 0");
  if (px <= 0) return VS_OK;
  const unsigned grid = static_cast<unsigned>(std::min<long long>((px + 127) / 128, 148 * 16));
  const size_t smem = (3 * Cf + 128 * 3) * sizeof(float);
  pts_tail_backward_kernel<<<grid, 128, smem, to_stream(stream)>>>(
      static_cast<const bf16*>(feat), Cf, w, b, d_xyz, d_ld, static_cast<bf16*>(d_feat), dw, db, px);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_gaussian_adapter_backward(const float* src, int64_t src_ld, int center_col, int param_col,
                                            int64_t G, int d_sh, const float* sh_mask, const float* d_raw,
                                            const float* d_means, const float* d_cov, const float* d_cov6,
                                            const float* d_shs, const float* d_opac, float* d_src,
                                            int64_t dsrc_ld, vs_stream_t stream) {
  VS_REQUIRE(src && d_src && G >= 0, "gaussian_adapter_backward: null tensor");
  VS_REQUIRE(d_shs == nullptr || sh_mask != nullptr, "gaussian_adapter_backward: sh_mask required");
  VS_REQUIRE(center_col >= 0 && param_col >= 0 && src_ld >= center_col + 3 && src_ld >= param_col + 8 + 3 * d_sh &&
                 dsrc_ld >= center_col + 3 && dsrc_ld >= param_col + 8 + 3 * d_sh,
             "gaussian_adapter_backward: leading dimensions too small for the column layout");
  if (G == 0) return VS_OK;
  adapter_backward_kernel<<<blocks_for(G, AB_G), AB_THREADS, 0, to_stream(stream)>>>(
      src, src_ld, center_col, param_col, G, d_sh, sh_mask, d_raw, d_means, d_cov, d_cov6, d_shs, d_opac, d_src,
      dsrc_ld);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_camera_head_backward(const float* cam_feat, int64_t ld, const float* w, const float* b, int B,
                                       int T, int C, const float* d_pred, float* d_feat, int64_t ldd, float* dw,
                                       float* db, vs_stream_t stream) {
  VS_REQUIRE(cam_feat && w && b && d_pred && d_feat && dw && db, "camera_head_backward: null tensor");
  if (B <= 0 || T <= 0) return VS_OK;
  camera_head_backward_kernel<<<B * T, 256, 0, to_stream(stream)>>>(cam_feat, ld, w, b, T, C, d_pred, d_feat, ldd,
                                                                    dw, db);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// ------------------------------------------------------------------ LPIPS consumer (src/loss/loss_lpips.py:27-54)
// The VGG16 feature extractor of LPIPS runs on the implicit-GEMM convolution kernel (conv3x3 + bias + ReLU
// epilogue); what is left is below: 2x2 max-pooling (forward / backward) and the per-layer distance
//   d = mean_pixels sum_c w_c (f0_c / (|f0| + eps) - f1_c / (|f1| + eps))^2        (lpips: normalize_tensor,
// eps = 1e-10, 1x1 `lin` layer with non-negative weights w, spatial average), whose VALUE and GRADIENT
// w.r.t. the prediction's features f0 come out of ONE pass, like vs_mse_loss.
namespace vs {
namespace {

__global__ void maxpool2_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int n, int h, int w, int c) {
  const unsigned c8 = c / 8;
  const int ho = h / 2, wo = w / 2;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= static_cast<unsigned>(wo) * c8) return;
  const int xo = idx / c8, cc = idx - xo * c8;
  const int yo = blockIdx.y, im = blockIdx.z;
  const bf16* base = x + ((static_cast<size_t>(im) * h + 2 * yo) * w + 2 * xo) * c + cc * 8;
  uint4 q[4] = {__ldg(reinterpret_cast<const uint4*>(base)), __ldg(reinterpret_cast<const uint4*>(base + c)),
                __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(w) * c)),
                __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(w) * c + c))};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 m = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&(&q[0].x)[j]));
#pragma unroll
    for (int t = 1; t < 4; ++t) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&(&q[t].x)[j]));
      m.x = fmaxf(m.x, f.x);
      m.y = fmaxf(m.y, f.y);
    }
    o[j] = pk2(m.x, m.y);
  }
  *reinterpret_cast<uint4*>(y + ((static_cast<size_t>(im) * ho + yo) * wo + xo) * c + cc * 8) =
      make_uint4(o[0], o[1], o[2], o[3]);
}

// dx[window] = dy routed to the FIRST element of the 2x2 window that equals the pooled maximum (torch's
// tie rule), + add (the gradient arriving at x from its other consumer; nullable)
__global__ void maxpool2_backward_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y,
                                         const bf16* __restrict__ dy, const bf16* __restrict__ add,
                                         bf16* __restrict__ dx, int n, int h, int w, int c, int relu_mask) {
  const unsigned c8 = c / 8;
  const int ho = h / 2, wo = w / 2;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= static_cast<unsigned>(wo) * c8) return;
  const int xo = idx / c8, cc = idx - xo * c8;
  const int yo = blockIdx.y, im = blockIdx.z;
  const size_t o_off = ((static_cast<size_t>(im) * ho + yo) * wo + xo) * c + cc * 8;
  const size_t i_off[4] = {((static_cast<size_t>(im) * h + 2 * yo) * w + 2 * xo) * c + cc * 8, 0, 0, 0};
  const size_t offs[4] = {i_off[0], i_off[0] + c, i_off[0] + static_cast<size_t>(w) * c,
                          i_off[0] + static_cast<size_t>(w) * c + c};
  const uint4 ym = __ldg(reinterpret_cast<const uint4*>(y + o_off));
  const uint4 g = __ldg(reinterpret_cast<const uint4*>(dy + o_off));
  const uint16_t* ymh = reinterpret_cast<const uint16_t*>(&ym);
  const uint16_t* gh = reinterpret_cast<const uint16_t*>(&g);
  bool taken[8] = {false, false, false, false, false, false, false, false};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + offs[t]));
    uint4 av = make_uint4(0u, 0u, 0u, 0u);
    if (add != nullptr) av = __ldg(reinterpret_cast<const uint4*>(add + offs[t]));
    const uint16_t* xh = reinterpret_cast<const uint16_t*>(&xv);
    const uint16_t* ah = reinterpret_cast<const uint16_t*>(&av);
    uint16_t oh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = add != nullptr ? __bfloat162float(*reinterpret_cast<const bf16*>(&ah[e])) : 0.f;
      if (!taken[e] && xh[e] == ymh[e]) {
        taken[e] = true;
        // relu_mask: x is a post-ReLU map and dx is wanted w.r.t. its PRE-activation: a zero maximum passes nothing
        if (!relu_mask || (xh[e] & 0x7fffu) != 0u) v += __bfloat162float(*reinterpret_cast<const bf16*>(&gh[e]));
      }
      const bf16 b = __float2bfloat16(v);
      oh[e] = *reinterpret_cast<const uint16_t*>(&b);
    }
    *reinterpret_cast<uint4*>(dx + offs[t]) = *reinterpret_cast<const uint4*>(oh);
  }
}

// A block owns 128 consecutive pixels (of ONE image when hw % 128 == 0), a warp 16 of them; a pixel's C channels are
// spread over GROUP = min(32, C / 8) lanes with 16-byte loads (C = 64: four pixels per warp instruction).  The
// image's value is reduced in the block: one atomic per 128 pixels (one per pixel serialised 6 M same-address
// atomics on the 256^2 layer: 2.9 ms of a 3.0 ms launch).
template <int C>
__global__ void __launch_bounds__(256)
lpips_layer_kernel(const bf16* __restrict__ f0, const bf16* __restrict__ f1, const float* __restrict__ wlin,
                   long long pixels, int hw, float grad_scale, float* __restrict__ per_image,
                   bf16* __restrict__ df0) {
  constexpr int GROUP = C >= 256 ? 32 : C / 8;   // lanes per pixel
  constexpr int NV = C / (GROUP * 8);            // 16-byte vectors per lane
  constexpr int PPW = 32 / GROUP;                // pixels per warp instruction
  __shared__ float s_val[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % GROUP;                   // lane within the pixel's group
  const long long pix0 = static_cast<long long>(blockIdx.x) * 128 + warp * 16;
  float w[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int e = 0; e < 8; ++e) w[v][e] = wlin[(v * GROUP + gl) * 8 + e];
  float val_acc = 0.f;
  const bool uniform = hw % 128 == 0;
#pragma unroll 2
  for (int it = 0; it < 16 / PPW; ++it) {
    const long long pix = pix0 + it * PPW + lane / GROUP;
    const bool ok = pix < pixels;
    float a[NV][8], b[NV][8];
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      uint4 qa = make_uint4(0u, 0u, 0u, 0u), qb = qa;
      if (ok) {
        qa = __ldg(reinterpret_cast<const uint4*>(f0 + pix * C) + v * GROUP + gl);
        qb = __ldg(reinterpret_cast<const uint4*>(f1 + pix * C) + v * GROUP + gl);
      }
      const uint32_t* pa = &qa.x;
      const uint32_t* pb = &qb.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pa + e));
        const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pb + e));
        a[v][2 * e] = fa.x; a[v][2 * e + 1] = fa.y;
        b[v][2 * e] = fb.x; b[v][2 * e + 1] = fb.y;
        sa += fa.x * fa.x + fa.y * fa.y;
        sb += fb.x * fb.x + fb.y * fb.y;
      }
    }
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    const float ra = sqrtf(sa), rb = sqrtf(sb);
    const float ia = 1.0f / (ra + 1e-10f), ib = 1.0f / (rb + 1e-10f);
    float val = 0.f, dot = 0.f;   // dot = f0 . dn
    float dn[NV][8];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = a[v][e] * ia - b[v][e] * ib;
        val += w[v][e] * d * d;
        dn[v][e] = 2.f * w[v][e] * d * grad_scale;
        dot += a[v][e] * dn[v][e];
      }
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (uniform) {
      if (ok) val_acc += val;
    } else {   // small maps: a block spans several images -- one atomic per pixel
#pragma unroll
      for (int o = GROUP / 2; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
      if (ok && gl == 0) atomicAdd(per_image + pix / hw, val / hw);
    }
    if (df0 != nullptr && ok) {
      // n = f / (r + eps):  d f = (dn - f (f . dn) / (r (r + eps))) / (r + eps); then through the ReLU that
      // produced f0 (features are taken after the activation)
      const float k = ra > 0.f ? dot * ia / ra : 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        uint32_t o4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float gx = a[v][2 * e] > 0.f ? (dn[v][2 * e] - a[v][2 * e] * k) * ia : 0.f;
          const float gy = a[v][2 * e + 1] > 0.f ? (dn[v][2 * e + 1] - a[v][2 * e + 1] * k) * ia : 0.f;
          const __nv_bfloat162 h = __floats2bfloat162_rn(gx, gy);
          o4[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        reinterpret_cast<uint4*>(df0 + pix * C)[v * GROUP + gl] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) val_acc += __shfl_xor_sync(0xffffffffu, val_acc, o);
  if (lane == 0) s_val[warp] = val_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_val[i];
    const long long first = static_cast<long long>(blockIdx.x) * 128;
    if (uniform && first < pixels) atomicAdd(per_image + first / hw, t / hw);
  }
}

}  // namespace
}  // namespace vs

extern "C" int vs_maxpool2(const void* x, void* y, int n, int h, int w, int c, vs_stream_t stream) {
  VS_REQUIRE(x && y && c % 8 == 0 && h % 2 == 0 && w % 2 == 0 && al16(x) && al16(y), "maxpool2: NHWC bf16, even map, c % 8 == 0");
  if (n <= 0) return VS_OK;
  VS_REQUIRE(h / 2 < 65536 && n < 65536, "maxpool2: map too large");
  const dim3 grid(blocks_for(static_cast<long long>(w / 2) * (c / 8), 256), h / 2, n);
  maxpool2_kernel<<<grid, 256, 0, to_stream(stream)>>>(static_cast<const bf16*>(x), static_cast<bf16*>(y), n, h, w, c);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_maxpool2_backward(const void* x, const void* y, const void* dy, const void* add, void* dx, int n,
                                    int h, int w, int c, int relu_mask, vs_stream_t stream) {
  VS_REQUIRE(x && y && dy && dx && c % 8 == 0 && h % 2 == 0 && w % 2 == 0, "maxpool2_backward: NHWC bf16, even map");
  if (n <= 0) return VS_OK;
  VS_REQUIRE(h / 2 < 65536 && n < 65536, "maxpool2_backward: map too large");
  const dim3 grid(blocks_for(static_cast<long long>(w / 2) * (c / 8), 256), h / 2, n);
  maxpool2_backward_kernel<<<grid, 256, 0, to_stream(stream)>>>(
      static_cast<const bf16*>(x), static_cast<const bf16*>(y), static_cast<const bf16*>(dy),
      static_cast<const bf16*>(add), static_cast<bf16*>(dx), n, h, w, c, relu_mask);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_lpips_layer(const void* f0, const void* f1, const float* wlin, int64_t pixels, int C, int hw,
                              float grad_scale, float* per_image, void* df0, vs_stream_t stream) {
  VS_REQUIRE(f0 && f1 && wlin && per_image, "lpips_layer: null tensor");
  VS_REQUIRE((C == 64 || C == 128 || C == 256 || C == 512) && hw > 0 && pixels % hw == 0,
             "lpips_layer: C in {64, 128, 256, 512}, whole images");
  VS_REQUIRE(al16(f0) && al16(f1) && (df0 == nullptr || al16(df0)), "lpips_layer: 16-byte aligned maps");
  if (pixels <= 0) return VS_OK;
  const unsigned grid = static_cast<unsigned>((pixels + 127) / 128);
  const bf16* a = static_cast<const bf16*>(f0);
  const bf16* b = static_cast<const bf16*>(f1);
  bf16* d = static_cast<bf16*>(df0);
  cudaStream_t st = to_stream(stream);
  switch (C) {
    case 64: lpips_layer_kernel<64><<<grid, 256, 0, st>>>(a, b, wlin, pixels, hw, grad_scale, per_image, d); break;
    case 128: lpips_layer_kernel<128><<<grid, 256, 0, st>>>(a, b, wlin, pixels, hw, grad_scale, per_image, d); break;
    case 256: lpips_layer_kernel<256><<<grid, 256, 0, st>>>(a, b, wlin, pixels, hw, grad_scale, per_image, d); break;
    default: lpips_layer_kernel<512><<<grid, 256, 0, st>>>(a, b, wlin, pixels, hw, grad_scale, per_image, d); break;
  }
  VS_LAUNCH_CHECK();
  return VS_OK;
}
