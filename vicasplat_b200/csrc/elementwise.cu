// Bandwidth-bound fused ops of the encoder path (everything that is not a GEMM or attention):
// RoPE (curope replacement), LayerNorm + AdaLN modulate, patch / im2col gathers, bilinear x2,
// pixel shuffle, token initialisers, camera head, pts-head tail and the Gaussian adapter.
// All kernels are written for coalesced 16-byte accesses with the innermost (channel) dimension on
// the lanes; none of them needs shared-memory tiling (no reuse).
#include <cuda_bf16.h>

#include "common.h"
#include "half16.cuh"
#include "ptx.cuh"

namespace vs {
namespace {

using bf16 = __nv_bfloat16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half(v); }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16(v); }

// ------------------------------------------------------------------ curope.rope_2d
// One warp per token; lanes stride over the D/2 (u,v) pairs, heads in the inner loop so that
// cos/sin are computed once per pair per token (kernels.cu:44-81 computes them once per thread).
template <typename T>
__global__ void rope_2d_kernel(T* __restrict__ tokens, int B, int N, int H, int D,
                               long long stride_b, long long stride_n,
                               const long long* __restrict__ pos, float base, float fwd) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * N) return;
  const int b = warp / N, n = warp - b * N;
  const int Q = D / 4;
  T* tok = tokens + b * stride_b + n * stride_n;
  for (int p = lane; p < D / 2; p += 32) {
    const int X = p / Q, d = p - X * Q;
    const float inv_freq = fwd / powf(base, d / static_cast<float>(Q));
    const float ang = static_cast<float>(pos[(static_cast<long long>(warp)) * 2 + X]) * inv_freq;
    float s, c;
    sincosf(ang, &s, &c);
    const int iu = X * (D / 2) + d, iv = iu + Q;
    for (int h = 0; h < H; ++h) {
      const float u = to_f<T>(tok[h * D + iu]), v = to_f<T>(tok[h * D + iv]);
      tok[h * D + iu] = from_f<T>(u * c - v * s);
      tok[h * D + iv] = from_f<T>(v * c + u * s);
    }
  }
}

// Row-wise rope on a packed bf16 qkv buffer, head_dim 64: one warp per row, each lane owns the
// element pair (2*lane, 2*lane+1) of every head; the image rope's partner (e +- 16) lives in
// lane ^ 8, the camera rope's partner is inside the lane.
__global__ void rope_rows_kernel(bf16* __restrict__ qkv, long long ld, int rows, int H, int q_col,
                                 int k_col, const int* __restrict__ pos, float base,
                                 float cam_theta, float sign) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int py = pos[row * 2 + 0], px = pos[row * 2 + 1];
  const bool cam = py < 0;
  float c0, s0, c1, s1;
  bool upper = false;  // lane holds the "v" element of an image pair
  if (cam) {
    const float t = static_cast<float>(-1 - py);
    const float ang = t / powf(cam_theta, (2 * lane) / 64.0f);
    sincosf(ang, &s0, &c0);
    s0 *= sign;
    c1 = c0; s1 = s0;
  } else {
    const int e0 = 2 * lane;
    const int half = e0 >> 5;         // 0: y block, 1: x block
    const int within = e0 & 31;
    upper = within >= 16;
    const int d0 = within & 15;
    const float p = static_cast<float>(half ? px : py);
    sincosf(p / powf(base, d0 / 16.0f), &s0, &c0);
    sincosf(p / powf(base, (d0 + 1) / 16.0f), &s1, &c1);
    s0 *= sign; s1 *= sign;   // sign = -1: the inverse rotation = the transpose = the backward pass
  }
  bf16* r = qkv + static_cast<long long>(row) * ld;
  // all loads of a batch of heads are issued before the first store (q/k alias the same buffer,
  // so the compiler cannot hoist them itself): 8 independent 4-byte loads in flight per lane
  constexpr int HB = 8;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    bf16* base_ptr = r + (which == 0 ? q_col : k_col);
#pragma unroll 1
    for (int h0 = 0; h0 < H; h0 += HB) {
      float2 me[HB];
#pragma unroll
      for (int i = 0; i < HB; ++i)
        if (h0 + i < H)
          me[i] = __bfloat1622float2(
              *(reinterpret_cast<const __nv_bfloat162*>(base_ptr + (h0 + i) * 64) + lane));
#pragma unroll
      for (int i = 0; i < HB; ++i) {
        if (h0 + i < H) {   // warp-uniform
          float2 out;
          const float ox = __shfl_xor_sync(0xffffffffu, me[i].x, 8);
          const float oy = __shfl_xor_sync(0xffffffffu, me[i].y, 8);
          if (cam) {
            out.x = me[i].x * c0 - me[i].y * s0;
            out.y = me[i].y * c0 + me[i].x * s0;
          } else if (!upper) {  // me = u, other = v
            out.x = me[i].x * c0 - ox * s0;
            out.y = me[i].y * c1 - oy * s1;
          } else {              // me = v, other = u
            out.x = me[i].x * c0 + ox * s0;
            out.y = me[i].y * c1 + oy * s1;
          }
          *(reinterpret_cast<__nv_bfloat162*>(base_ptr + (h0 + i) * 64) + lane) =
              __floats2bfloat162_rn(out.x, out.y);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ LayerNorm (+ modulate)
// One warp per row, row kept in registers (C <= 1024, C % 128 == 0): two-pass mean / variance.
__global__ void layernorm_kernel(vs_layernorm_params p) {
  pdl_launch_dependents();   // the GEMM that consumes these rows may start its prologue now
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= p.rows) return;
  const int nv = p.C / 128;  // float4 per lane
  const float4* x4 = reinterpret_cast<const float4*>(p.x + static_cast<long long>(row) * p.ldx);
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) v[i] = x4[i * 32 + lane];
  float mean = 0.f, rstd = 1.f;
  if (p.normalize) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) s += v[i].x + v[i].y + v[i].z + v[i].w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mean = s / p.C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += a * a + b * b + c * c + d * d;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    rstd = rsqrtf(q / p.C + p.eps);
  }
  const bool framed = p.rows_per_frame > 0;
  const int frame = framed ? row / p.rows_per_frame : 0;
  const bool first = framed && (row - frame * p.rows_per_frame) == 0 && p.w0 != nullptr;
  const float4* w4 = reinterpret_cast<const float4*>(first ? p.w0 : p.w);
  const float4* b4 = reinterpret_cast<const float4*>(first ? p.b0 : p.b);
  const bool mod = p.scale != nullptr && !first;
  const float4* sc4 =
      mod ? reinterpret_cast<const float4*>(p.scale + static_cast<long long>(frame) * p.mod_ld)
          : nullptr;
  const float4* sh4 =
      mod ? reinterpret_cast<const float4*>(p.shift + static_cast<long long>(frame) * p.mod_ld)
          : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < nv) {
      const int idx = i * 32 + lane;
      float4 y = v[i];
      if (p.normalize) {
        y.x = (y.x - mean) * rstd; y.y = (y.y - mean) * rstd;
        y.z = (y.z - mean) * rstd; y.w = (y.w - mean) * rstd;
      }
      if (w4 != nullptr) {
        const float4 w = w4[idx], b = b4[idx];
        y.x = y.x * w.x + b.x; y.y = y.y * w.y + b.y; y.z = y.z * w.z + b.z; y.w = y.w * w.w + b.w;
      }
      if (mod) {
        const float4 sc = sc4[idx], sh = sh4[idx];
        y.x = y.x * (1.f + sc.x) + sh.x; y.y = y.y * (1.f + sc.y) + sh.y;
        y.z = y.z * (1.f + sc.z) + sh.z; y.w = y.w * (1.f + sc.w) + sh.w;
      }
      if (p.y_f32 != nullptr)
        reinterpret_cast<float4*>(p.y_f32 + static_cast<long long>(row) * p.ldy_f32)[idx] = y;
      if (p.y_bf16 != nullptr) {
        const bool f16 = p.y16_dtype == VS_F16;
        uint2 pk;
        pk.x = f2_to_h2(y.x, y.y, f16);
        pk.y = f2_to_h2(y.z, y.w, f16);
        reinterpret_cast<uint2*>(static_cast<bf16*>(p.y_bf16) +
                                 static_cast<long long>(row) * p.ldy_bf16)[idx] = pk;
      }
    }
  }
}

// ------------------------------------------------------------------ gathers
__global__ void patchify_kernel(const float* __restrict__ img, bf16* __restrict__ out, int n, int h,
                                int w, int P, int f16) {
  const long long total = static_cast<long long>(n) * 3 * h * w;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // i indexes the OUTPUT: [token][c][py][px]
  const int K = 3 * P * P;
  const long long tok = i / K;
  int k = static_cast<int>(i - tok * K);
  const int c = k / (P * P);
  k -= c * P * P;
  const int py = k / P, px = k - py * P;
  const int gw = w / P, gh = h / P;
  const int im = static_cast<int>(tok / (gw * gh));
  const int t = static_cast<int>(tok - static_cast<long long>(im) * gw * gh);
  const int ty = t / gw, tx = t - ty * gw;
  const float v = img[((static_cast<long long>(im) * 3 + c) * h + ty * P + py) * w + tx * P + px];
  reinterpret_cast<uint16_t*>(out)[i] = f_to_h(v, f16);
}

// one thread = 8 consecutive output columns (one 16-byte store); for an NHWC source whose channel
// count is a multiple of 8 those are 8 contiguous channels of one tap (one 16-byte load).
// grid = (segments of an output row of pixels, output y, image).
__global__ void im2col_kernel(const void* __restrict__ src, int nchw_f32, bf16* __restrict__ out,
                              int n, int h, int w, int c, int k, int stride, int pad, int kpad,
                              int ho, int wo, int f16) {
  const unsigned k8 = kpad / 8;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= static_cast<unsigned>(wo) * k8) return;
  const int xo = idx / k8;
  const int col0 = (idx - xo * k8) * 8;
  const int yo = blockIdx.y, im = blockIdx.z;
  const size_t row = (static_cast<size_t>(im) * ho + yo) * wo + xo;
  const int kkc = k * k * c;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (!nchw_f32 && (c & 7) == 0) {
    if (col0 < kkc) {
      const int tap = col0 / c, ch = col0 - tap * c;
      const int dy = tap / k, dx = tap - dy * k;
      const int y = yo * stride + dy - pad, x = xo * stride + dx - pad;
      if (y >= 0 && y < h && x >= 0 && x < w)
        o = __ldg(reinterpret_cast<const uint4*>(static_cast<const bf16*>(src) +
                                                 ((static_cast<size_t>(im) * h + y) * w + x) * c + ch));
    }
  } else {
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float v2[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int col = col0 + e + u;
        float v = 0.f;
        if (col < kkc) {
          const int tap = col / c, ch = col - tap * c;
          const int dy = tap / k, dx = tap - dy * k;
          const int y = yo * stride + dy - pad, x = xo * stride + dx - pad;
          if (y >= 0 && y < h && x >= 0 && x < w) {
            if (nchw_f32)
              v = __ldg(static_cast<const float*>(src) +
                        ((static_cast<size_t>(im) * c + ch) * h + y) * w + x);
            else
              v = h_to_f(static_cast<const uint16_t*>(src)[((static_cast<size_t>(im) * h + y) * w + x) * c + ch], f16);
          }
        }
        v2[u] = v;
      }
      pk[e >> 1] = f2_to_h2(v2[0], v2[1], f16);
    }
    o = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  *reinterpret_cast<uint4*>(out + row * kpad + col0) = o;
}

// bilinear x2 align_corners=True on NHWC bf16; one thread per 8 channels of UP2_ROWS vertically adjacent output
// pixels (one output per thread made the launch block-rate-bound: 262 144 blocks of 256 single-output threads ran
// at 2.2 TB/s); grid = (segments of an output row, groups of output rows, image): no 64-bit index arithmetic
constexpr int UP2_ROWS = 4;
template <bool f16, bool ADD>
__global__ void upsample2x_kernel(const bf16* __restrict__ src, const bf16* __restrict__ add,
                                  bf16* __restrict__ dst, int n, int h, int w, int c) {
  const unsigned c8 = c / 8;
  const int ho = 2 * h, wo = 2 * w;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= static_cast<unsigned>(wo) * c8) return;
  const int xo = idx / c8, cc = idx - xo * c8;
  const int im = blockIdx.z;
  const float sy = ho > 1 ? static_cast<float>(h - 1) / (ho - 1) : 0.f;
  const float sx = wo > 1 ? static_cast<float>(w - 1) / (wo - 1) : 0.f;
  const float fx = xo * sx;
  const int x0 = static_cast<int>(fx);
  const int x1 = min(x0 + 1, w - 1);
  const float lx = fx - x0;
  const bf16* base = src + static_cast<size_t>(im) * h * w * c + cc * 8;
  auto ld = [&](int y, int x) {
    return __ldg(reinterpret_cast<const uint4*>(base + (static_cast<size_t>(y) * w + x) * c));
  };
#pragma unroll
  for (int r = 0; r < UP2_ROWS; ++r) {
    const int yo = blockIdx.y * UP2_ROWS + r;
    if (yo >= ho) break;
    const float fy = yo * sy;
    const int y0 = static_cast<int>(fy);
    const int y1 = min(y0 + 1, h - 1);
    const float ly = fy - y0;
    const float w00 = (1 - ly) * (1 - lx), w01 = (1 - ly) * lx, w10 = ly * (1 - lx), w11 = ly * lx;
    const uint4 a = ld(y0, x0), b = ld(y0, x1), cq = ld(y1, x0), d = ld(y1, x1);
    uint4 o;
    const size_t oidx = ((static_cast<size_t>(im) * ho + yo) * wo + xo) * c + cc * 8;
    uint4 e = make_uint4(0u, 0u, 0u, 0u);
    if (ADD) e = __ldg(reinterpret_cast<const uint4*>(add + oidx));
    const uint32_t* pa = &a.x; const uint32_t* pb = &b.x; const uint32_t* pc = &cq.x;
    const uint32_t* pd = &d.x; const uint32_t* pe = &e.x; uint32_t* po = &o.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = h2_to_f2(pa[j], f16), fb = h2_to_f2(pb[j], f16);
      const float2 fc = h2_to_f2(pc[j], f16), fd = h2_to_f2(pd[j], f16);
      uint32_t r2 = f2_to_h2(w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fd.x,
                             w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fd.y, f16);
      if (ADD) {   // the upsampled map is rounded to 16 bits first, like the fused GEMM epilogue does
        const float2 fu = h2_to_f2(r2, f16);
        const float2 fe = h2_to_f2(pe[j], f16);
        r2 = f2_to_h2(fu.x + fe.x, fu.y + fe.y, f16);
      }
      po[j] = r2;
    }
    *reinterpret_cast<uint4*>(dst + oidx) = o;
  }
}

__global__ void pixel_shuffle_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int n,
                                     int h, int w, int c, int k) {
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
  const long long srow = (static_cast<long long>(im) * h + y) * w + x;
  const uint4 v = *reinterpret_cast<const uint4*>(src + srow * (static_cast<long long>(k) * k * c) +
                                                  (dy * k + dx) * c + cc * 8);
  *reinterpret_cast<uint4*>(dst + ((static_cast<long long>(im) * ho + yo) * wo + xo) * c + cc * 8) =
      v;
}

// thread per padded pixel: 16-byte store of (r, g, b, 0, 0, 0, 0, 0) or zeros on the border
__global__ void image_nhwc8_kernel(const float* __restrict__ img, bf16* __restrict__ out, int h,
                                   int w, int pad, int f16) {
  const int wp = w + 8;
  const int xp = blockIdx.x * blockDim.x + threadIdx.x;
  if (xp >= wp) return;
  const int yp = blockIdx.y, im = blockIdx.z;
  const int x = xp - pad, y = yp - pad;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (x >= 0 && x < w && y >= 0 && y < h) {
    const size_t hw = static_cast<size_t>(h) * w;
    const float* p = img + static_cast<size_t>(im) * 3 * hw + static_cast<size_t>(y) * w + x;
    o.x = f2_to_h2(__ldg(p), __ldg(p + hw), f16);
    o.y = f2_to_h2(__ldg(p + 2 * hw), 0.f, f16);
  }
  *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(im) * (h + 2 * pad) + yp) * wp + xp) * 8) = o;
}

__global__ void intrinsic_token_kernel(const float* __restrict__ K9, const float* __restrict__ w,
                                       const float* __restrict__ b, float* __restrict__ x,
                                       int frames, int E, int rows_per_frame, int row_off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= frames * E) return;
  const int f = i / E, e = i - f * E;
  float acc = b[e];
#pragma unroll
  for (int j = 0; j < 9; ++j) acc += K9[f * 9 + j] * w[e * 9 + j];
  x[(static_cast<long long>(f) * rows_per_frame + row_off) * E + e] = acc;
}

__global__ void camera_tokens_kernel(const float* __restrict__ intr, const float* __restrict__ extr,
                                     float* __restrict__ x, int frames, int T, int C,
                                     int rows_per_frame) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= frames * C) return;
  const int f = i / C, c = i - f * C;
  float v = intr[c];
  if (f % T != 0) v += extr[c];
  x[static_cast<long long>(f) * rows_per_frame * C + c] = v;
}

__global__ void silu_bf16_kernel(const float* __restrict__ x, long long ldx, bf16* __restrict__ y,
                                 long long ldy, int rows, int C, int f16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const int r = i / C, c = i - r * C;
  const float v = x[r * ldx + c];
  reinterpret_cast<uint16_t*>(y)[r * ldy + c] = f_to_h(v / (1.0f + expf(-v)), f16);
}

// one block (8 warps) per (b, t): warp j computes output channel j of Linear(C->8) on relu(feat)
__global__ void camera_head_kernel(const float* __restrict__ feat, long long ld,
                                   const float* __restrict__ w, const float* __restrict__ bias,
                                   int B, int T, int C, float* __restrict__ pred,
                                   float* __restrict__ c2w) {
  __shared__ float o[8];
  const int bt = blockIdx.x;
  const int b = bt / T, t = bt - b * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* M = c2w + static_cast<long long>(bt) * 16;
  if (t == 0) {
    if (threadIdx.x < 16) M[threadIdx.x] = (threadIdx.x % 5 == 0) ? 1.f : 0.f;
    return;
  }
  const float* f = feat + static_cast<long long>(bt) * ld;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) acc += fmaxf(f[c], 0.f) * w[warp * C + c];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) o[warp] = acc + bias[warp];
  __syncthreads();
  if (threadIdx.x == 0) {
    float q[8];
    for (int i = 0; i < 8; ++i) q[i] = o[i];
    q[3] += 1.0f;
    const float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 8; ++i) q[i] /= nrm;
    float* pd = pred + (static_cast<long long>(b) * (T - 1) + (t - 1)) * 8;
    for (int i = 0; i < 8; ++i) pd[i] = q[i];
    const float x = q[0], y = q[1], z = q[2], ww = q[3];
    // rotation of the unit quaternion (xyzw)
    const float R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * ww),     2 * (x * z + y * ww),
                        2 * (x * y + z * ww),     1 - 2 * (x * x + z * z), 2 * (y * z - x * ww),
                        2 * (x * z - y * ww),     2 * (y * z + x * ww),     1 - 2 * (x * x + y * y)};
    // translation = (2 q_d) * conj(q_r), vector part (Hamilton product, xyzw)
    const float x1 = 2 * q[4], y1 = 2 * q[5], z1 = 2 * q[6], w1 = 2 * q[7];
    const float x2 = -x, y2 = -y, z2 = -z, w2 = ww;
    const float tx = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2;
    const float ty = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2;
    const float tz = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2;
    M[0] = R[0]; M[1] = R[1]; M[2] = R[2];  M[3] = tx;
    M[4] = R[3]; M[5] = R[4]; M[6] = R[5];  M[7] = ty;
    M[8] = R[6]; M[9] = R[7]; M[10] = R[8]; M[11] = tz;
    M[12] = 0.f; M[13] = 0.f; M[14] = 0.f;  M[15] = 1.f;
  }
}

// thread per pixel: the pixel's Cf-channel row is read with 16-byte loads (whole sectors per
// thread), the 3 x Cf weights are broadcast from shared memory; then the exp-depth postprocess
__global__ void __launch_bounds__(128)
    pts_tail_kernel(const bf16* __restrict__ feat, int Cf, const float* __restrict__ w,
                    const float* __restrict__ b, float* __restrict__ raw, long long raw_ld,
                    long long px, int f16) {
  extern __shared__ float s_w[];   // [3][Cf]
  for (int i = threadIdx.x; i < 3 * Cf; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pix >= px) return;
  const uint4* f = reinterpret_cast<const uint4*>(feat + pix * Cf);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
  for (int c8 = 0; c8 < Cf / 8; ++c8) {
    const uint4 t = __ldg(f + c8);
    const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 v = h2_to_f2(u[j], f16);
      const int c = c8 * 8 + 2 * j;
      a0 = fmaf(v.x, s_w[c], fmaf(v.y, s_w[c + 1], a0));
      a1 = fmaf(v.x, s_w[Cf + c], fmaf(v.y, s_w[Cf + c + 1], a1));
      a2 = fmaf(v.x, s_w[2 * Cf + c], fmaf(v.y, s_w[2 * Cf + c + 1], a2));
    }
  }
  a0 += b[0]; a1 += b[1]; a2 += b[2];
  const float d = sqrtf(a0 * a0 + a1 * a1 + a2 * a2);
  const float sc = expm1f(d) / fmaxf(d, 1e-8f);
  float* o = raw + pix * raw_ld;
  o[0] = a0 * sc; o[1] = a1 * sc; o[2] = a2 * sc;
}

// Gaussian adapter (common/gaussian_adapter.py:167-212), one block per 64 Gaussians:
//   A  the 86 used columns of the 64 head rows (xyz | opacity, scale, quaternion | SH) are staged in
//      shared memory in the reference's raw layout, with coalesced loads;
//   B  64 threads turn the 11 scalars of their Gaussian into opacity / scales / rotation / covariance
//      (into shared memory), while the others already stream out raw (a flat float4 copy of the
//      staged block) and SH * mask;
//   C  the per-Gaussian outputs leave as flat, contiguous block writes as well.
// Every global store of the ~3 GB this writes per 8 scenes is a full-line coalesced access (the
// thread-per-Gaussian version scattered 4-byte stores at 12..344-byte strides).
constexpr int AD_G = 64, AD_THREADS = 256;

__device__ __forceinline__ void flat_store(float* __restrict__ dst, const float* s_src, int n, bool vec) {
  // dst: block-contiguous global range of n floats, s_src: the same range in shared memory
  if (vec) {
    for (int i = threadIdx.x; i < (n >> 2); i += AD_THREADS)
      reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_src)[i];
    for (int i = (n & ~3) + threadIdx.x; i < n; i += AD_THREADS) dst[i] = s_src[i];
  } else {
    for (int i = threadIdx.x; i < n; i += AD_THREADS) dst[i] = s_src[i];
  }
}

__global__ void __launch_bounds__(AD_THREADS)
    adapter_kernel(const float* __restrict__ src, long long src_ld, int center_col, int param_col,
                   long long G, int d_sh, const float* __restrict__ mask, float* __restrict__ raw_out,
                   float* __restrict__ means, float* __restrict__ cov, float* __restrict__ cov6,
                   float* __restrict__ sh, float* __restrict__ opac, float* __restrict__ scales,
                   float* __restrict__ rot) {
  extern __shared__ __align__(16) float ad_smem[];
  const int raw_w = 11 + 3 * d_sh;
  float* s_raw = ad_smem;                       // [64][raw_w]
  float* s_out = ad_smem + ((AD_G * raw_w + 3) & ~3);   // means 3 | cov 9 | cov6 6 | opac 1 | scales 3 | rot 4
  float* s_means = s_out, *s_cov = s_means + AD_G * 3, *s_cov6 = s_cov + AD_G * 9,
        *s_opac = s_cov6 + AD_G * 6, *s_scales = s_opac + AD_G, *s_rot = s_scales + AD_G * 3;
  const long long g0 = static_cast<long long>(blockIdx.x) * AD_G;
  const int cnt = static_cast<int>(min(static_cast<long long>(AD_G), G - g0));
  // ---- A: stage
  const float* rows = src + g0 * src_ld;
  for (int i = threadIdx.x; i < cnt * raw_w; i += AD_THREADS) {
    const int g = i / raw_w, e = i - g * raw_w;
    s_raw[i] = rows[g * src_ld + (e < 3 ? center_col + e : param_col + e - 3)];
  }
  __syncthreads();
  // ---- B: per-Gaussian parameters (threads 0..63)
  if (threadIdx.x < cnt) {
    const int g = threadIdx.x;
    const float* r = s_raw + g * raw_w;   // xyz | opacity | scale(3) | quaternion xyzw(4)
    s_means[g * 3] = r[0]; s_means[g * 3 + 1] = r[1]; s_means[g * 3 + 2] = r[2];
    s_opac[g] = 1.0f / (1.0f + expf(-r[3]));
    float sc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float x = r[4 + i];
      const float sp = x > 20.0f ? x : log1pf(expf(x));  // F.softplus (beta 1, threshold 20)
      sc[i] = fminf(0.001f * sp, 0.3f);
      s_scales[g * 3 + i] = sc[i];
    }
    float q[4] = {r[7], r[8], r[9], r[10]};
    const float qn = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
#pragma unroll
    for (int i = 0; i < 4; ++i) { q[i] /= qn; s_rot[g * 4 + i] = q[i]; }
    const float i_ = q[0], j_ = q[1], k_ = q[2], rr = q[3];
    const float two_s = 2.0f / (i_ * i_ + j_ * j_ + k_ * k_ + rr * rr + 1e-8f);
    const float R[9] = {1 - two_s * (j_ * j_ + k_ * k_), two_s * (i_ * j_ - k_ * rr),
                        two_s * (i_ * k_ + j_ * rr),     two_s * (i_ * j_ + k_ * rr),
                        1 - two_s * (i_ * i_ + k_ * k_), two_s * (j_ * k_ - i_ * rr),
                        two_s * (i_ * k_ - j_ * rr),     two_s * (j_ * k_ + i_ * rr),
                        1 - two_s * (i_ * i_ + j_ * j_)};
    const float s2[3] = {sc[0] * sc[0], sc[1] * sc[1], sc[2] * sc[2]};
    float Cm[9];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int bb = 0; bb < 3; ++bb)
        Cm[a * 3 + bb] = R[a * 3 + 0] * s2[0] * R[bb * 3 + 0] + R[a * 3 + 1] * s2[1] * R[bb * 3 + 1] +
                         R[a * 3 + 2] * s2[2] * R[bb * 3 + 2];
#pragma unroll
    for (int a = 0; a < 9; ++a) s_cov[g * 9 + a] = Cm[a];
    s_cov6[g * 6 + 0] = Cm[0]; s_cov6[g * 6 + 1] = Cm[1]; s_cov6[g * 6 + 2] = Cm[2];
    s_cov6[g * 6 + 3] = Cm[4]; s_cov6[g * 6 + 4] = Cm[5]; s_cov6[g * 6 + 5] = Cm[8];
  }
  // ---- raw and SH do not depend on B
  const bool full = cnt == AD_G;   // a full block starts 16-byte aligned in every output (64 rows)
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (raw_out) flat_store(raw_out + g0 * raw_w, s_raw, cnt * raw_w, full && al16(raw_out));
  if (sh) {
    const int per = 3 * d_sh, n = cnt * per;
    float* dst = sh + g0 * per;
    if (full && al16(sh)) {
      for (int i4 = threadIdx.x; i4 < (n >> 2); i4 += AD_THREADS) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = 4 * i4 + u, g = i / per, j = i - g * per;
          v[u] = s_raw[g * raw_w + 11 + j] * __ldg(mask + (j % d_sh));
        }
        reinterpret_cast<float4*>(dst)[i4] = make_float4(v[0], v[1], v[2], v[3]);
      }
    } else {
      for (int i = threadIdx.x; i < n; i += AD_THREADS) {
        const int g = i / per, j = i - g * per;
        dst[i] = s_raw[g * raw_w + 11 + j] * __ldg(mask + (j % d_sh));
      }
    }
  }
  __syncthreads();
  // ---- C
  if (means) flat_store(means + g0 * 3, s_means, cnt * 3, full && al16(means));
  if (cov) flat_store(cov + g0 * 9, s_cov, cnt * 9, full && al16(cov));
  if (cov6) flat_store(cov6 + g0 * 6, s_cov6, cnt * 6, full && al16(cov6));
  if (opac) flat_store(opac + g0, s_opac, cnt, full && al16(opac));
  if (scales) flat_store(scales + g0 * 3, s_scales, cnt * 3, full && al16(scales));
  if (rot) flat_store(rot + g0 * 4, s_rot, cnt * 4, full && al16(rot));
}

// ------------------------------------------------------------------ MSE loss + gradient
constexpr int MSE_BLOCKS = 148 * 8, MSE_THREADS = 256;

__global__ void __launch_bounds__(MSE_THREADS)
    mse_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long n,
                    float weight, float* __restrict__ loss_out, float* __restrict__ grad,
                    float* __restrict__ partial, unsigned int* __restrict__ counter) {
  const float gs = 2.0f * weight / static_cast<float>(n);
  float acc = 0.f;
  const long long n4 = n >> 2;
  const bool vec = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target) |
                     reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (vec) {
    for (long long i = tid; i < n4; i += stride) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(pred) + i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(target) + i);
      const float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
      acc += (d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w);
      if (grad) reinterpret_cast<float4*>(grad)[i] = make_float4(gs * d.x, gs * d.y, gs * d.z, gs * d.w);
    }
    for (long long i = n4 * 4 + tid; i < n; i += stride) {
      const float d = pred[i] - target[i];
      acc += d * d;
      if (grad) grad[i] = gs * d;
    }
  } else {
    for (long long i = tid; i < n; i += stride) {
      const float d = pred[i] - target[i];
      acc += d * d;
      if (grad) grad[i] = gs * d;
    }
  }
  __shared__ float red[MSE_THREADS / 32];
  __shared__ bool last;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < MSE_THREADS / 32; ++i) t += red[i];
    partial[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {   // fixed-order final reduction: the result does not depend on block scheduling
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += MSE_THREADS)
      t += *(volatile float*)(partial + i);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) t += __shfl_xor_sync(0xffffffffu, t, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < MSE_THREADS / 32; ++i) tot += red[i];
      *loss_out = weight * tot / static_cast<float>(n);
      *counter = 0u;   // ready for the next call on this workspace
    }
  }
}

// ------------------------------------------------------------------ pose update (one thread per camera)
__device__ __forceinline__ bool inv4(const float (&m)[16], float (&o)[16]) {
  // cofactor expansion (adjugate / determinant), as a general 4x4 inverse
  const float s0 = m[0] * m[5] - m[4] * m[1], s1 = m[0] * m[6] - m[4] * m[2];
  const float s2 = m[0] * m[7] - m[4] * m[3], s3 = m[1] * m[6] - m[5] * m[2];
  const float s4 = m[1] * m[7] - m[5] * m[3], s5 = m[2] * m[7] - m[6] * m[3];
  const float c5 = m[10] * m[15] - m[14] * m[11], c4 = m[9] * m[15] - m[13] * m[11];
  const float c3 = m[9] * m[14] - m[13] * m[10], c2 = m[8] * m[15] - m[12] * m[11];
  const float c1 = m[8] * m[14] - m[12] * m[10], c0 = m[8] * m[13] - m[12] * m[9];
  const float det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
  const float id = 1.0f / det;
  o[0] = (m[5] * c5 - m[6] * c4 + m[7] * c3) * id;
  o[1] = (-m[1] * c5 + m[2] * c4 - m[3] * c3) * id;
  o[2] = (m[13] * s5 - m[14] * s4 + m[15] * s3) * id;
  o[3] = (-m[9] * s5 + m[10] * s4 - m[11] * s3) * id;
  o[4] = (-m[4] * c5 + m[6] * c2 - m[7] * c1) * id;
  o[5] = (m[0] * c5 - m[2] * c2 + m[3] * c1) * id;
  o[6] = (-m[12] * s5 + m[14] * s2 - m[15] * s1) * id;
  o[7] = (m[8] * s5 - m[10] * s2 + m[11] * s1) * id;
  o[8] = (m[4] * c4 - m[5] * c2 + m[7] * c0) * id;
  o[9] = (-m[0] * c4 + m[1] * c2 - m[3] * c0) * id;
  o[10] = (m[12] * s4 - m[13] * s2 + m[15] * s0) * id;
  o[11] = (-m[8] * s4 + m[9] * s2 - m[11] * s0) * id;
  o[12] = (-m[4] * c3 + m[5] * c1 - m[6] * c0) * id;
  o[13] = (m[0] * c3 - m[1] * c1 + m[2] * c0) * id;
  o[14] = (-m[12] * s3 + m[13] * s1 - m[14] * s0) * id;
  o[15] = (m[8] * s3 - m[9] * s1 + m[10] * s0) * id;
  return det != 0.f;
}

__global__ void update_pose_kernel(const float* __restrict__ rho, const float* __restrict__ theta,
                                   const float* c2w, float* c2w_out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float E[16], w2c[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) E[k] = c2w[i * 16 + k];
  inv4(E, w2c);
  const float tx = theta[i * 3], ty = theta[i * 3 + 1], tz = theta[i * 3 + 2];
  const float rx = rho[i * 3], ry = rho[i * 3 + 1], rz = rho[i * 3 + 2];
  // W = skew(theta), W2 = W W
  const float W[9] = {0.f, -tz, ty, tz, 0.f, -tx, -ty, tx, 0.f};
  float W2[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      W2[r * 3 + c] = W[r * 3] * W[c] + W[r * 3 + 1] * W[3 + c] + W[r * 3 + 2] * W[6 + c];
  const float angle = sqrtf(tx * tx + ty * ty + tz * tz);
  float a, b, c_, d;   // R = I + a W + b W2,  V = I + c W + d W2
  if (angle < 1e-5f) {
    a = 1.f; b = 0.5f; c_ = 0.5f; d = 1.0f / 6.0f;
  } else {
    float sn, cs;
    sincosf(angle, &sn, &cs);
    a = sn / angle;
    b = (1.f - cs) / (angle * angle);
    c_ = b;
    d = (angle - sn) / (angle * angle * angle);
  }
  float T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  float Vm[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float I = (k % 4 == 0) ? 1.f : 0.f;
    T[(k / 3) * 4 + (k % 3)] = I + a * W[k] + b * W2[k];
    Vm[k] = I + c_ * W[k] + d * W2[k];
  }
  T[3] = Vm[0] * rx + Vm[1] * ry + Vm[2] * rz;
  T[7] = Vm[3] * rx + Vm[4] * ry + Vm[5] * rz;
  T[11] = Vm[6] * rx + Vm[7] * ry + Vm[8] * rz;
  float nw[16], out[16];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c)
      nw[r * 4 + c] = T[r * 4] * w2c[c] + T[r * 4 + 1] * w2c[4 + c] + T[r * 4 + 2] * w2c[8 + c] +
                      T[r * 4 + 3] * w2c[12 + c];
  inv4(nw, out);
#pragma unroll
  for (int k = 0; k < 16; ++k) c2w_out[i * 16 + k] = out[k];
}

// ------------------------------------------------------------------ fused AdamW
constexpr int AW_THREADS = 256;

// torch.nan_to_num_ (the reference's GradientNanCheckCallback, src/main.py:40-45): NaN -> 0, +-inf -> +-FLT_MAX
__device__ __forceinline__ float sanitize(float v) {
  if (isnan(v)) return 0.f;
  if (isinf(v)) return v > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
  return v;
}

__global__ void __launch_bounds__(AW_THREADS)
    adamw_norm_kernel(const vs_adamw_params p) {
  const int c = blockIdx.x;
  const int t = p.chunk_tensor[c];
  const long long start = p.chunk_start[c];
  const long long n = min(static_cast<long long>(VS_ADAMW_CHUNK), p.sizes[t] - start);
  const float* g = p.grads[t] + start;
  float acc = 0.f;
  bool bad = false;
  for (long long i = threadIdx.x; i < n; i += AW_THREADS) {
    float v = g[i];
    bad |= !isfinite(v);
    if (p.skip_nonfinite == 2) v = sanitize(v);
    acc = fmaf(v, v, acc);
  }
  __shared__ float red[AW_THREADS / 32];
  __shared__ int s_bad;
  __shared__ bool last;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  if (bad) s_bad = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < AW_THREADS / 32; ++i) tot += red[i];
    p.partials[c] = tot;
    if (s_bad) atomicExch(p.found_inf_out, 1);
    __threadfence();
    last = atomicAdd(p.counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {   // fixed-order final reduction (independent of block scheduling)
    __threadfence();
    float tsum = 0.f;
    for (int i = threadIdx.x; i < p.n_chunks; i += AW_THREADS) tsum += *(volatile float*)(p.partials + i);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < AW_THREADS / 32; ++i) tot += red[i];
      *p.grad_norm_out = sqrtf(tot);
      *p.counter = 0u;
      // the step counter lives on the device: it advances only when the update is applied, so the bias
      // corrections never run ahead of the moments after a skipped step
      if (p.step_counter != nullptr && !(p.skip_nonfinite == 1 && *(volatile int*)p.found_inf_out))
        *p.step_counter += 1;
    }
  }
}

__global__ void __launch_bounds__(AW_THREADS)
    adamw_update_kernel(const vs_adamw_params p, float bc1, float bc2_sqrt) {
  if (p.skip_nonfinite == 1 && *p.found_inf_out) return;
  if (p.step_counter != nullptr) {
    const float t = static_cast<float>(*p.step_counter);
    bc1 = 1.0f - powf(p.beta1, t);
    bc2_sqrt = sqrtf(1.0f - powf(p.beta2, t));
  }
  const int c = blockIdx.x;
  const int t = p.chunk_tensor[c];
  const long long start = p.chunk_start[c];
  const long long n = min(static_cast<long long>(VS_ADAMW_CHUNK), p.sizes[t] - start);
  float* w = p.params[t] + start;
  const float* g = p.grads[t] + start;
  float* m = p.exp_avg[t] + start;
  float* v = p.exp_avg_sq[t] + start;
  const float lr = p.lrs[t];
  float clip = 1.0f;
  if (p.max_grad_norm > 0.f) clip = fminf(1.0f, p.max_grad_norm / (*p.grad_norm_out + 1e-6f));
  const float decay = 1.0f - lr * p.weight_decay;
  const float step_size = lr / bc1;
  for (long long i = threadIdx.x; i < n; i += AW_THREADS) {
    const float gi = (p.skip_nonfinite == 2 ? sanitize(g[i]) : g[i]) * clip;
    const float mi = p.beta1 * m[i] + (1.0f - p.beta1) * gi;
    const float vi = p.beta2 * v[i] + (1.0f - p.beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + p.eps;
    w[i] = w[i] * decay - step_size * (mi / denom);
  }
}

inline unsigned blocks_for(long long n, int threads) {
  return static_cast<unsigned>((n + threads - 1) / threads);
}

}  // namespace
}  // namespace vs

using namespace vs;

extern "C" int vs_rope_2d(void* tokens, int dtype, int B, int N, int H, int D, int64_t stride_b,
                          int64_t stride_n, const int64_t* positions, float base, float fwd,
                          vs_stream_t stream) {
  VS_REQUIRE(D % 4 == 0, "token dim must be multiple of 4");  // kernels.cu:94
  VS_REQUIRE(B >= 0 && N >= 0 && H >= 0, "rope_2d: negative size");
  if (B * N == 0 || H == 0) return VS_OK;
  VS_REQUIRE(tokens && positions, "rope_2d: null tensor");
  const int threads = 256;
  const unsigned blocks = blocks_for(static_cast<long long>(B) * N * 32, threads);
  cudaStream_t s = to_stream(stream);
  const long long* pos = reinterpret_cast<const long long*>(positions);
  if (dtype == VS_F32)
    rope_2d_kernel<float><<<blocks, threads, 0, s>>>(static_cast<float*>(tokens), B, N, H, D,
                                                     stride_b, stride_n, pos, base, fwd);
  else if (dtype == VS_F16)
    rope_2d_kernel<__half><<<blocks, threads, 0, s>>>(static_cast<__half*>(tokens), B, N, H, D,
                                                      stride_b, stride_n, pos, base, fwd);
  else if (dtype == VS_BF16)
    rope_2d_kernel<bf16><<<blocks, threads, 0, s>>>(static_cast<bf16*>(tokens), B, N, H, D,
                                                    stride_b, stride_n, pos, base, fwd);
  else {
    set_error("rope_2d: unsupported dtype %d", dtype);
    return VS_ERR_UNSUPPORTED;
  }
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_rope_rows(void* qkv, int64_t ld, int rows, int H, int q_col, int k_col,
                            const int32_t* pos, float base, float cam_theta, vs_stream_t stream) {
  VS_REQUIRE(qkv && pos, "rope_rows: null tensor");
  VS_REQUIRE(ld % 2 == 0 && q_col % 2 == 0 && k_col % 2 == 0, "rope_rows: columns must be even");
  if (rows <= 0) return VS_OK;
  rope_rows_kernel<<<blocks_for(static_cast<long long>(rows) * 32, 256), 256, 0,
                     to_stream(stream)>>>(static_cast<bf16*>(qkv), ld, rows, H, q_col, k_col, pos,
                                          base, cam_theta, 1.0f);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_rope_rows_backward(void* dqkv, int64_t ld, int rows, int H, int q_col, int k_col,
                                     const int32_t* pos, float base, float cam_theta,
                                     vs_stream_t stream) {
  VS_REQUIRE(dqkv && pos, "rope_rows_backward: null tensor");
  VS_REQUIRE(ld % 2 == 0 && q_col % 2 == 0 && k_col % 2 == 0, "rope_rows_backward: columns must be even");
  if (rows <= 0) return VS_OK;
  rope_rows_kernel<<<blocks_for(static_cast<long long>(rows) * 32, 256), 256, 0,
                     to_stream(stream)>>>(static_cast<bf16*>(dqkv), ld, rows, H, q_col, k_col, pos,
                                          base, cam_theta, -1.0f);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_layernorm(const vs_layernorm_params* p, vs_stream_t stream) {
  VS_REQUIRE(p && p->x, "layernorm: null input");
  VS_REQUIRE(p->C % 128 == 0 && p->C <= 1024 && p->C > 0, "layernorm: C must be k*128 <= 1024");
  VS_REQUIRE(p->ldx % 4 == 0 && p->ldy_f32 % 4 == 0 && p->ldy_bf16 % 4 == 0 && p->mod_ld % 4 == 0,
             "layernorm: leading dimensions must be multiples of 4");
  VS_REQUIRE((p->scale == nullptr) == (p->shift == nullptr), "layernorm: scale/shift go together");
  VS_REQUIRE(p->scale == nullptr || p->rows_per_frame > 0, "layernorm: modulation needs frames");
  if (p->rows <= 0) return VS_OK;
  layernorm_kernel<<<blocks_for(static_cast<long long>(p->rows) * 32, 256), 256, 0,
                     to_stream(stream)>>>(*p);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_patchify(const float* img, void* out, int n, int h, int w, int P, int half_dtype,
                           vs_stream_t stream) {
  VS_REQUIRE(half_dtype == 0 || half_dtype == VS_BF16 || half_dtype == VS_F16, "half_dtype must be bf16 or fp16");
  const int f16 = half_dtype == VS_F16;
  VS_REQUIRE(img && out, "patchify: null tensor");
  VS_REQUIRE(P > 0 && h % P == 0 && w % P == 0, "Input image size is not a multiple of patch size");
  const long long total = static_cast<long long>(n) * 3 * h * w;
  if (total == 0) return VS_OK;
  patchify_kernel<<<blocks_for(total, 256), 256, 0, to_stream(stream)>>>(
      img, static_cast<bf16*>(out), n, h, w, P, f16);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_im2col(const void* src, int src_nchw_f32, void* out, int n, int h, int w, int c,
                         int k, int stride, int pad, int kpad, int half_dtype, vs_stream_t stream) {
  VS_REQUIRE(half_dtype == 0 || half_dtype == VS_BF16 || half_dtype == VS_F16, "half_dtype must be bf16 or fp16");
  const int f16 = half_dtype == VS_F16;
  VS_REQUIRE(src && out, "im2col: null tensor");
  VS_REQUIRE(k > 0 && stride > 0 && kpad >= k * k * c, "im2col: bad geometry");
  VS_REQUIRE(kpad % 8 == 0, "im2col: kpad must be a multiple of 8");
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  if (static_cast<long long>(n) * ho * wo <= 0) return VS_OK;
  VS_REQUIRE(ho <= 65535 && n <= 65535, "im2col: map too large for the launch grid");
  dim3 grid(blocks_for(static_cast<long long>(wo) * (kpad / 8), 256), ho, n);
  im2col_kernel<<<grid, 256, 0, to_stream(stream)>>>(
      src, src_nchw_f32, static_cast<bf16*>(out), n, h, w, c, k, stride, pad, kpad, ho, wo, f16);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_upsample2x(const void* src, void* dst, int n, int h, int w, int c, int half_dtype,
                             vs_stream_t stream) {
  VS_REQUIRE(half_dtype == 0 || half_dtype == VS_BF16 || half_dtype == VS_F16, "half_dtype must be bf16 or fp16");
  const int f16 = half_dtype == VS_F16;
  VS_REQUIRE(src && dst, "upsample2x: null tensor");
  VS_REQUIRE(c % 8 == 0, "upsample2x: channels must be a multiple of 8");
  if (static_cast<long long>(n) * h * w == 0) return VS_OK;
  VS_REQUIRE(2 * h <= 65535 && n <= 65535, "upsample2x: map too large for the launch grid");
  dim3 grid(blocks_for(static_cast<long long>(2 * w) * (c / 8), 256), (2 * h + UP2_ROWS - 1) / UP2_ROWS, n);
  (f16 ? upsample2x_kernel<true, false> : upsample2x_kernel<false, false>)<<<grid, 256, 0, to_stream(stream)>>>(
      static_cast<const bf16*>(src), nullptr, static_cast<bf16*>(dst), n, h, w, c);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_upsample2x_add(const void* src, const void* add, void* dst, int n, int h, int w, int c,
                                 int half_dtype, vs_stream_t stream) {
  VS_REQUIRE(half_dtype == 0 || half_dtype == VS_BF16 || half_dtype == VS_F16, "half_dtype must be bf16 or fp16");
  const int f16 = half_dtype == VS_F16;
  VS_REQUIRE(src && add && dst, "upsample2x_add: null tensor");
  VS_REQUIRE(c % 8 == 0, "upsample2x_add: channels must be a multiple of 8");
  if (static_cast<long long>(n) * h * w == 0) return VS_OK;
  VS_REQUIRE(2 * h <= 65535 && n <= 65535, "upsample2x_add: map too large for the launch grid");
  dim3 grid(blocks_for(static_cast<long long>(2 * w) * (c / 8), 256), (2 * h + UP2_ROWS - 1) / UP2_ROWS, n);
  (f16 ? upsample2x_kernel<true, true> : upsample2x_kernel<false, true>)<<<grid, 256, 0, to_stream(stream)>>>(
      static_cast<const bf16*>(src), static_cast<const bf16*>(add), static_cast<bf16*>(dst), n, h, w, c);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_pixel_shuffle(const void* src, void* dst, int n, int h, int w, int c, int k,
                                vs_stream_t stream) {
  VS_REQUIRE(src && dst, "pixel_shuffle: null tensor");
  VS_REQUIRE(c % 8 == 0 && k > 0, "pixel_shuffle: channels must be a multiple of 8");
  const long long total = static_cast<long long>(n) * h * k * w * k * (c / 8);
  if (total == 0) return VS_OK;
  pixel_shuffle_kernel<<<blocks_for(total, 256), 256, 0, to_stream(stream)>>>(
      static_cast<const bf16*>(src), static_cast<bf16*>(dst), n, h, w, c, k);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_image_nhwc8(const float* img, void* out, int n, int h, int w, int pad, int half_dtype,
                              vs_stream_t stream) {
  VS_REQUIRE(half_dtype == 0 || half_dtype == VS_BF16 || half_dtype == VS_F16, "half_dtype must be bf16 or fp16");
  const int f16 = half_dtype == VS_F16;
  VS_REQUIRE(img && out, "image_nhwc8: null tensor");
  VS_REQUIRE(pad >= 0 && pad <= 8 && h + 2 * pad <= 65535 && n <= 65535, "image_nhwc8: bad geometry");
  if (n * h * w == 0) return VS_OK;
  dim3 grid(blocks_for(w + 8, 128), h + 2 * pad, n);
  image_nhwc8_kernel<<<grid, 128, 0, to_stream(stream)>>>(img, static_cast<bf16*>(out), h, w, pad, f16);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_intrinsic_token(const float* K9, const float* w, const float* b, float* x,
                                  int frames, int E, int rows_per_frame, int row_off,
                                  vs_stream_t stream) {
  VS_REQUIRE(K9 && w && b && x, "intrinsic_token: null tensor");
  if (frames * E == 0) return VS_OK;
  intrinsic_token_kernel<<<blocks_for(static_cast<long long>(frames) * E, 256), 256, 0,
                           to_stream(stream)>>>(K9, w, b, x, frames, E, rows_per_frame, row_off);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_camera_tokens(const float* intr_tok, const float* extr_tok, float* x, int frames,
                                int T, int C, int rows_per_frame, vs_stream_t stream) {
  VS_REQUIRE(intr_tok && extr_tok && x && T > 0, "camera_tokens: bad arguments");
  if (frames * C == 0) return VS_OK;
  camera_tokens_kernel<<<blocks_for(static_cast<long long>(frames) * C, 256), 256, 0,
                         to_stream(stream)>>>(intr_tok, extr_tok, x, frames, T, C, rows_per_frame);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_silu_bf16(const float* x, int64_t ldx, void* y, int64_t ldy, int rows, int C,
                            int half_dtype, vs_stream_t stream) {
  VS_REQUIRE(half_dtype == 0 || half_dtype == VS_BF16 || half_dtype == VS_F16, "half_dtype must be bf16 or fp16");
  const int f16 = half_dtype == VS_F16;
  VS_REQUIRE(x && y, "silu: null tensor");
  if (rows * C == 0) return VS_OK;
  silu_bf16_kernel<<<blocks_for(static_cast<long long>(rows) * C, 256), 256, 0,
                     to_stream(stream)>>>(x, ldx, static_cast<bf16*>(y), ldy, rows, C, f16);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_camera_head(const float* cam_feat, int64_t ld, const float* w, const float* b,
                              int B, int T, int C, float* pred_dq, float* c2w, vs_stream_t stream) {
  VS_REQUIRE(cam_feat && w && b && pred_dq && c2w, "camera_head: null tensor");
  VS_REQUIRE(T >= 1 && B >= 1, "camera_head: bad sizes");
  camera_head_kernel<<<B * T, 256, 0, to_stream(stream)>>>(cam_feat, ld, w, b, B, T, C, pred_dq,
                                                           c2w);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_pts_tail(const void* feat, int Cf, const float* w, const float* b, float* raw,
                           int64_t raw_ld, int64_t px, int half_dtype, vs_stream_t stream) {
  VS_REQUIRE(half_dtype == 0 || half_dtype == VS_BF16 || half_dtype == VS_F16, "half_dtype must be bf16 or fp16");
  const int f16 = half_dtype == VS_F16;
  VS_REQUIRE(feat && w && b && raw, "pts_tail: null tensor");
  VS_REQUIRE(Cf % 64 == 0, "pts_tail: Cf must be a multiple of 64");
  if (px == 0) return VS_OK;
  VS_REQUIRE(Cf <= 1024, "pts_tail: Cf too large");
  pts_tail_kernel<<<blocks_for(px, 128), 128, 3 * Cf * sizeof(float), to_stream(stream)>>>(
      static_cast<const bf16*>(feat), Cf, w, b, raw, raw_ld, px, f16);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_gaussian_adapter(const float* src, int64_t src_ld, int center_col, int param_col,
                                   int64_t G, int d_sh, const float* sh_mask, float* raw_out,
                                   float* means, float* cov, float* cov6, float* sh, float* opac,
                                   float* scales, float* rot, vs_stream_t stream) {
  VS_REQUIRE(src, "gaussian_adapter: null input");
  VS_REQUIRE(center_col >= 0 && param_col >= 0 && src_ld >= center_col + 3 &&
                 src_ld >= param_col + 8 + 3 * d_sh,
             "gaussian_adapter: src_ld too small for the column layout");
  if (G == 0) return VS_OK;
  VS_REQUIRE(sh == nullptr || sh_mask != nullptr, "gaussian_adapter: sh_mask required");
  VS_REQUIRE(d_sh >= 0 && d_sh <= 49, "gaussian_adapter: d_sh out of range");
  const int raw_w = 11 + 3 * d_sh;
  const size_t smem = (static_cast<size_t>((AD_G * raw_w + 3) & ~3) + AD_G * 26) * sizeof(float);
    VS_CONFIGURE_PER_DEVICE(
    VS_CUDA(cudaFuncSetAttribute(adapter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  );
  adapter_kernel<<<blocks_for(G, AD_G), AD_THREADS, smem, to_stream(stream)>>>(
      src, src_ld, center_col, param_col, G, d_sh, sh_mask, raw_out, means, cov, cov6, sh, opac,
      scales, rot);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int64_t vs_mse_workspace_bytes(void) { return (vs::MSE_BLOCKS + 4) * sizeof(float); }

extern "C" int vs_mse_loss(const float* pred, const float* target, int64_t n, float weight,
                           float* loss_out, float* grad_out, void* workspace, vs_stream_t stream) {
  using namespace vs;
  VS_REQUIRE(pred && target && loss_out && workspace, "mse_loss: null tensor");
  VS_REQUIRE(n > 0, "mse_loss: empty input");
  float* partial = static_cast<float*>(workspace);
  unsigned int* counter = reinterpret_cast<unsigned int*>(partial + MSE_BLOCKS);
  const long long want = (n / 4 + MSE_THREADS - 1) / MSE_THREADS;
  const unsigned blocks = static_cast<unsigned>(want < 1 ? 1 : (want > MSE_BLOCKS ? MSE_BLOCKS : want));
  mse_loss_kernel<<<blocks, MSE_THREADS, 0, to_stream(stream)>>>(pred, target, n, weight, loss_out,
                                                               grad_out, partial, counter);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_update_pose(const float* rho, const float* theta, const float* c2w, float* c2w_out,
                              int n, vs_stream_t stream) {
  using namespace vs;
  if (n <= 0) return VS_OK;
  VS_REQUIRE(rho && theta && c2w && c2w_out, "update_pose: null tensor");
  update_pose_kernel<<<blocks_for(n, 64), 64, 0, to_stream(stream)>>>(rho, theta, c2w, c2w_out, n);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

extern "C" int vs_adamw_step(const vs_adamw_params* p, vs_stream_t stream) {
  using namespace vs;
  VS_REQUIRE(p != nullptr, "adamw_step: null params");
  if (p->n_tensors <= 0 || p->n_chunks <= 0) return VS_OK;
  VS_REQUIRE(p->params && p->grads && p->exp_avg && p->exp_avg_sq && p->sizes && p->lrs &&
                 p->chunk_tensor && p->chunk_start,
             "adamw_step: null table");
  VS_REQUIRE(p->partials && p->counter && p->grad_norm_out && p->found_inf_out,
             "adamw_step: null scratch / output");
  VS_REQUIRE((p->step >= 1 || p->step_counter != nullptr) && p->skip_nonfinite >= 0 && p->skip_nonfinite <= 2 &&
                 p->beta1 >= 0.f && p->beta1 < 1.f && p->beta2 >= 0.f && p->beta2 < 1.f,
             "adamw_step: bad step / betas");
  cudaStream_t st = to_stream(stream);
  VS_CUDA(cudaMemsetAsync(p->found_inf_out, 0, sizeof(int32_t), st));
  adamw_norm_kernel<<<p->n_chunks, AW_THREADS, 0, st>>>(*p);
  VS_LAUNCH_CHECK();
  const float bc1 = 1.0f - powf(p->beta1, static_cast<float>(p->step));
  const float bc2 = 1.0f - powf(p->beta2, static_cast<float>(p->step));
  adamw_update_kernel<<<p->n_chunks, AW_THREADS, 0, st>>>(*p, bc1, sqrtf(bc2));
  VS_LAUNCH_CHECK();
  return VS_OK;
}
