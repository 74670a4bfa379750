// Per-element chain rules of the encoder plugin's tails, shared by device code and by a host build
// (oracle/tail_math_host.cpp) that tests/test_oracle_decoder_backward_cpu.py checks against the
// hand-derived fp64 oracle (oracle/adapter_backward_ref.py) -- so the arithmetic of the backward
// kernels of the next rows is verified without a GPU:
//   * MyGaussianAdapter.forward backward (common/gaussian_adapter.py:167-212, common/gaussians.py:8-44):
//     from d means, d cov6 (the packed upper triangle vs_raster_backward writes) and d opacity to the
//     gradient of the 11 leading raw channels (xyz | opacity | scale 3 | quaternion xyzw 4)
//   * 'exp' depth postprocess backward (heads/postprocess.py:42-61)
//   * dual-quaternion normalisation backward (vicasplat.py:183-190)
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define VS_HD __host__ __device__ __forceinline__
#else
#define VS_HD inline
#endif

namespace vs {

// raw = (x, y, z | o | s0, s1, s2 | qx, qy, qz, qw); G = dL/d covariance as ANY 3x3 matrix (a packed
// upper-triangle gradient is the upper-triangular G, a full-matrix gradient is G itself): only
// G + G^T enters.
VS_HD void adapter_backward_general(const float* raw, const float* d_means, const float* G,
                                    float d_opac, float* d_raw) {
  d_raw[0] = d_means[0]; d_raw[1] = d_means[1]; d_raw[2] = d_means[2];
  // opacity = sigmoid(o)
  const float op = 1.0f / (1.0f + expf(-raw[3]));
  d_raw[3] = d_opac * op * (1.0f - op);
  // scales = min(0.001 * softplus(s), 0.3)
  float sc[3], dsc_ds[3];
  for (int a = 0; a < 3; ++a) {
    const float s = raw[4 + a];
    const float sp = s > 20.0f ? s : log1pf(expf(s));          // F.softplus (threshold 20)
    const float un = 0.001f * sp;
    sc[a] = fminf(un, 0.3f);
    dsc_ds[a] = un < 0.3f ? 0.001f / (1.0f + expf(-s)) : 0.0f;
  }
  // rotation = r / max(|r|, 1e-12);  R = I + s_q A(q), s_q = 2 / (q.q + 1e-8)
  const float n = fmaxf(sqrtf(raw[7] * raw[7] + raw[8] * raw[8] + raw[9] * raw[9] + raw[10] * raw[10]), 1e-12f);
  const float i = raw[7] / n, j = raw[8] / n, k = raw[9] / n, r = raw[10] / n;
  const float sq = 2.0f / (i * i + j * j + k * k + r * r + 1e-8f);
  const float A[9] = {-(j * j + k * k), i * j - k * r, i * k + j * r,
                      i * j + k * r, -(i * i + k * k), j * k - i * r,
                      i * k - j * r, j * k + i * r, -(i * i + j * j)};
  float R[9];
  for (int a = 0; a < 9; ++a) R[a] = sq * A[a] + ((a == 0 || a == 4 || a == 8) ? 1.0f : 0.0f);
  // covariance = R diag(sc^2) R^T
  float GR[9], GtR[9];   // G R and G^T R
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      float u = 0.f, v = 0.f;
      for (int c = 0; c < 3; ++c) {
        u += G[a * 3 + c] * R[c * 3 + b];
        v += G[c * 3 + a] * R[c * 3 + b];
      }
      GR[a * 3 + b] = u;
      GtR[a * 3 + b] = v;
    }
  float dR[9];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) dR[a * 3 + b] = (GR[a * 3 + b] + GtR[a * 3 + b]) * sc[b] * sc[b];
  for (int b = 0; b < 3; ++b) {
    const float diag = R[0 * 3 + b] * GR[0 * 3 + b] + R[1 * 3 + b] * GR[1 * 3 + b] + R[2 * 3 + b] * GR[2 * 3 + b];
    d_raw[4 + b] = 2.0f * sc[b] * diag * dsc_ds[b];
  }
  // dL/dq through A (scaled by s_q) and through s_q = 2 / u (ds_q/dq = -s_q^2 q)
  const float g00 = dR[0], g01 = dR[1], g02 = dR[2], g10 = dR[3], g11 = dR[4], g12 = dR[5], g20 = dR[6],
              g21 = dR[7], g22 = dR[8];
  float dq[4];
  dq[0] = -2.f * i * (g11 + g22) + j * (g01 + g10) + k * (g02 + g20) + r * (g21 - g12);
  dq[1] = -2.f * j * (g00 + g22) + i * (g01 + g10) + k * (g12 + g21) + r * (g02 - g20);
  dq[2] = -2.f * k * (g00 + g11) + i * (g02 + g20) + j * (g12 + g21) + r * (g10 - g01);
  dq[3] = k * (g10 - g01) + j * (g02 - g20) + i * (g21 - g12);
  float dsq = 0.f;
  for (int a = 0; a < 9; ++a) dsq += dR[a] * A[a];
  const float q[4] = {i, j, k, r};
  float dot = 0.f;
  for (int a = 0; a < 4; ++a) {
    dq[a] = sq * dq[a] - dsq * sq * sq * q[a];
    dot += q[a] * dq[a];
  }
  // through the normalisation r / |r|
  for (int a = 0; a < 4; ++a) d_raw[7 + a] = (dq[a] - q[a] * dot) / n;
}

// d_cov6 = dL/d(xx, xy, xz, yy, yz, zz): the packed upper triangle vs_raster_backward writes
VS_HD void adapter_backward_one(const float* raw, const float* d_means, const float* d_cov6,
                                float d_opac, float* d_raw) {
  const float G[9] = {d_cov6[0], d_cov6[1], d_cov6[2], 0.f, d_cov6[3], d_cov6[4], 0.f, 0.f, d_cov6[5]};
  adapter_backward_general(raw, d_means, G, d_opac, d_raw);
}

// xyz = x / max(|x|, 1e-8) * expm1(|x|)
VS_HD void exp_postprocess_backward_one(const float* x, const float* g, float* dx) {
  const float d = fmaxf(sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]), 1e-8f);
  const float em = expm1f(d);
  const float f = em / d;
  const float fp = (expf(d) * d - em) / (d * d);
  const float xg = (x[0] * g[0] + x[1] * g[1] + x[2] * g[2]) * fp / d;
  for (int a = 0; a < 3; ++a) dx[a] = f * g[a] + x[a] * xg;
}

// pred = v / |v[0..3]|  (all 8 components)
VS_HD void dq_normalise_backward_one(const float* v, const float* dp, float* dv) {
  const float n = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
  float dot = 0.f;
  for (int a = 0; a < 8; ++a) dot += dp[a] * v[a] / n;
  for (int a = 0; a < 8; ++a) dv[a] = (dp[a] - (a < 4 ? dot * v[a] / n : 0.f)) / n;
}

}  // namespace vs
