/* CPU oracle for the 2-D rotary embedding.
 *
 * TEST INFRASTRUCTURE ONLY: loaded by tests/ (ctypes) as the checker; the product never links it.
 *
 * Plain-C restatement of the reference's CPU path rope_2d_cpu
 * (src/model/encoder/backbone/croco/curope/curope.cpp:11-47), which the CUDA kernel
 * (curope/kernels.cu:18-82) and the PyTorch fallback (croco/pos_embed.py:112-159) both mirror:
 * tokens (B, N, H, D) fp32 in place, positions (B, N, 2) int64 (y, x), Q = D/4.  For axis
 * X in {y, x} and d in [0, Q): the pair (tok[X*2Q + d], tok[X*2Q + Q + d]) is rotated by the angle
 * fwd * pos[X] / base^(d/Q).
 *
 * Parity pinned: tests/golden/rope_2d.npz holds outputs of the reference's own curope.cpp compiled
 * here (oracle/_ref, recipe in oracle/Makefile); tests/test_oracle_rope_cpu.py checks this file
 * against them.
 */
#include <math.h>
#include <stdint.h>

void rope_2d_ref(float* tok, const int64_t* pos, int B, int N, int H, int D, float base, float fwd) {
  const int Q = D / 4;
  for (int b = 0; b < B; ++b)
    for (int n = 0; n < N; ++n)
      for (int X = 0; X < 2; ++X) {
        const int64_t p = pos[((int64_t)b * N + n) * 2 + X];
        for (int d = 0; d < Q; ++d) {
          const float ang = fwd * (float)p / powf(base, (float)d / (float)Q);
          const float c = cosf(ang), s = sinf(ang);
          for (int h = 0; h < H; ++h) {
            float* t = tok + (((int64_t)b * N + n) * H + h) * D + X * 2 * Q;
            const float u = t[d], v = t[d + Q];
            t[d] = u * c - v * s;
            t[d + Q] = v * c + u * s;
          }
        }
      }
}
