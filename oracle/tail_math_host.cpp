// TEST INFRASTRUCTURE: host build of vicasplat_b200/csrc/tail_math.h (the per-element chain rules the
// backward kernels of the encoder tails will run) so that their arithmetic can be checked against the
// fp64 oracle without a GPU.  Built by oracle/Makefile into oracle/_build/libtail_math_host.so.
#include "../vicasplat_b200/csrc/tail_math.h"

extern "C" {
void tm_adapter_backward(const float* raw, long long ld, const float* d_means, const float* d_cov6,
                         const float* d_opac, float* d_raw, long long n) {
  for (long long i = 0; i < n; ++i)
    vs::adapter_backward_one(raw + i * ld, d_means + 3 * i, d_cov6 + 6 * i, d_opac[i], d_raw + 11 * i);
}
void tm_exp_postprocess_backward(const float* x, const float* g, float* dx, long long n) {
  for (long long i = 0; i < n; ++i) vs::exp_postprocess_backward_one(x + 3 * i, g + 3 * i, dx + 3 * i);
}
void tm_dq_normalise_backward(const float* v, const float* dp, float* dv, long long n) {
  for (long long i = 0; i < n; ++i) vs::dq_normalise_backward_one(v + 8 * i, dp + 8 * i, dv + 8 * i);
}
}
