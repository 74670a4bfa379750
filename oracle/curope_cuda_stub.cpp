// Link-time stand-in for the reference's CUDA branch (curope/kernels.cu), which is not compiled for
// the CPU-only oracle build: oracle/_ref only ever runs the reference's rope_2d_cpu.
#include <torch/extension.h>
void rope_2d_cuda(torch::Tensor, const torch::Tensor, const float, const float) {
  TORCH_CHECK(false, "oracle/_ref is the CPU build of the reference curope extension");
}
