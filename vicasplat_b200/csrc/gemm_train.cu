// gemm_tc05_kernel, variant 2: training path: bf16, MN-major / conv-wgrad operands, split-K, atomic and ReLU-mask epilogues.  See gemm_kernel.cuh.
#define VS_GEMM_VARIANT 2
#include "gemm_kernel.cuh"
