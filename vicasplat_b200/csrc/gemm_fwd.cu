// gemm_tc05_kernel, variant 0: forward path, bf16 operands (speed mode).  See gemm_kernel.cuh.
#define VS_GEMM_VARIANT 0
#include "gemm_kernel.cuh"
