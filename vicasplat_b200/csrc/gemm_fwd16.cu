// gemm_tc05_kernel, variant 1: forward path, fp16 operands (parity mode).  See gemm_kernel.cuh.
#define VS_GEMM_VARIANT 1
#include "gemm_kernel.cuh"
