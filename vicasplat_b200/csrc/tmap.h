// Host-side TMA tensor-map construction (bf16, 128-byte swizzle), shared by the GEMM and attention
// translation units.  The driver entry point is resolved through the runtime so the library does
// not link against libcuda.
#pragma once
#include <cuda.h>

#include "common.h"

namespace vs {
// dims[0] is the innermost (contiguous) dimension; strides_bytes has rank-1 entries (dims 1..).
int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
               const cuuint64_t* strides_bytes, const cuuint32_t* box);
}  // namespace vs
