// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/vicasplat_b200.h"

namespace vs {

void set_error(const char* fmt, ...);

#define VS_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      ::vs::set_error(__VA_ARGS__);  \
      return VS_ERR_INVALID;         \
    }                                \
  } while (0)

#define VS_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      ::vs::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                      __LINE__);                                                        \
      return VS_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

// every kernel launch of this library goes through VS_LAUNCH_CHECK: it also feeds vs_launch_count()
void count_launch();
#define VS_LAUNCH_CHECK()          \
  do {                             \
    ::vs::count_launch();          \
    VS_CUDA(cudaGetLastError());   \
  } while (0)

inline cudaStream_t to_stream(vs_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace vs
