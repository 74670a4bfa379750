// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
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

// cudaFuncSetAttribute is PER DEVICE: one-time kernel configuration is keyed by the current device (a
// bit per device, set after the attributes are in place, so a racing second thread at worst repeats
// the idempotent calls).
#define VS_CONFIGURE_PER_DEVICE(...)                                                   \
  do {                                                                                 \
    static std::atomic<unsigned long long> vs_done_{0ull};                             \
    int vs_dev_ = 0;                                                                   \
    VS_CUDA(cudaGetDevice(&vs_dev_));                                                  \
    const unsigned long long vs_bit_ = 1ull << (vs_dev_ & 63);                         \
    if (!(vs_done_.load(std::memory_order_acquire) & vs_bit_)) {                       \
      __VA_ARGS__                                                                      \
      vs_done_.fetch_or(vs_bit_, std::memory_order_release);                           \
    }                                                                                  \
  } while (0)

inline cudaStream_t to_stream(vs_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace vs
