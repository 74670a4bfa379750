// The 16-bit operand format of the encoder path is a RUN-TIME choice between
//   bf16  (speed mode: 8 exponent / 7 mantissa bits) and
//   fp16  (parity mode: 5 exponent / 10 mantissa bits -- the mantissa of TF32, which is what the reference's
//          matmuls round their fp32 operands to, backbone_vica.py:9; tcgen05 kind::f16 runs both at the
//          same rate, i.e. twice kind::tf32's).
// Both are 2-byte types: tensor maps, shared-memory tiles, strides and descriptors are identical, only the
// conversions at the producers / consumers and the MMA instruction descriptor's format bits differ.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>

namespace vs {

__device__ __forceinline__ float2 h2_to_f2(uint32_t w, bool f16) {
  return f16 ? __half22float2(*reinterpret_cast<const __half2*>(&w))
             : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
__device__ __forceinline__ uint32_t f2_to_h2(float a, float b, bool f16) {
  if (f16) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float h_to_f(uint16_t w, bool f16) {
  return f16 ? __half2float(*reinterpret_cast<const __half*>(&w))
             : __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&w));
}
__device__ __forceinline__ uint16_t f_to_h(float a, bool f16) {
  if (f16) {
    const __half h = __float2half_rn(a);
    return *reinterpret_cast<const uint16_t*>(&h);
  }
  const __nv_bfloat16 h = __float2bfloat16(a);
  return *reinterpret_cast<const uint16_t*>(&h);
}
// round-trip through the 16-bit format (what a stand-alone kernel's store + the consumer's load do)
__device__ __forceinline__ float round_h(float a, bool f16) { return h_to_f(f_to_h(a, f16), f16); }

}  // namespace vs
