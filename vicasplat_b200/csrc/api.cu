// Error reporting shared by all C-ABI entry points.
#include <cstring>

#include "common.h"
#include "tmap.h"

#include <atomic>

namespace vs {
namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace vs

namespace vs {
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
}  // namespace

int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
               const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
    return VS_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                  const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu %llu)",
              static_cast<int>(r), rank, (unsigned long long)dims[0],
              (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)(rank > 3 ? dims[3] : 0));
    return VS_ERR_CUDA;
  }
  return VS_OK;
}
}  // namespace vs

extern "C" const char* vs_last_error(void) { return vs::g_err; }
extern "C" int vs_version(void) { return 100; }
extern "C" int64_t vs_launch_count(void) { return vs::g_launches.load(std::memory_order_relaxed); }

extern "C" int64_t vs_struct_size(const char* name) {
  if (name == nullptr) return -1;
#define VS_SZ(T) if (std::strcmp(name, #T) == 0) return static_cast<int64_t>(sizeof(T));
  VS_SZ(vs_gemm_params)
  VS_SZ(vs_layernorm_params)
  VS_SZ(vs_attention_params)
  VS_SZ(vs_raster_params)
  VS_SZ(vs_raster_bwd_params)
  VS_SZ(vs_adamw_params)
  VS_SZ(vs_layernorm_bwd_params)
  VS_SZ(vs_attention_bwd_params)
  VS_SZ(vs_ln_mod_bwd_params)
#undef VS_SZ
  return -1;
}
