// Tile-based differentiable 3-D Gaussian splatting rasterizer for sm_100a (forward).
//
// Semantics follow the diff_gaussian_rasterization extension the reference calls at
// src/model/decoder/cuda_splatting.py:207-235 (EWA projection, 16x16 tiles, depth-sorted
// front-to-back alpha compositing; constants in SURVEY.md Appendix D).  The organisation is new:
//   * all V views of a scene are rendered by ONE launch chain: the Gaussians are read once
//     (coalesced, SH staged through shared memory), each thread projects its Gaussian into every
//     view, and the (view, tile, depth) keys of all views go through a single radix sort;
//   * per-view intermediates are SoA so the blend kernel's gathers are 16-byte vector loads;
//   * the blend kernel stages 256-splat batches of a tile's list in shared memory, each warp
//     ballots its done-mask for early exit.
#include <cub/block/block_radix_sort.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.h"

namespace vs {
namespace {

constexpr int TILE = 16;
constexpr int BLEND_THREADS = TILE * TILE;

// upstream constants (named so the oracle and the kernel agree)
constexpr float kNearCullZ = 0.2f;
constexpr float kLowpass = 0.3f;
constexpr float kFovClamp = 1.3f;
constexpr float kLambdaFloor = 0.1f;
constexpr float kAlphaMax = 0.99f;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kTStop = 1e-4f;
constexpr float kWEps = 1e-7f;
constexpr float kNTouchedT = 0.5f;

constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
__constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
__constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f,  -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

struct Workspace {
  float2* xy;        // (V*G)
  float4* conic_o;   // (V*G)  conic a,b,c + opacity
  float4* rgbd;      // (V*G)  rgb + view-space depth
  uint32_t* tiles;   // (V*G)  tiles touched
  uint32_t* offsets; // (V*G)  inclusive scan of tiles
  int32_t* radii;    // (V*G)  (only used when the caller passes radii == NULL)
  uint64_t* keys_in;
  uint64_t* keys_out;
  uint32_t* vals_in;
  uint32_t* vals_out;
  uint2* ranges;     // (V*tiles)
  uint32_t* tile_counts;  // (V*tiles)  per-tile binning path
  uint32_t* tile_cursor;  // (V*tiles)
  void* cub_tmp;
  size_t cub_bytes;
  size_t total;
};

inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

Workspace carve(void* base, int V, int G, int H, int W, int64_t max_pairs) {
  Workspace w{};
  const size_t vg = static_cast<size_t>(V) * G;
  const size_t nt = static_cast<size_t>(V) * ceil_div(H, TILE) * ceil_div(W, TILE);
  size_t scan_bytes = 0, sort_bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, static_cast<uint32_t*>(nullptr),
                                static_cast<uint32_t*>(nullptr), static_cast<int>(vg));
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, static_cast<uint64_t*>(nullptr),
                                  static_cast<uint64_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                  static_cast<uint32_t*>(nullptr), static_cast<int>(max_pairs), 0,
                                  64);
  w.cub_bytes = scan_bytes > sort_bytes ? scan_bytes : sort_bytes;
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = p ? p + off : nullptr;
    off += align256(bytes);
    return r;
  };
  w.xy = static_cast<float2*>(take(vg * sizeof(float2)));
  w.conic_o = static_cast<float4*>(take(vg * sizeof(float4)));
  w.rgbd = static_cast<float4*>(take(vg * sizeof(float4)));
  w.tiles = static_cast<uint32_t*>(take(vg * 4));
  w.offsets = static_cast<uint32_t*>(take(vg * 4));
  w.radii = static_cast<int32_t*>(take(vg * 4));
  w.keys_in = static_cast<uint64_t*>(take(static_cast<size_t>(max_pairs) * 8));
  w.keys_out = static_cast<uint64_t*>(take(static_cast<size_t>(max_pairs) * 8));
  w.vals_in = static_cast<uint32_t*>(take(static_cast<size_t>(max_pairs) * 4));
  w.vals_out = static_cast<uint32_t*>(take(static_cast<size_t>(max_pairs) * 4));
  w.ranges = static_cast<uint2*>(take(nt * sizeof(uint2)));
  w.tile_counts = static_cast<uint32_t*>(take(nt * 4));
  w.tile_cursor = static_cast<uint32_t*>(take(nt * 4));
  w.cub_tmp = take(w.cub_bytes);
  w.total = off;
  return w;
}

struct View {
  float vm[16];  // column-major view matrix (as passed by the reference, transposed)
  float pm[16];  // column-major full projection
  float cam[3];
  float tanx, tany;
};

__device__ __forceinline__ void tile_rect(float px, float py, int radius, int gx, int gy,
                                          int& x0, int& x1, int& y0, int& y1) {
  x0 = min(gx, max(0, static_cast<int>((px - radius) / TILE)));
  y0 = min(gy, max(0, static_cast<int>((py - radius) / TILE)));
  x1 = min(gx, max(0, static_cast<int>((px + radius + TILE - 1) / TILE)));
  y1 = min(gy, max(0, static_cast<int>((py + radius + TILE - 1) / TILE)));
}

// ------------------------------------------------------------------ preprocess
// One thread per Gaussian; loops over the views [v_begin, v_end) (all V when the set is shared,
// exactly one when every view has its own set).  SH coefficients of the warp's 32 Gaussians are
// staged through shared memory with coalesced 16-byte loads.
__global__ void __launch_bounds__(128)
    preprocess_kernel(int G, int V, int shared_set, int H, int W, const float* __restrict__ means,
                      const float* __restrict__ cov6, const float* __restrict__ opac,
                      const float* __restrict__ shs, int M, int sh_cs, int sh_ch, int degree,
                      const float* __restrict__ colors, const float* __restrict__ viewm,
                      const float* __restrict__ projm, const float* __restrict__ campos,
                      const float* __restrict__ tanfov, float2* __restrict__ xy,
                      float4* __restrict__ conic_o, float4* __restrict__ rgbd,
                      uint32_t* __restrict__ tiles, int32_t* __restrict__ radii) {
  extern __shared__ float sh_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int v_begin = shared_set ? 0 : blockIdx.y;
  const int v_end = shared_set ? V : blockIdx.y + 1;
  const size_t set_off = shared_set ? 0 : static_cast<size_t>(blockIdx.y) * G;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;

  // ---- stage SH for this warp
  const int n_sh = (degree >= 3 ? 16 : (degree + 1) * (degree + 1));
  float* my_sh = nullptr;
  if (shs != nullptr) {
    const int per = 3 * M;
    float* ws = sh_smem + static_cast<size_t>(warp) * 32 * per;
    const int gbase = blockIdx.x * blockDim.x + warp * 32;
    const int cnt = max(0, min(32, G - gbase)) * per;
    const float* src = shs + (set_off + gbase) * per;
    if ((cnt & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(ws);
      for (int i = lane; i < cnt / 4; i += 32) d4[i] = __ldg(s4 + i);
    } else {
      for (int i = lane; i < cnt; i += 32) ws[i] = __ldg(src + i);
    }
    __syncwarp();
    my_sh = ws + lane * per;
  }
  if (g >= G) return;

  const size_t gi = set_off + g;
  const float mx = means[gi * 3 + 0], my = means[gi * 3 + 1], mz = means[gi * 3 + 2];
  float c6[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) c6[i] = cov6[gi * 6 + i];
  const float op = opac[gi];

  for (int v = v_begin; v < v_end; ++v) {
    const float* vm = viewm + v * 16;
    const float* pm = projm + v * 16;
    const size_t o = static_cast<size_t>(v) * G + g;
    // view-space point
    const float tx = vm[0] * mx + vm[4] * my + vm[8] * mz + vm[12];
    const float ty = vm[1] * mx + vm[5] * my + vm[9] * mz + vm[13];
    const float tz = vm[2] * mx + vm[6] * my + vm[10] * mz + vm[14];
    int radius = 0;
    uint32_t ntiles = 0;
    if (tz > kNearCullZ) {
      const float hx = pm[0] * mx + pm[4] * my + pm[8] * mz + pm[12];
      const float hy = pm[1] * mx + pm[5] * my + pm[9] * mz + pm[13];
      const float hw = pm[3] * mx + pm[7] * my + pm[11] * mz + pm[15];
      const float pw = 1.0f / (hw + kWEps);
      const float px = ((hx * pw + 1.0f) * W - 1.0f) * 0.5f;
      const float py = ((hy * pw + 1.0f) * H - 1.0f) * 0.5f;
      // EWA 2-D covariance
      const float tanx = tanfov[v * 2 + 0], tany = tanfov[v * 2 + 1];
      const float fx = W / (2.0f * tanx), fy = H / (2.0f * tany);
      const float limx = kFovClamp * tanx, limy = kFovClamp * tany;
      const float txc = fminf(limx, fmaxf(-limx, tx / tz)) * tz;
      const float tyc = fminf(limy, fmaxf(-limy, ty / tz)) * tz;
      const float j00 = fx / tz, j02 = -fx * txc / (tz * tz);
      const float j11 = fy / tz, j12 = -fy * tyc / (tz * tz);
      // T = J * Wrot, Wrot[r][c] = vm[c*4 + r]
      const float t00 = j00 * vm[0] + j02 * vm[2];
      const float t01 = j00 * vm[4] + j02 * vm[6];
      const float t02 = j00 * vm[8] + j02 * vm[10];
      const float t10 = j11 * vm[1] + j12 * vm[2];
      const float t11 = j11 * vm[5] + j12 * vm[6];
      const float t12 = j11 * vm[9] + j12 * vm[10];
      // S * T^T rows
      const float s0x = c6[0] * t00 + c6[1] * t01 + c6[2] * t02;
      const float s0y = c6[1] * t00 + c6[3] * t01 + c6[4] * t02;
      const float s0z = c6[2] * t00 + c6[4] * t01 + c6[5] * t02;
      const float s1x = c6[0] * t10 + c6[1] * t11 + c6[2] * t12;
      const float s1y = c6[1] * t10 + c6[3] * t11 + c6[4] * t12;
      const float s1z = c6[2] * t10 + c6[4] * t11 + c6[5] * t12;
      const float a = t00 * s0x + t01 * s0y + t02 * s0z + kLowpass;
      const float b = t00 * s1x + t01 * s1y + t02 * s1z;
      const float c = t10 * s1x + t11 * s1y + t12 * s1z + kLowpass;
      const float det = a * c - b * b;
      if (det != 0.0f) {
        const float inv = 1.0f / det;
        const float mid = 0.5f * (a + c);
        const float lam = mid + sqrtf(fmaxf(kLambdaFloor, mid * mid - det));
        const int rad = static_cast<int>(ceilf(3.0f * sqrtf(lam)));
        int x0, x1, y0, y1;
        tile_rect(px, py, rad, gx, gy, x0, x1, y0, y1);
        const uint32_t nt = static_cast<uint32_t>((x1 - x0) * (y1 - y0));
        if (nt > 0) {
          radius = rad;
          ntiles = nt;
          float r, gcol, bcol;
          if (colors != nullptr) {
            r = colors[gi * 3 + 0];
            gcol = colors[gi * 3 + 1];
            bcol = colors[gi * 3 + 2];
          } else {
            float dx = mx - campos[v * 3 + 0], dy = my - campos[v * 3 + 1],
                  dz = mz - campos[v * 3 + 2];
            const float il = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
            dx *= il; dy *= il; dz *= il;
            float basis[16];
            basis[0] = SH_C0;
            if (n_sh > 1) {
              basis[1] = -SH_C1 * dy; basis[2] = SH_C1 * dz; basis[3] = -SH_C1 * dx;
            }
            if (n_sh > 4) {
              const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
              const float xy_ = dx * dy, yz = dy * dz, xz = dx * dz;
              basis[4] = SH_C2[0] * xy_;
              basis[5] = SH_C2[1] * yz;
              basis[6] = SH_C2[2] * (2.0f * zz - xx - yy);
              basis[7] = SH_C2[3] * xz;
              basis[8] = SH_C2[4] * (xx - yy);
              if (n_sh > 9) {
                basis[9] = SH_C3[0] * dy * (3.0f * xx - yy);
                basis[10] = SH_C3[1] * xy_ * dz;
                basis[11] = SH_C3[2] * dy * (4.0f * zz - xx - yy);
                basis[12] = SH_C3[3] * dz * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                basis[13] = SH_C3[4] * dx * (4.0f * zz - xx - yy);
                basis[14] = SH_C3[5] * dz * (xx - yy);
                basis[15] = SH_C3[6] * dx * (xx - 3.0f * yy);
              }
            }
            r = 0.f; gcol = 0.f; bcol = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              if (k < n_sh) {
                r += basis[k] * my_sh[k * sh_cs + 0 * sh_ch];
                gcol += basis[k] * my_sh[k * sh_cs + 1 * sh_ch];
                bcol += basis[k] * my_sh[k * sh_cs + 2 * sh_ch];
              }
            }
            r = fmaxf(r + 0.5f, 0.0f);
            gcol = fmaxf(gcol + 0.5f, 0.0f);
            bcol = fmaxf(bcol + 0.5f, 0.0f);
          }
          xy[o] = make_float2(px, py);
          conic_o[o] = make_float4(c * inv, -b * inv, a * inv, op);
          rgbd[o] = make_float4(r, gcol, bcol, tz);
        }
      }
    }
    tiles[o] = ntiles;
    radii[o] = radius;
  }
}

// ------------------------------------------------------------------ binning
__global__ void emit_pairs_kernel(size_t VG, int G, int H, int W, const float2* __restrict__ xy,
                                  const float4* __restrict__ rgbd, const int32_t* __restrict__ radii,
                                  const uint32_t* __restrict__ offsets, uint64_t* __restrict__ keys,
                                  uint32_t* __restrict__ vals, int64_t max_pairs) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= VG) return;
  const int rad = radii[i];
  if (rad <= 0) return;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const uint32_t v = static_cast<uint32_t>(i / G);
  uint32_t off = (i == 0) ? 0u : offsets[i - 1];
  const float2 p = xy[i];
  int x0, x1, y0, y1;
  tile_rect(p.x, p.y, rad, gx, gy, x0, x1, y0, y1);
  const uint32_t dbits = __float_as_uint(rgbd[i].w);
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      if (static_cast<int64_t>(off) < max_pairs) {
        const uint64_t tile = static_cast<uint64_t>(v) * (gx * gy) + y * gx + x;
        keys[off] = (tile << 32) | dbits;
        vals[off] = static_cast<uint32_t>(i);
      }
      ++off;
    }
}

__global__ void pad_keys_kernel(const uint32_t* __restrict__ offsets, size_t VG,
                                uint64_t* __restrict__ keys, int64_t max_pairs,
                                int64_t* __restrict__ num_pairs_out) {
  const int64_t n = offsets[VG - 1];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i == 0 && num_pairs_out != nullptr) *num_pairs_out = n;
  if (i >= n && i < max_pairs) keys[i] = ~0ull;
}

__global__ void tile_ranges_kernel(const uint64_t* __restrict__ keys, int64_t max_pairs,
                                   const uint32_t* __restrict__ offsets, size_t VG,
                                   uint2* __restrict__ ranges) {
  const int64_t n = min(static_cast<int64_t>(offsets[VG - 1]), max_pairs);
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t t = static_cast<uint32_t>(keys[i] >> 32);
  if (i == 0) ranges[t].x = 0;
  else {
    const uint32_t tp = static_cast<uint32_t>(keys[i - 1] >> 32);
    if (tp != t) {
      ranges[tp].y = static_cast<uint32_t>(i);
      ranges[t].x = static_cast<uint32_t>(i);
    }
  }
  if (i == n - 1) ranges[t].y = static_cast<uint32_t>(n);
}

// Half extents of the axis-aligned box around the region where a splat's alpha can reach 1/255:
// alpha >= 1/255  <=>  power >= -tau, tau = ln(255 o): an ellipse with half extents
// sqrt(2 tau C / det), sqrt(2 tau A / det) (conic = (A, B, C)); 1 % + 0.01 px safety margin.
// Negative extents = never visible.  Used by the blend's sub-block culling AND by the per-tile
// binning (a (splat, tile) pair whose box misses the tile's pixel centres cannot contribute to any
// pixel of it, so dropping it leaves the image bit-identical and shortens sort and blend lists).
__device__ __forceinline__ float2 alpha_extent(const float4 co) {
  const float tau = __logf(255.0f * co.w) * 1.01f + 1e-3f;
  const float det = co.x * co.z - co.y * co.y;
  float ex = -1.f, ey = -1.f;
  if (tau > 0.f) {
    if (det > 0.f) {
      ex = sqrtf(2.0f * tau * co.z / det) + 0.01f;
      ey = sqrtf(2.0f * tau * co.x / det) + 0.01f;
    } else {
      ex = ey = 1e30f;  // degenerate conic: never cull
    }
  }
  return make_float2(ex, ey);
}

__device__ __forceinline__ bool tile_may_hit(float cx, float cy, float2 e, int tx, int ty, int W, int H) {
  const float bx0 = static_cast<float>(tx * TILE), bx1 = static_cast<float>(min(tx * TILE + TILE - 1, W - 1));
  const float by0 = static_cast<float>(ty * TILE), by1 = static_cast<float>(min(ty * TILE + TILE - 1, H - 1));
  const float ddx = fmaxf(fmaxf(bx0 - cx, cx - bx1), 0.f);
  const float ddy = fmaxf(fmaxf(by0 - cy, cy - by1), 0.f);
  return ddx <= e.x && ddy <= e.y;
}

// ------------------------------------------------------------------ per-tile binning + sort
// Alternative to the global 64-bit radix sort (6 passes over every pair): pairs are counted per
// tile, scattered into their tile's segment, and each segment is sorted by ONE CTA in shared
// memory (cub::BlockRadixSort on (depth bits, Gaussian id): the id makes the order independent of
// the scatter order and reproduces the stable global sort exactly).  HBM traffic drops from
// ~12 x 12 B to ~3 x 8 B per pair.  Used when the caller provides max_tile_pairs <= 16384.
// Pixel-aligned Gaussians of one block land in a handful of tiles: count in shared memory first,
// then ONE global atomic per touched tile per block (13 M same-address global atomics otherwise).
// grid = (blocks over g, V); dynamic smem = tiles_per_view counters (0 -> direct global atomics).
__global__ void tile_count_kernel(int G, int H, int W, const float2* __restrict__ xy,
                                  const float4* __restrict__ conic_o,
                                  const int32_t* __restrict__ radii,
                                  uint32_t* __restrict__ tile_counts, int use_smem) {
  extern __shared__ uint32_t s_cnt[];
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int tpv = gx * gy;
  const uint32_t v = blockIdx.y;
  if (use_smem) {
    for (int i = threadIdx.x; i < tpv; i += blockDim.x) s_cnt[i] = 0u;
    __syncthreads();
  }
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G) {
    const size_t i = static_cast<size_t>(v) * G + g;
    const int rad = radii[i];
    if (rad > 0) {
      const float2 p = xy[i];
      int x0, x1, y0, y1;
      tile_rect(p.x, p.y, rad, gx, gy, x0, x1, y0, y1);
      const float2 ext = alpha_extent(conic_o[i]);
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
          if (!tile_may_hit(p.x, p.y, ext, x, y, W, H)) continue;
          if (use_smem) atomicAdd(s_cnt + y * gx + x, 1u);
          else atomicAdd(tile_counts + v * tpv + y * gx + x, 1u);
        }
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < tpv; i += blockDim.x)
      if (s_cnt[i] != 0u) atomicAdd(tile_counts + v * tpv + i, s_cnt[i]);
  }
}

// single CTA: exclusive scan of the tile counts -> ranges (clipped to the capacity), totals
__global__ void __launch_bounds__(1024)
    tile_offsets_kernel(const uint32_t* __restrict__ counts, int n_tiles, int64_t max_pairs,
                        uint2* __restrict__ ranges, uint32_t* __restrict__ cursor,
                        int64_t* __restrict__ num_pairs_out) {
  __shared__ unsigned long long s_part[1024];
  __shared__ unsigned s_max[1024];
  const int t = threadIdx.x;
  const int per = (n_tiles + 1023) / 1024;
  const int b = t * per, e = min(n_tiles, b + per);
  unsigned long long sum = 0;
  unsigned mx = 0;
  for (int i = b; i < e; ++i) { sum += counts[i]; mx = max(mx, counts[i]); }
  s_part[t] = sum;
  s_max[t] = mx;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {   // Hillis-Steele inclusive scan / max
    unsigned long long v = t >= off ? s_part[t - off] : 0ull;
    unsigned m = t >= off ? s_max[t - off] : 0u;
    __syncthreads();
    s_part[t] += v;
    s_max[t] = max(s_max[t], m);
    __syncthreads();
  }
  unsigned long long run = s_part[t] - sum;   // exclusive prefix of this thread's chunk
  for (int i = b; i < e; ++i) {
    const unsigned long long lo = run < static_cast<unsigned long long>(max_pairs) ? run : max_pairs;
    run += counts[i];
    const unsigned long long hi = run < static_cast<unsigned long long>(max_pairs) ? run : max_pairs;
    ranges[i] = make_uint2(static_cast<unsigned>(lo), static_cast<unsigned>(hi));
    cursor[i] = 0u;
  }
  if (t == 1023 && num_pairs_out != nullptr) {
    num_pairs_out[0] = static_cast<int64_t>(s_part[1023]);
    num_pairs_out[1] = static_cast<int64_t>(s_max[1023]);
  }
}

// Same aggregation for the scatter: the block counts per tile in shared memory, reserves one
// contiguous chunk per touched tile with a single global atomic, then hands out slots locally.
__global__ void tile_scatter_kernel(int G, int H, int W, const float2* __restrict__ xy,
                                    const float4* __restrict__ conic_o,
                                    const float4* __restrict__ rgbd,
                                    const int32_t* __restrict__ radii,
                                    const uint2* __restrict__ ranges, uint32_t* __restrict__ cursor,
                                    uint64_t* __restrict__ keys, int id_bits, int use_smem) {
  extern __shared__ uint32_t s_buf[];   // [tpv] counts -> local cursors, [tpv] chunk bases
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int tpv = gx * gy;
  const uint32_t v = blockIdx.y;
  uint32_t* s_cnt = s_buf;
  uint32_t* s_base = s_buf + tpv;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t i = static_cast<size_t>(v) * G + (g < G ? g : 0);
  int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
  float2 p = make_float2(0.f, 0.f), ext = make_float2(-1.f, -1.f);
  if (g < G) {
    const int rad = radii[i];
    if (rad > 0) {
      p = xy[i];
      tile_rect(p.x, p.y, rad, gx, gy, x0, x1, y0, y1);
      ext = alpha_extent(conic_o[i]);
    }
  }
  const uint64_t key = (static_cast<uint64_t>(__float_as_uint(g < G ? rgbd[i].w : 0.f)) << id_bits) | i;
  if (use_smem) {
    for (int t = threadIdx.x; t < tpv; t += blockDim.x) s_cnt[t] = 0u;
    __syncthreads();
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x)
        if (tile_may_hit(p.x, p.y, ext, x, y, W, H)) atomicAdd(s_cnt + y * gx + x, 1u);
    __syncthreads();
    for (int t = threadIdx.x; t < tpv; t += blockDim.x) {
      const uint32_t c = s_cnt[t];
      s_base[t] = c != 0u ? atomicAdd(cursor + v * tpv + t, c) : 0u;
      s_cnt[t] = 0u;
    }
    __syncthreads();
  }
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      if (!tile_may_hit(p.x, p.y, ext, x, y, W, H)) continue;
      const int t = y * gx + x;
      const uint2 r = ranges[v * tpv + t];
      const uint32_t pos = r.x + (use_smem ? s_base[t] + atomicAdd(s_cnt + t, 1u)
                                           : atomicAdd(cursor + v * tpv + t, 1u));
      if (pos < r.y) keys[pos] = key;
    }
}

constexpr int SORT_THREADS = 512;

// Sort one tile's pairs by (depth bits, Gaussian id) = the stable global order.
// Fast path: the 64-bit keys are split into a 32-bit depth key (minus the tile's minimum, so only
// the bits of the tile's depth RANGE are sorted: ~26 instead of 55 -> 6 radix passes instead of 11)
// and the id as the carried value.  The radix sort is stable but the scatter order is not, so a
// tile in which two splats share their depth bits exactly (a few percent of the tiles) is re-sorted
// on the full key.
template <int ITEMS>
__device__ __forceinline__ void sort_tile(unsigned char* smem, const uint2 r, int n,
                                          const uint64_t* __restrict__ keys,
                                          uint32_t* __restrict__ vals_out, int key_bits, int id_bits) {
  using Sort64 = cub::BlockRadixSort<uint64_t, SORT_THREADS, ITEMS, cub::NullType, 5>;
  using Sort32 = cub::BlockRadixSort<uint32_t, SORT_THREADS, ITEMS, uint32_t, 5>;
  __shared__ uint32_t s_edge[SORT_THREADS];
  __shared__ uint32_t s_red[2][SORT_THREADS / 32];
  const uint64_t id_mask = (1ull << id_bits) - 1ull;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t d[ITEMS], id[ITEMS];
  uint32_t dmin = 0xffffffffu, dmax = 0u;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int idx = tid * ITEMS + i;
    if (idx < n) {
      const uint64_t k = keys[r.x + idx];
      d[i] = static_cast<uint32_t>(k >> id_bits);
      id[i] = static_cast<uint32_t>(k & id_mask);
      dmin = min(dmin, d[i]);
      dmax = max(dmax, d[i]);
    } else {
      d[i] = 0xffffffffu;
      id[i] = 0u;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dmin = min(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  }
  if (lane == 0) { s_red[0][warp] = dmin; s_red[1][warp] = dmax; }
  __syncthreads();
  dmin = s_red[0][0]; dmax = s_red[1][0];
#pragma unroll
  for (int w = 1; w < SORT_THREADS / 32; ++w) { dmin = min(dmin, s_red[0][w]); dmax = max(dmax, s_red[1][w]); }
  const int nbits = 32 - __clz(dmax - dmin);   // 0 when every depth is equal
#pragma unroll
  for (int i = 0; i < ITEMS; ++i)
    if (tid * ITEMS + i < n) d[i] -= dmin;      // padding keeps all ones: stable sort leaves it last
  Sort32(*reinterpret_cast<typename Sort32::TempStorage*>(smem)).Sort(d, id, 0, nbits);
  // exact ties among valid neighbours (blocked arrangement: thread t holds elements t*ITEMS ..)
  s_edge[tid] = d[ITEMS - 1];
  __syncthreads();
  bool tie = false;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int idx = tid * ITEMS + i;
    const uint32_t prev = i > 0 ? d[i - 1] : (tid > 0 ? s_edge[tid - 1] : ~d[0]);
    tie |= idx > 0 && idx < n && d[i] == prev;
  }
  if (__syncthreads_or(tie)) {
    uint64_t k[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int idx = tid * ITEMS + i;
      k[i] = idx < n ? keys[r.x + idx] : ~0ull;
    }
    Sort64(*reinterpret_cast<typename Sort64::TempStorage*>(smem)).Sort(k, 0, key_bits);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) id[i] = static_cast<uint32_t>(k[i] & id_mask);
  }
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int idx = tid * ITEMS + i;
    if (idx < n) vals_out[r.x + idx] = id[i];
  }
}

template <int ITEMS>
constexpr size_t sort_smem_bytes() {
  using Sort64 = cub::BlockRadixSort<uint64_t, SORT_THREADS, ITEMS, cub::NullType, 5>;
  using Sort32 = cub::BlockRadixSort<uint32_t, SORT_THREADS, ITEMS, uint32_t, 5>;
  return sizeof(typename Sort64::TempStorage) > sizeof(typename Sort32::TempStorage)
             ? sizeof(typename Sort64::TempStorage) : sizeof(typename Sort32::TempStorage);
}

// tiles of up to 4096 pairs (the common case): one CTA per tile, size class chosen in the kernel --
// one launch instead of one per class (a class kernel whose CTAs all exit still costs ~35 us of
// block launches at 3072 tiles)
__global__ void __launch_bounds__(SORT_THREADS, 2)   // <= 64 registers: two (smem: three) CTAs per SM
    tile_sort_small_kernel(const uint2* __restrict__ ranges, const uint64_t* __restrict__ keys,
                           uint32_t* __restrict__ vals_out, int key_bits, int id_bits, int max_cap) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  const uint2 r = ranges[blockIdx.x];
  const int n = static_cast<int>(r.y - r.x);
  if (n > max_cap) {
    // larger than the caller's bound: ids are passed through UNSORTED (valid indices, wrong order);
    // the caller sees num_pairs_out[1] > max_tile_pairs and re-runs on the global-sort path
    const uint64_t id_mask = (1ull << id_bits) - 1ull;
    for (int i = threadIdx.x; i < n; i += SORT_THREADS)
      vals_out[r.x + i] = static_cast<uint32_t>(keys[r.x + i] & id_mask);
    return;
  }
  if (n == 0 || n > 4096) return;
  if (n <= 2048) sort_tile<4>(sort_smem, r, n, keys, vals_out, key_bits, id_bits);
  else if (n <= 3072) sort_tile<6>(sort_smem, r, n, keys, vals_out, key_bits, id_bits);
  else sort_tile<8>(sort_smem, r, n, keys, vals_out, key_bits, id_bits);
}

// tiles of 4097 .. 16384 pairs: a persistent grid walks the tile list (nothing to do on most
// scenes: a sweep over the ranges costs a few microseconds)
__global__ void __launch_bounds__(SORT_THREADS)
    tile_sort_large_kernel(const uint2* __restrict__ ranges, int n_tiles,
                           const uint64_t* __restrict__ keys, uint32_t* __restrict__ vals_out,
                           int key_bits, int id_bits, int max_cap) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const uint2 r = ranges[t];
    const int n = static_cast<int>(r.y - r.x);
    if (n <= 4096 || n > max_cap) continue;   // block-uniform
    if (n <= 5120) sort_tile<10>(sort_smem, r, n, keys, vals_out, key_bits, id_bits);
    else if (n <= 6144) sort_tile<12>(sort_smem, r, n, keys, vals_out, key_bits, id_bits);
    else if (n <= 8192) sort_tile<16>(sort_smem, r, n, keys, vals_out, key_bits, id_bits);
    else if (n <= 12288) sort_tile<24>(sort_smem, r, n, keys, vals_out, key_bits, id_bits);
    else sort_tile<32>(sort_smem, r, n, keys, vals_out, key_bits, id_bits);
    __syncthreads();   // shared memory is reused by the next tile
  }
}

int launch_tile_sorts(const Workspace& ws, int n_tiles, int key_bits, int id_bits, int max_cap,
                      cudaStream_t stream) {
  const int smem_small = static_cast<int>(sort_smem_bytes<8>());
  const int smem_large = static_cast<int>(sort_smem_bytes<32>());
    VS_CONFIGURE_PER_DEVICE(
    VS_CUDA(cudaFuncSetAttribute(tile_sort_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 smem_small));
    VS_CUDA(cudaFuncSetAttribute(tile_sort_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 smem_large));
  );
  tile_sort_small_kernel<<<n_tiles, SORT_THREADS, smem_small, stream>>>(
      ws.ranges, ws.keys_in, ws.vals_out, key_bits, id_bits, max_cap);
  VS_LAUNCH_CHECK();
  if (max_cap > 4096) {
    const int grid = n_tiles < 148 ? n_tiles : 148;
    tile_sort_large_kernel<<<grid, SORT_THREADS, smem_large, stream>>>(
        ws.ranges, n_tiles, ws.keys_in, ws.vals_out, key_bits, id_bits, max_cap);
    VS_LAUNCH_CHECK();
  }
  return VS_OK;
}

// Shared-memory loads by 32-bit shared address (computed ONCE per kernel): inside the divergent per-splat
// loop the compiler rebuilt the generic-to-shared window base (S2R SR_CgaCtaId + LEA) for every access.
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("mov.u32 %0, %0;" : "+r"(a));   // opaque: not to be re-materialised inside the loop
  return a;
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 t;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(a));
  return t;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 t;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "r"(a));
  return t;
}
// exp(x) for x <= 0 as ONE ex2.approx of x * log2(e): what __expf does minus its denormal-range fix-up (three
// more instructions), whose results (< 1e-38) can never reach the alpha >= 1/255 test.  The backward pass
// recomputes alpha with the same two instructions, so both passes take identical decisions.
__device__ __forceinline__ float exp_neg(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}

// ------------------------------------------------------------------ blend
// One CTA per 16x16 tile, 8 warps; warp w owns the 8x4-pixel sub-block (w & 1, w >> 1).  The tile's
// depth-sorted splat list is staged through shared memory in batches of 256 (16-byte gathers of
// the SoA intermediates).  Per batch every warp first culls: lane l tests splat (32 k + l) against
// the warp's sub-block with the exact bound of the alpha >= 1/255 test (the axis-aligned box of the
// ellipse power >= -ln(255 o)), ballots, and then walks only the set bits front to back -- pixels
// outside that box would have rejected the splat anyway, so the image is bit-identical to
// evaluating every (pixel, splat) pair.  A warp leaves the list as soon as all its pixels are opaque.
__global__ void __launch_bounds__(BLEND_THREADS, 6)  /* <= 40 registers: a blend CTA then fits beside a resident 320-thread GEMM CTA */
    blend_kernel(int G, int H, int W, const uint2* __restrict__ ranges,
                 const uint32_t* __restrict__ vals, const float2* __restrict__ xy,
                 const float4* __restrict__ conic_o, const float4* __restrict__ rgbd,
                 const float* __restrict__ bg, float* __restrict__ out_color,
                 float* __restrict__ out_depth, float* __restrict__ out_alpha,
                 float* __restrict__ final_T, int32_t* __restrict__ n_contrib,
                 int32_t* __restrict__ n_touched) {
  __shared__ uint32_t s_id[BLEND_THREADS];
  __shared__ float4 s_xe[BLEND_THREADS];   // centre x, y, half extents ex, ey of the alpha box
  __shared__ float4 s_co[BLEND_THREADS];
  __shared__ float4 s_cd[BLEND_THREADS];

  const int gx = gridDim.x, gy = gridDim.y;
  const int v = blockIdx.z;
  const int tile = (v * gy + blockIdx.y) * gx + blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sbx = blockIdx.x * TILE + (warp & 1) * 8, sby = blockIdx.y * TILE + (warp >> 1) * 4;
  const int px = sbx + (lane & 7);
  const int py = sby + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = static_cast<float>(px), pyf = static_cast<float>(py);
  // pixel-centre range of this warp's sub-block
  const float bx0 = static_cast<float>(sbx), bx1 = static_cast<float>(min(sbx + 7, W - 1));
  const float by0 = static_cast<float>(sby), by1 = static_cast<float>(min(sby + 3, H - 1));

  const uint2 range = ranges[tile];
  const int total = static_cast<int>(range.y - range.x);
  const int rounds = (total + BLEND_THREADS - 1) / BLEND_THREADS;

  bool done = !inside;
  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f, A = 0.f;
  int last_contributor = 0;
  const uint32_t a_xe = smem_addr(s_xe), a_co = smem_addr(s_co), a_cd = smem_addr(s_cd);

  for (int r = 0; r < rounds; ++r) {
    if (__syncthreads_count(done) == BLEND_THREADS) break;
    const int idx = r * BLEND_THREADS + threadIdx.x;
    if (idx < total) {
      const uint32_t id = vals[range.x + idx];
      const float2 p = xy[id];
      const float4 co = conic_o[id];
      s_id[threadIdx.x] = id;
      s_co[threadIdx.x] = co;
      s_cd[threadIdx.x] = rgbd[id];
      const float2 ext = alpha_extent(co);
      const float ex = ext.x, ey = ext.y;
      s_xe[threadIdx.x] = make_float4(p.x, p.y, ex, ey);
    }
    __syncthreads();
    const int nb = min(BLEND_THREADS, total - r * BLEND_THREADS);
    for (int base = 0; base < nb; base += 32) {
      if (__all_sync(0xffffffffu, done)) break;
      bool near_me = false;
      if (base + lane < nb) {
        const float4 xe = lds_f4(a_xe + (base + lane) * 16);
        const float ddx = fmaxf(fmaxf(bx0 - xe.x, xe.x - bx1), 0.f);
        const float ddy = fmaxf(fmaxf(by0 - xe.y, xe.y - by1), 0.f);
        near_me = ddx <= xe.z && ddy <= xe.w;
      }
      uint32_t todo = __ballot_sync(0xffffffffu, near_me);
      while (todo != 0u) {
        const int j = base + __ffs(todo) - 1;
        todo &= todo - 1;
        bool hit = false;
        if (!done) {
          const float2 xe = lds_f2(a_xe + j * 16);
          const float4 co = lds_f4(a_co + j * 16);
          const float dx = xe.x - pxf, dy = xe.y - pyf;
          const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
          if (power <= 0.0f) {
            const float alpha = fminf(kAlphaMax, co.w * exp_neg(power));
            if (alpha >= kAlphaMin) {
              const float test_T = T * (1.0f - alpha);
              if (test_T < kTStop) {
                done = true;
              } else {
                const float4 cd = lds_f4(a_cd + j * 16);
                const float w = alpha * T;
                C0 += cd.x * w; C1 += cd.y * w; C2 += cd.z * w;
                D += cd.w * w;
                A += w;
                hit = T > kNTouchedT;
                T = test_T;
                last_contributor = r * BLEND_THREADS + j + 1;
              }
            }
          }
        }
        if (n_touched != nullptr) {
          const uint32_t m = __ballot_sync(0xffffffffu, hit);
          if (m != 0u && lane == 0) atomicAdd(n_touched + s_id[j], __popc(m));
        }
        if (__all_sync(0xffffffffu, done)) break;
      }
    }
  }
  if (inside) {
    const size_t hw = static_cast<size_t>(H) * W;
    const size_t pix = static_cast<size_t>(py) * W + px;
    const float* b = bg + v * 3;
    out_color[(static_cast<size_t>(v) * 3 + 0) * hw + pix] = C0 + T * b[0];
    out_color[(static_cast<size_t>(v) * 3 + 1) * hw + pix] = C1 + T * b[1];
    out_color[(static_cast<size_t>(v) * 3 + 2) * hw + pix] = C2 + T * b[2];
    if (out_depth) out_depth[static_cast<size_t>(v) * hw + pix] = D;
    if (out_alpha) out_alpha[static_cast<size_t>(v) * hw + pix] = A;
    if (final_T) final_T[static_cast<size_t>(v) * hw + pix] = T;
    if (n_contrib) n_contrib[static_cast<size_t>(v) * hw + pix] = last_contributor;
  }
}

// ------------------------------------------------------------------ backward: blend
// Per-(view, Gaussian) gradient slots accumulated by the blend backward pass (SoA, stride VG)
enum { GA_X = 0, GA_Y, GA_CA, GA_CB, GA_CC, GA_OP, GA_R, GA_G, GA_B, GA_D, GA_COUNT };

// Same tile / sub-block / ballot-culling structure as the forward kernel, walked back to front.
// Each pixel replays only the splats in front of its last contributor (n_contrib), rebuilding
// T_before = T_after / (1 - alpha); the per-splat partials of the 8x4 pixels of a warp are reduced
// with shuffles and leave the warp as ONE atomicAdd per quantity.
__global__ void __launch_bounds__(BLEND_THREADS)
    blend_backward_kernel(int H, int W, const uint2* __restrict__ ranges,
                          const uint32_t* __restrict__ vals, const float2* __restrict__ xy,
                          const float4* __restrict__ conic_o, const float4* __restrict__ rgbd,
                          const float* __restrict__ bg, const float* __restrict__ final_T,
                          const int32_t* __restrict__ n_contrib, const float* __restrict__ dL_dcolor,
                          const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha,
                          float* __restrict__ gacc, size_t VG) {
  __shared__ uint32_t s_id[BLEND_THREADS];
  __shared__ float4 s_xe[BLEND_THREADS];
  __shared__ float4 s_co[BLEND_THREADS];
  __shared__ float4 s_cd[BLEND_THREADS];

  const int gx = gridDim.x, gy = gridDim.y;
  const int v = blockIdx.z;
  const int tile = (v * gy + blockIdx.y) * gx + blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sbx = blockIdx.x * TILE + (warp & 1) * 8, sby = blockIdx.y * TILE + (warp >> 1) * 4;
  const int px = sbx + (lane & 7), py = sby + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = static_cast<float>(px), pyf = static_cast<float>(py);
  const float bx0 = static_cast<float>(sbx), bx1 = static_cast<float>(min(sbx + 7, W - 1));
  const float by0 = static_cast<float>(sby), by1 = static_cast<float>(min(sby + 3, H - 1));
  const size_t hw = static_cast<size_t>(H) * W;
  const size_t pix = static_cast<size_t>(py) * W + px;

  float dpix[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float T_final = 1.f;
  int last = 0;
  if (inside) {
    T_final = final_T[static_cast<size_t>(v) * hw + pix];
    last = n_contrib[static_cast<size_t>(v) * hw + pix];
#pragma unroll
    for (int c = 0; c < 3; ++c) dpix[c] = dL_dcolor[(static_cast<size_t>(v) * 3 + c) * hw + pix];
    if (dL_ddepth) dpix[3] = dL_ddepth[static_cast<size_t>(v) * hw + pix];
    if (dL_dalpha) dpix[4] = dL_dalpha[static_cast<size_t>(v) * hw + pix];
  }
  const float bg_dot = bg[v * 3] * dpix[0] + bg[v * 3 + 1] * dpix[1] + bg[v * 3 + 2] * dpix[2];
  float T = T_final, last_alpha = 0.f;
  float last_val[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, accum[5] = {0.f, 0.f, 0.f, 0.f, 0.f};

  const uint2 range = ranges[tile];
  const int total = static_cast<int>(range.y - range.x);
  const int rounds = (total + BLEND_THREADS - 1) / BLEND_THREADS;
  for (int r = rounds - 1; r >= 0; --r) {
    __syncthreads();
    const int idx = r * BLEND_THREADS + threadIdx.x;
    if (idx < total) {
      const uint32_t id = vals[range.x + idx];
      const float2 p = xy[id];
      const float4 co = conic_o[id];
      s_id[threadIdx.x] = id;
      s_co[threadIdx.x] = co;
      s_cd[threadIdx.x] = rgbd[id];
      const float tau = __logf(255.0f * co.w) * 1.01f + 1e-3f;
      const float det = co.x * co.z - co.y * co.y;
      float ex = -1.f, ey = -1.f;
      if (tau > 0.f) {
        if (det > 0.f) {
          ex = sqrtf(2.0f * tau * co.z / det) + 0.01f;
          ey = sqrtf(2.0f * tau * co.x / det) + 0.01f;
        } else {
          ex = ey = 1e30f;
        }
      }
      s_xe[threadIdx.x] = make_float4(p.x, p.y, ex, ey);
    }
    __syncthreads();
    const int nb = min(BLEND_THREADS, total - r * BLEND_THREADS);
    for (int base = ((nb - 1) >> 5) << 5; base >= 0; base -= 32) {
      bool near_me = false;
      if (base + lane < nb) {
        const float4 xe = s_xe[base + lane];
        const float ddx = fmaxf(fmaxf(bx0 - xe.x, xe.x - bx1), 0.f);
        const float ddy = fmaxf(fmaxf(by0 - xe.y, xe.y - by1), 0.f);
        near_me = ddx <= xe.z && ddy <= xe.w;
      }
      uint32_t todo = __ballot_sync(0xffffffffu, near_me);
      while (todo != 0u) {
        const int bit = 31 - __clz(todo);   // back to front
        todo &= ~(1u << bit);
        const int j = base + bit;
        const int listidx = r * BLEND_THREADS + j;
        float g[GA_COUNT];
#pragma unroll
        for (int k = 0; k < GA_COUNT; ++k) g[k] = 0.f;
        bool contrib = false;
        if (listidx < last) {
          const float4 xe = s_xe[j];
          const float4 co = s_co[j];
          const float dx = xe.x - pxf, dy = xe.y - pyf;
          const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
          if (power <= 0.0f) {
            const float Gv = exp_neg(power);
            const float a_raw = co.w * Gv;
            const float alpha = fminf(kAlphaMax, a_raw);
            if (alpha >= kAlphaMin) {
              contrib = true;
              const float4 cd = s_cd[j];
              T = T / (1.0f - alpha);
              const float wgt = alpha * T;
              const float val[5] = {cd.x, cd.y, cd.z, cd.w, 1.0f};
              float dLa = 0.f;
#pragma unroll
              for (int c = 0; c < 5; ++c) {
                accum[c] = last_alpha * last_val[c] + (1.0f - last_alpha) * accum[c];
                last_val[c] = val[c];
                dLa += (val[c] - accum[c]) * dpix[c];
              }
              g[GA_R] = wgt * dpix[0]; g[GA_G] = wgt * dpix[1]; g[GA_B] = wgt * dpix[2];
              g[GA_D] = wgt * dpix[3];
              dLa *= T;
              last_alpha = alpha;
              dLa += (-T_final / (1.0f - alpha)) * bg_dot;
              const float dLraw = a_raw <= kAlphaMax ? dLa : 0.f;   // d min(0.99, x)
              const float dLdG = co.w * dLraw * Gv;                   // includes dG/dpower = G
              g[GA_OP] = Gv * dLraw;
              g[GA_X] = dLdG * (-(co.x * dx + co.y * dy));
              g[GA_Y] = dLdG * (-(co.z * dy + co.y * dx));
              g[GA_CA] = dLdG * (-0.5f * dx * dx);
              g[GA_CB] = dLdG * (-dx * dy);
              g[GA_CC] = dLdG * (-0.5f * dy * dy);
            }
          }
        }
        if (__ballot_sync(0xffffffffu, contrib) != 0u) {
          // Sum the 10 partials over the 32 pixels of the warp with 12 shuffles instead of 50: every
          // butterfly step HALVES the set of values a lane is responsible for (it hands the other half to
          // its partner), so after the five steps lane l holds the warp total of ONE value
          //   k = 5 * bit4 + {0, 1, 2 | 3, 4}   (selected by bits 3..1; six of the 16 bit patterns are idle)
          // and ten lanes issue one atomic each.
          const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
          float a5[5], b[3], c[2];
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const float send = b4 ? g[i] : g[i + 5], keep = b4 ? g[i + 5] : g[i];
            a5[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float hi = i < 2 ? a5[i + 3] : 0.f;
            const float send = b3 ? a5[i] : hi, keep = b3 ? hi : a5[i];
            b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float hi = i < 1 ? b[i + 2] : 0.f;
            const float send = b2 ? b[i] : hi, keep = b2 ? hi : b[i];
            c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
          float d = (b1 ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, b1 ? c[0] : c[1], 2);
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          // value index held by this lane (-1: none)
          int k = -1;
          if (!b3) k = !b2 ? (b1 ? 1 : 0) : (b1 ? -1 : 2);
          else k = !b2 ? (b1 ? 4 : 3) : -1;
          if (k >= 0 && !(lane & 1)) atomicAdd(gacc + (k + (b4 ? 5 : 0)) * VG + s_id[j], d);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ backward: per Gaussian
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) r += red[i];
  return r;
}

// One thread per Gaussian, looping over the views that share it: turns the screen-space partials
// of the blend pass into gradients of means3D / cov6 / opacity / SH (or colours) and, reduced over
// the block, of the camera twist tau = (rho, theta) of the left perturbation W2C <- exp(tau) W2C
// (src/misc/cam_utils.py:123-142).  Assumes projmatrix = viewmatrix o standard perspective with
// the given tan(fov), which is how cuda_splatting.py:187-194 builds it.
__global__ void __launch_bounds__(128, 4)   // 128 registers (a few spills) beat 242 at two CTAs per SM: the pass is latency-bound
    preprocess_backward_kernel(int G, int V, int shared_set, int H, int W,
                               const float* __restrict__ means, const float* __restrict__ cov6,
                               const float* __restrict__ shs, int M, int sh_cs, int sh_ch,
                               int degree, int has_colors, const float* __restrict__ viewm,
                               const float* __restrict__ campos, const float* __restrict__ tanfov,
                               const int32_t* __restrict__ radii, const float4* __restrict__ rgbd,
                               const float* __restrict__ gacc, size_t VG,
                               float* __restrict__ d_means, float* __restrict__ d_cov,
                               float* __restrict__ d_opac, float* __restrict__ d_shs,
                               float* __restrict__ d_colors, float* __restrict__ d_tau) {
  __shared__ float red[4];
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = g < G;
  const int v_begin = shared_set ? 0 : blockIdx.y;
  const int v_end = shared_set ? V : blockIdx.y + 1;
  const size_t gi = (shared_set ? 0 : static_cast<size_t>(blockIdx.y) * G) + (live ? g : 0);
  const int n_sh = shs ? (degree >= 3 ? 16 : (degree + 1) * (degree + 1)) : 0;

  float mx = 0.f, my = 0.f, mz = 0.f, c6[6] = {0, 0, 0, 0, 0, 0};
  if (live) {
    mx = means[gi * 3]; my = means[gi * 3 + 1]; mz = means[gi * 3 + 2];
#pragma unroll
    for (int i = 0; i < 6; ++i) c6[i] = cov6[gi * 6 + i];
  }
  float dmean[3] = {0.f, 0.f, 0.f}, dcov[6] = {0, 0, 0, 0, 0, 0}, dop = 0.f, dcol[3] = {0, 0, 0};
  float dsh[48];
#pragma unroll
  for (int i = 0; i < 48; ++i) dsh[i] = 0.f;
  // SH coefficients of the warp's 32 Gaussians staged through shared memory with coalesced loads, like the
  // forward kernel (thread-per-Gaussian reads of the (G, 3, M) array touch 32 lines per instruction, 48 times
  // per view); the same buffer carries d SH back out at the end
  extern __shared__ float sh_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = 3 * M;
  float* wsh = sh_smem + static_cast<size_t>(warp) * 32 * per;
  const int gbase = blockIdx.x * blockDim.x + warp * 32;
  const int cnt = max(0, min(32, G - gbase)) * per;
  const size_t set_off = shared_set ? 0 : static_cast<size_t>(blockIdx.y) * G;
  const float* my_sh = nullptr;
  if (shs != nullptr) {
    const float* src = shs + (set_off + gbase) * per;
    for (int i = lane; i < cnt; i += 32) wsh[i] = __ldg(src + i);
    __syncwarp();
    my_sh = wsh + lane * per;
  }

  for (int v = v_begin; v < v_end; ++v) {
    const float* vm = viewm + v * 16;
    float dtau[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const size_t o = static_cast<size_t>(v) * G + (live ? g : 0);
    if (live && radii[o] > 0) {
      const float gX = gacc[GA_X * VG + o], gY = gacc[GA_Y * VG + o];
      const float gA = gacc[GA_CA * VG + o], gB = gacc[GA_CB * VG + o], gC = gacc[GA_CC * VG + o];
      dop += gacc[GA_OP * VG + o];
      float gr[3] = {gacc[GA_R * VG + o], gacc[GA_G * VG + o], gacc[GA_B * VG + o]};
      const float gD = gacc[GA_D * VG + o];
      const float4 cd = rgbd[o];
      // ---- forward recomputation
      const float tx = vm[0] * mx + vm[4] * my + vm[8] * mz + vm[12];
      const float ty = vm[1] * mx + vm[5] * my + vm[9] * mz + vm[13];
      const float tz = vm[2] * mx + vm[6] * my + vm[10] * mz + vm[14];
      const float tanx = tanfov[v * 2], tany = tanfov[v * 2 + 1];
      const float fx = W / (2.0f * tanx), fy = H / (2.0f * tany);
      const float limx = kFovClamp * tanx, limy = kFovClamp * tany;
      const float rx = tx / tz, ry = ty / tz;
      const bool clx = rx < -limx || rx > limx, cly = ry < -limy || ry > limy;
      const float txc = fminf(limx, fmaxf(-limx, rx)) * tz;
      const float tyc = fminf(limy, fmaxf(-limy, ry)) * tz;
      const float itz = 1.0f / tz, itz2 = itz * itz;
      const float j00 = fx * itz, j02 = -fx * txc * itz2, j11 = fy * itz, j12 = -fy * tyc * itz2;
      const float T0[3] = {j00 * vm[0] + j02 * vm[2], j00 * vm[4] + j02 * vm[6], j00 * vm[8] + j02 * vm[10]};
      const float T1[3] = {j11 * vm[1] + j12 * vm[2], j11 * vm[5] + j12 * vm[6], j11 * vm[9] + j12 * vm[10]};
      const float S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
      float ST0[3], ST1[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        ST0[i] = S[i][0] * T0[0] + S[i][1] * T0[1] + S[i][2] * T0[2];
        ST1[i] = S[i][0] * T1[0] + S[i][1] * T1[1] + S[i][2] * T1[2];
      }
      const float a = T0[0] * ST0[0] + T0[1] * ST0[1] + T0[2] * ST0[2] + kLowpass;
      const float b = T0[0] * ST1[0] + T0[1] * ST1[1] + T0[2] * ST1[2];
      const float c = T1[0] * ST1[0] + T1[1] * ST1[1] + T1[2] * ST1[2] + kLowpass;
      const float det = a * c - b * b;
      const float id = 1.0f / det, id2 = id * id;
      // ---- conic (c, -b, a) / det  ->  a, b, c
      const float dLa = gA * (-c * c * id2) + gB * (b * c * id2) + gC * (id - a * c * id2);
      const float dLc = gA * (id - a * c * id2) + gB * (a * b * id2) + gC * (-a * a * id2);
      const float dLb = gA * (2.f * b * c * id2) + gB * (-id - 2.f * b * b * id2) + gC * (2.f * a * b * id2);
      // ---- cov2D = T S T^T -> S (6 unique entries) and T
      dcov[0] += dLa * T0[0] * T0[0] + dLb * T0[0] * T1[0] + dLc * T1[0] * T1[0];
      dcov[3] += dLa * T0[1] * T0[1] + dLb * T0[1] * T1[1] + dLc * T1[1] * T1[1];
      dcov[5] += dLa * T0[2] * T0[2] + dLb * T0[2] * T1[2] + dLc * T1[2] * T1[2];
      dcov[1] += 2.f * dLa * T0[0] * T0[1] + dLb * (T0[0] * T1[1] + T0[1] * T1[0]) + 2.f * dLc * T1[0] * T1[1];
      dcov[2] += 2.f * dLa * T0[0] * T0[2] + dLb * (T0[0] * T1[2] + T0[2] * T1[0]) + 2.f * dLc * T1[0] * T1[2];
      dcov[4] += 2.f * dLa * T0[1] * T0[2] + dLb * (T0[1] * T1[2] + T0[2] * T1[1]) + 2.f * dLc * T1[1] * T1[2];
      float dT0[3], dT1[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        dT0[i] = 2.f * dLa * ST0[i] + dLb * ST1[i];
        dT1[i] = 2.f * dLc * ST1[i] + dLb * ST0[i];
      }
      // ---- T = J Wrot (Wrot[r][i] = vm[i*4 + r])
      const float W0[3] = {vm[0], vm[4], vm[8]}, W1[3] = {vm[1], vm[5], vm[9]}, W2[3] = {vm[2], vm[6], vm[10]};
      const float dj00 = dT0[0] * W0[0] + dT0[1] * W0[1] + dT0[2] * W0[2];
      const float dj02 = dT0[0] * W2[0] + dT0[1] * W2[1] + dT0[2] * W2[2];
      const float dj11 = dT1[0] * W1[0] + dT1[1] * W1[1] + dT1[2] * W1[2];
      const float dj12 = dT1[0] * W2[0] + dT1[1] * W2[1] + dT1[2] * W2[2];
      float dW0[3], dW1[3], dW2[3];   // gradient w.r.t. the rows of Wrot (for the pose)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        dW0[i] = dT0[i] * j00;
        dW1[i] = dT1[i] * j11;
        dW2[i] = dT0[i] * j02 + dT1[i] * j12;
      }
      // ---- J(t) with the frustum clamp, depth, and the projected centre
      float dt[3];
      dt[0] = dj02 * (-fx * itz2) * (clx ? 0.f : 1.f);
      dt[1] = dj12 * (-fy * itz2) * (cly ? 0.f : 1.f);
      dt[2] = dj00 * (-fx * itz2) + dj11 * (-fy * itz2) +
              dj02 * (2.f * fx * txc * itz2 * itz - (clx ? fx * itz2 * (txc * itz) : 0.f)) +
              dj12 * (2.f * fy * tyc * itz2 * itz - (cly ? fy * itz2 * (tyc * itz) : 0.f)) + gD;
      const float pw = 1.0f / (tz + kWEps);
      const float hx = tx / tanx, hy = ty / tany;
      const float sxp = gX * 0.5f * W, syp = gY * 0.5f * H;
      dt[0] += sxp * pw / tanx;
      dt[1] += syp * pw / tany;
      dt[2] += -(sxp * hx + syp * hy) * pw * pw;
      // ---- colour: SH (bands 0..3) of dir = normalize(p - campos), clamped at 0
      float dcam[3] = {0.f, 0.f, 0.f};
      if (has_colors) {
        dcol[0] += gr[0]; dcol[1] += gr[1]; dcol[2] += gr[2];
      } else if (my_sh != nullptr) {
        if (cd.x <= 0.f) gr[0] = 0.f;
        if (cd.y <= 0.f) gr[1] = 0.f;
        if (cd.z <= 0.f) gr[2] = 0.f;
        float ux = mx - campos[v * 3], uy = my - campos[v * 3 + 1], uz = mz - campos[v * 3 + 2];
        const float il = 1.0f / sqrtf(ux * ux + uy * uy + uz * uz);
        const float x = ux * il, y = uy * il, z = uz * il;
        const float xx = x * x, yy = y * y, zz = z * z, xy_ = x * y, yz = y * z, xz = x * z;
        float bas[16], bdx[16], bdy[16], bdz[16];
        bas[0] = SH_C0; bdx[0] = bdy[0] = bdz[0] = 0.f;
        bas[1] = -SH_C1 * y; bdx[1] = 0.f; bdy[1] = -SH_C1; bdz[1] = 0.f;
        bas[2] = SH_C1 * z; bdx[2] = 0.f; bdy[2] = 0.f; bdz[2] = SH_C1;
        bas[3] = -SH_C1 * x; bdx[3] = -SH_C1; bdy[3] = 0.f; bdz[3] = 0.f;
        bas[4] = SH_C2[0] * xy_; bdx[4] = SH_C2[0] * y; bdy[4] = SH_C2[0] * x; bdz[4] = 0.f;
        bas[5] = SH_C2[1] * yz; bdx[5] = 0.f; bdy[5] = SH_C2[1] * z; bdz[5] = SH_C2[1] * y;
        bas[6] = SH_C2[2] * (2.f * zz - xx - yy);
        bdx[6] = SH_C2[2] * -2.f * x; bdy[6] = SH_C2[2] * -2.f * y; bdz[6] = SH_C2[2] * 4.f * z;
        bas[7] = SH_C2[3] * xz; bdx[7] = SH_C2[3] * z; bdy[7] = 0.f; bdz[7] = SH_C2[3] * x;
        bas[8] = SH_C2[4] * (xx - yy); bdx[8] = SH_C2[4] * 2.f * x; bdy[8] = SH_C2[4] * -2.f * y; bdz[8] = 0.f;
        bas[9] = SH_C3[0] * y * (3.f * xx - yy);
        bdx[9] = SH_C3[0] * 6.f * xy_; bdy[9] = SH_C3[0] * (3.f * xx - 3.f * yy); bdz[9] = 0.f;
        bas[10] = SH_C3[1] * xy_ * z; bdx[10] = SH_C3[1] * yz; bdy[10] = SH_C3[1] * xz; bdz[10] = SH_C3[1] * xy_;
        bas[11] = SH_C3[2] * y * (4.f * zz - xx - yy);
        bdx[11] = SH_C3[2] * -2.f * xy_; bdy[11] = SH_C3[2] * (4.f * zz - xx - 3.f * yy); bdz[11] = SH_C3[2] * 8.f * yz;
        bas[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
        bdx[12] = SH_C3[3] * -6.f * xz; bdy[12] = SH_C3[3] * -6.f * yz; bdz[12] = SH_C3[3] * (6.f * zz - 3.f * xx - 3.f * yy);
        bas[13] = SH_C3[4] * x * (4.f * zz - xx - yy);
        bdx[13] = SH_C3[4] * (4.f * zz - 3.f * xx - yy); bdy[13] = SH_C3[4] * -2.f * xy_; bdz[13] = SH_C3[4] * 8.f * xz;
        bas[14] = SH_C3[5] * z * (xx - yy);
        bdx[14] = SH_C3[5] * 2.f * xz; bdy[14] = SH_C3[5] * -2.f * yz; bdz[14] = SH_C3[5] * (xx - yy);
        bas[15] = SH_C3[6] * x * (xx - 3.f * yy);
        bdx[15] = SH_C3[6] * (3.f * xx - 3.f * yy); bdy[15] = SH_C3[6] * -6.f * xy_; bdz[15] = 0.f;
        float ddx = 0.f, ddy = 0.f, ddz = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (k < n_sh) {
            float wsum = 0.f;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
              dsh[k * 3 + ch] += bas[k] * gr[ch];
              wsum += gr[ch] * my_sh[k * sh_cs + ch * sh_ch];
            }
            ddx += wsum * bdx[k]; ddy += wsum * bdy[k]; ddz += wsum * bdz[k];
          }
        }
        // through the normalisation
        const float dot = ddx * x + ddy * y + ddz * z;
        const float dux = (ddx - x * dot) * il, duy = (ddy - y * dot) * il, duz = (ddz - z * dot) * il;
        dmean[0] += dux; dmean[1] += duy; dmean[2] += duz;
        dcam[0] = -dux; dcam[1] = -duy; dcam[2] = -duz;
      }
      // ---- t = Wrot p + tw  ->  p
      dmean[0] += W0[0] * dt[0] + W1[0] * dt[1] + W2[0] * dt[2];
      dmean[1] += W0[1] * dt[0] + W1[1] * dt[1] + W2[1] * dt[2];
      dmean[2] += W0[2] * dt[0] + W1[2] * dt[1] + W2[2] * dt[2];
      // ---- camera twist at tau = 0: t' = t + rho + theta x t ; Wrot' = (I + [theta]x) Wrot ;
      //      campos' = campos - Wrot^T rho
      dtau[0] = dt[0] - (W0[0] * dcam[0] + W0[1] * dcam[1] + W0[2] * dcam[2]);
      dtau[1] = dt[1] - (W1[0] * dcam[0] + W1[1] * dcam[1] + W1[2] * dcam[2]);
      dtau[2] = dt[2] - (W2[0] * dcam[0] + W2[1] * dcam[1] + W2[2] * dcam[2]);
      float th[3] = {ty * dt[2] - tz * dt[1], tz * dt[0] - tx * dt[2], tx * dt[1] - ty * dt[0]};
#pragma unroll
      for (int i = 0; i < 3; ++i) {   // column i of Wrot crossed with column i of dL/dWrot
        const float wx = W0[i], wy = W1[i], wz = W2[i];
        const float gx_ = dW0[i], gy_ = dW1[i], gz_ = dW2[i];
        th[0] += wy * gz_ - wz * gy_;
        th[1] += wz * gx_ - wx * gz_;
        th[2] += wx * gy_ - wy * gx_;
      }
      dtau[3] = th[0]; dtau[4] = th[1]; dtau[5] = th[2];
    }
    if (d_tau != nullptr) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const float s = block_sum(dtau[k], red);
        if (threadIdx.x == 0 && s != 0.f) atomicAdd(d_tau + v * 6 + k, s);
      }
    }
  }
  if (live) {
    d_means[gi * 3] += dmean[0]; d_means[gi * 3 + 1] += dmean[1]; d_means[gi * 3 + 2] += dmean[2];
#pragma unroll
    for (int i = 0; i < 6; ++i) d_cov[gi * 6 + i] += dcov[i];
    d_opac[gi] += dop;
    if (has_colors && d_colors != nullptr) {
      d_colors[gi * 3] += dcol[0]; d_colors[gi * 3 + 1] += dcol[1]; d_colors[gi * 3 + 2] += dcol[2];
    }
  }
  if (!has_colors && d_shs != nullptr && shs != nullptr) {   // warp-uniform
    __syncwarp();
    for (int i = lane; i < cnt; i += 32) wsh[i] = 0.f;       // bands above the active degree get no gradient
    __syncwarp();
    if (live) {
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < n_sh)
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) wsh[lane * per + k * sh_cs + ch * sh_ch] = dsh[k * 3 + ch];
    }
    __syncwarp();
    float* dst = d_shs + (set_off + gbase) * per;
    for (int i = lane; i < cnt; i += 32) dst[i] += wsh[i];
  }
}

int highest_bit(uint64_t x) {
  int b = 0;
  while (x) { ++b; x >>= 1; }
  return b;
}

}  // namespace
}  // namespace vs

extern "C" int64_t vs_raster_workspace_bytes(int V, int G, int H, int W, int64_t max_pairs) {
  if (V <= 0 || G <= 0 || H <= 0 || W <= 0 || max_pairs <= 0) return 0;
  return static_cast<int64_t>(vs::carve(nullptr, V, G, H, W, max_pairs).total);
}

extern "C" int vs_raster_forward(const vs_raster_params* p, vs_stream_t stream_) {
  using namespace vs;
  VS_REQUIRE(p != nullptr, "vs_raster_forward: null params");
  VS_REQUIRE(p->V > 0 && p->H > 0 && p->W > 0 && p->G >= 0, "vs_raster_forward: bad sizes");
  if (p->G > 0) {
    VS_REQUIRE(p->means3D && p->cov3D && p->opacities,
               "vs_raster_forward: missing Gaussian inputs");
    VS_REQUIRE(p->shs != nullptr || p->colors_precomp != nullptr,
               "vs_raster_forward: provide either shs or colors_precomp");
    VS_REQUIRE(!(p->shs != nullptr && p->colors_precomp != nullptr),
               "vs_raster_forward: provide only one of shs / colors_precomp");
  }
  VS_REQUIRE(p->viewmatrix && p->projmatrix && p->campos && p->tanfov && p->bg,
             "vs_raster_forward: missing camera inputs");
  VS_REQUIRE(p->out_color != nullptr, "vs_raster_forward: out_color is required");
  VS_REQUIRE(static_cast<int64_t>(p->V) * p->G < (1ll << 31), "vs_raster_forward: V*G too large");
  cudaStream_t stream = to_stream(stream_);
  const int gx = ceil_div(p->W, TILE), gy = ceil_div(p->H, TILE);
  const size_t hw = static_cast<size_t>(p->H) * p->W;

  VS_REQUIRE(p->workspace != nullptr && p->max_pairs > 0, "vs_raster_forward: workspace required");
  const int Gc = p->G > 0 ? p->G : 1;
  Workspace ws = carve(p->workspace, p->V, Gc, p->H, p->W, p->max_pairs);
  if (static_cast<int64_t>(ws.total) > p->workspace_bytes) {
    set_error("vs_raster_forward: workspace too small (%lld < %lld bytes)",
              (long long)p->workspace_bytes, (long long)ws.total);
    return VS_ERR_WORKSPACE;
  }
  const size_t VG = static_cast<size_t>(p->V) * p->G;
  const size_t n_tiles = static_cast<size_t>(p->V) * gx * gy;
  int32_t* radii = p->radii ? p->radii : ws.radii;
  VS_CUDA(cudaMemsetAsync(ws.ranges, 0, n_tiles * sizeof(uint2), stream));
  if (p->n_touched) VS_CUDA(cudaMemsetAsync(p->n_touched, 0, VG * sizeof(int32_t), stream));
  if (p->num_pairs_out) VS_CUDA(cudaMemsetAsync(p->num_pairs_out, 0, 16, stream));

  if (p->G > 0) {
    int sh_cs = p->sh_stride_coef, sh_ch = p->sh_stride_chan;
    if (sh_cs == 0 && sh_ch == 0) { sh_cs = 3; sh_ch = 1; }  // reference layout (G, M, 3)
    if (p->shs) {
      VS_REQUIRE(p->sh_M >= 1 && p->sh_M <= 32, "vs_raster_forward: sh_M out of range");
      const int deg = p->sh_degree > 3 ? 3 : p->sh_degree;
      VS_REQUIRE((deg + 1) * (deg + 1) <= p->sh_M,
                 "vs_raster_forward: sh_degree needs more coefficients than sh_M");
    }
    const int threads = 128;
    dim3 grid(ceil_div(p->G, threads), p->gaussians_shared ? 1 : p->V);
    const size_t smem = p->shs ? static_cast<size_t>(threads) * 3 * p->sh_M * sizeof(float) : 0;
    if (smem > 48 * 1024) {
      VS_CUDA(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    }
    preprocess_kernel<<<grid, threads, smem, stream>>>(
        p->G, p->V, p->gaussians_shared, p->H, p->W, p->means3D, p->cov3D, p->opacities, p->shs,
        p->sh_M, sh_cs, sh_ch, p->sh_degree, p->colors_precomp, p->viewmatrix, p->projmatrix,
        p->campos, p->tanfov, ws.xy, ws.conic_o, ws.rgbd, ws.tiles, radii);
    VS_LAUNCH_CHECK();

    if (p->max_tile_pairs > 0 && p->max_tile_pairs <= 16384 && n_tiles < (1u << 20)) {
      // ---- per-tile binning + shared-memory sort
      VS_CUDA(cudaMemsetAsync(ws.tile_counts, 0, n_tiles * 4, stream));
      const int tpv = gx * gy;
      const int use_smem = tpv <= 4096;
      dim3 cgrid(ceil_div(p->G, 256), p->V);
      tile_count_kernel<<<cgrid, 256, use_smem ? tpv * 4 : 0, stream>>>(
          p->G, p->H, p->W, ws.xy, ws.conic_o, radii, ws.tile_counts, use_smem);
      VS_LAUNCH_CHECK();
      tile_offsets_kernel<<<1, 1024, 0, stream>>>(ws.tile_counts, static_cast<int>(n_tiles),
                                                  p->max_pairs, ws.ranges, ws.tile_cursor,
                                                  p->num_pairs_out);
      VS_LAUNCH_CHECK();
      const int id_bits = highest_bit(VG > 1 ? VG - 1 : 1);
      tile_scatter_kernel<<<cgrid, 256, use_smem ? tpv * 8 : 0, stream>>>(
          p->G, p->H, p->W, ws.xy, ws.conic_o, ws.rgbd, radii, ws.ranges, ws.tile_cursor, ws.keys_in, id_bits,
          use_smem);
      VS_LAUNCH_CHECK();
      const int key_bits = 32 + id_bits, nt = static_cast<int>(n_tiles);
      const int mt = p->max_tile_pairs;
      // size classes (capacity = 512 x ITEMS): a tile is sorted at the narrowest class that fits
      static const int caps[8] = {2048, 3072, 4096, 5120, 6144, 8192, 12288, 16384};
      int max_cap = 16384;
      for (int c = 7; c >= 0; --c)
        if (mt <= caps[c]) max_cap = caps[c];
      int rc = launch_tile_sorts(ws, nt, key_bits, id_bits, max_cap, stream);
      if (rc != VS_OK) return rc;
    } else {
    // ---- global path: scan, emit (tile | depth) keys, one 64-bit radix sort
    size_t tmp = ws.cub_bytes;
    VS_CUDA(cub::DeviceScan::InclusiveSum(ws.cub_tmp, tmp, ws.tiles, ws.offsets,
                                          static_cast<int>(VG), stream));
    emit_pairs_kernel<<<static_cast<unsigned>(ceil_div64(VG, 256)), 256, 0, stream>>>(
        VG, p->G, p->H, p->W, ws.xy, ws.rgbd, radii, ws.offsets, ws.keys_in, ws.vals_in,
        p->max_pairs);
    VS_LAUNCH_CHECK();
    pad_keys_kernel<<<static_cast<unsigned>(ceil_div64(p->max_pairs, 256)), 256, 0, stream>>>(
        ws.offsets, VG, ws.keys_in, p->max_pairs, p->num_pairs_out);
    VS_LAUNCH_CHECK();
    const int end_bit = 32 + highest_bit(n_tiles);  // padded keys (all ones) sort last
    tmp = ws.cub_bytes;
    VS_CUDA(cub::DeviceRadixSort::SortPairs(ws.cub_tmp, tmp, ws.keys_in, ws.keys_out, ws.vals_in,
                                            ws.vals_out, static_cast<int>(p->max_pairs), 0,
                                            end_bit, stream));
    tile_ranges_kernel<<<static_cast<unsigned>(ceil_div64(p->max_pairs, 256)), 256, 0, stream>>>(
        ws.keys_out, p->max_pairs, ws.offsets, VG, ws.ranges);
    VS_LAUNCH_CHECK();
    }
  }
  dim3 bgrid(gx, gy, p->V);
  blend_kernel<<<bgrid, BLEND_THREADS, 0, stream>>>(
      p->G, p->H, p->W, ws.ranges, ws.vals_out, ws.xy, ws.conic_o, ws.rgbd, p->bg, p->out_color,
      p->out_depth, p->out_alpha, p->final_T, p->n_contrib, p->n_touched);
  VS_LAUNCH_CHECK();
  (void)hw;
  return VS_OK;
}

extern "C" int vs_raster_backward(const vs_raster_bwd_params* p, vs_stream_t stream_) {
  using namespace vs;
  VS_REQUIRE(p != nullptr, "vs_raster_backward: null params");
  const vs_raster_params& f = p->fwd;
  VS_REQUIRE(f.V > 0 && f.H > 0 && f.W > 0 && f.G >= 0, "vs_raster_backward: bad sizes");
  if (f.G == 0) return VS_OK;
  VS_REQUIRE(p->dL_dcolor != nullptr, "vs_raster_backward: dL_dcolor is required");
  VS_REQUIRE(p->dL_dmeans3D && p->dL_dcov3D && p->dL_dopacity, "vs_raster_backward: missing outputs");
  VS_REQUIRE(f.final_T && f.n_contrib && f.radii && f.workspace,
             "vs_raster_backward: forward state (final_T, n_contrib, radii, workspace) is required");
  VS_REQUIRE(f.shs == nullptr || p->dL_dshs != nullptr, "vs_raster_backward: dL_dshs is required");
  VS_REQUIRE(f.colors_precomp == nullptr || p->dL_dcolors != nullptr,
             "vs_raster_backward: dL_dcolors is required");
  const size_t VG = static_cast<size_t>(f.V) * f.G;
  VS_REQUIRE(p->bwd_workspace != nullptr &&
                 p->bwd_workspace_bytes >= static_cast<int64_t>(VG * GA_COUNT * sizeof(float)),
             "vs_raster_backward: bwd_workspace too small (need V*G*10 floats)");
  cudaStream_t stream = to_stream(stream_);
  Workspace ws = carve(f.workspace, f.V, f.G, f.H, f.W, f.max_pairs);
  float* gacc = static_cast<float*>(p->bwd_workspace);
  VS_CUDA(cudaMemsetAsync(gacc, 0, VG * GA_COUNT * sizeof(float), stream));
  const int gx = ceil_div(f.W, TILE), gy = ceil_div(f.H, TILE);
  dim3 bgrid(gx, gy, f.V);
  blend_backward_kernel<<<bgrid, BLEND_THREADS, 0, stream>>>(
      f.H, f.W, ws.ranges, ws.vals_out, ws.xy, ws.conic_o, ws.rgbd, f.bg, f.final_T, f.n_contrib,
      p->dL_dcolor, p->dL_ddepth, p->dL_dalpha, gacc, VG);
  VS_LAUNCH_CHECK();
  int sh_cs = f.sh_stride_coef, sh_ch = f.sh_stride_chan;
  if (sh_cs == 0 && sh_ch == 0) { sh_cs = 3; sh_ch = 1; }
  dim3 grid(ceil_div(f.G, 128), f.gaussians_shared ? 1 : f.V);
  const size_t sh_smem_bytes = f.shs != nullptr ? static_cast<size_t>(128) * 3 * f.sh_M * sizeof(float) : 0;
  preprocess_backward_kernel<<<grid, 128, sh_smem_bytes, stream>>>(
      f.G, f.V, f.gaussians_shared, f.H, f.W, f.means3D, f.cov3D, f.shs, f.sh_M, sh_cs, sh_ch,
      f.sh_degree, f.colors_precomp != nullptr, f.viewmatrix, f.campos, f.tanfov, f.radii, ws.rgbd,
      gacc, VG, p->dL_dmeans3D, p->dL_dcov3D, p->dL_dopacity, p->dL_dshs, p->dL_dcolors, p->dL_dtau);
  VS_LAUNCH_CHECK();
  return VS_OK;
}
