// Multi-level RoIAlign (aligned=True, adaptive sampling) forward / backward on channels-last feature maps.
//
// Replaces detectron2 ROIPooler -> torchvision.ops.roi_align (per-level launches + index_put scatter):
// one launch covers all FPN levels; the level of each RoI is computed in-kernel
// (floor(4 + log2(sqrt(area)/224 + 1e-8)) clamped to [2,5], D2 poolers.py assign_boxes_to_levels).
// One warp per (RoI, output bin); lanes span channels so every bilinear tap is a coalesced 512 B
// (bf16) / 1 KB (fp32) row read of a 256-channel pixel.  Output is (M, 7, 7, C): the flatten order the
// box head's fc1 operand is packed for.  Backward scatters with 128-bit fp32 reductions.
#include "common.cuh"
#include "../../include/aldi_b200.h"
#include <stdlib.h>

namespace {

struct RoiArgs {
  const void* feat[4];
  float* dfeat[4];
  int fh[4], fw[4];
  float scale[4];
  int c, pooled;
  const float* rois;      // (M,4) xyxy in input-image pixels
  const int* roi_batch;   // (M)
  const int* num_valid;   // device count (nullable)
  int m;
  void* out;              // fwd: (M,P,P,C)
  const void* dout;       // bwd
  int min_level, max_level;
  float canonical_size;
  int canonical_level;
};

__device__ __forceinline__ int roi_level(const float* b, const RoiArgs& a) {
  const float area = (b[2] - b[0]) * (b[3] - b[1]);
  const float sz = sqrtf(area);
  float lvl = floorf((float)a.canonical_level + log2f(sz / a.canonical_size + 1e-8f));
  lvl = fminf(fmaxf(lvl, (float)a.min_level), (float)a.max_level);
  return (int)lvl - a.min_level;
}

template <typename T, int VEC>
__device__ __forceinline__ void load_vec(const T* p, float* f);
template <>
__device__ __forceinline__ void load_vec<float, 4>(const float* p, float* f) {
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <>
__device__ __forceinline__ void load_vec<__nv_bfloat16, 4>(const __nv_bfloat16* p, float* f) {
  uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
template <typename T>
__device__ __forceinline__ void load_vec8(const T* p, float* f);
template <>
__device__ __forceinline__ void load_vec8<float>(const float* p, float* f) {
  load_vec<float, 4>(p, f);
  load_vec<float, 4>(p + 4, f + 4);
}
template <>
__device__ __forceinline__ void load_vec8<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
template <typename T>
__device__ __forceinline__ void store_vec8(T* p, const float* f);
template <>
__device__ __forceinline__ void store_vec8<float>(float* p, const float* f) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
template <>
__device__ __forceinline__ void store_vec8<__nv_bfloat16>(__nv_bfloat16* p, const float* f) {
  uint4 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = q;
}
template <typename T>
__device__ __forceinline__ void store_vec4(T* p, const float* f);
template <>
__device__ __forceinline__ void store_vec4<float>(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
template <>
__device__ __forceinline__ void store_vec4<__nv_bfloat16>(__nv_bfloat16* p, const float* f) {
  uint2 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
  h[0] = __floats2bfloat162_rn(f[0], f[1]);
  h[1] = __floats2bfloat162_rn(f[2], f[3]);
  *reinterpret_cast<uint2*>(p) = q;
}

struct Tap {
  int y_low, x_low, y_high, x_high;
  float w1, w2, w3, w4;
  bool valid;
};

// torchvision roi_align bilinear_interpolate index/weight computation
__device__ __forceinline__ Tap make_tap(float y, float x, int height, int width) {
  Tap t;
  t.valid = !(y < -1.0f || y > (float)height || x < -1.0f || x > (float)width);
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  t.y_low = (int)y;
  t.x_low = (int)x;
  if (t.y_low >= height - 1) { t.y_high = t.y_low = height - 1; y = (float)t.y_low; } else { t.y_high = t.y_low + 1; }
  if (t.x_low >= width - 1) { t.x_high = t.x_low = width - 1; x = (float)t.x_low; } else { t.x_high = t.x_low + 1; }
  const float ly = y - t.y_low, lx = x - t.x_low, hy = 1.f - ly, hx = 1.f - lx;
  t.w1 = hy * hx; t.w2 = hy * lx; t.w3 = ly * hx; t.w4 = ly * lx;
  return t;
}

// block = 256 threads = 8 warps; warp -> (roi, bin); lanes -> channel groups of 4 (C/4 <= 64 groups -> 2 per lane max)
template <typename T, bool BWD>
__global__ void __launch_bounds__(256) roi_align_kernel(const RoiArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int bins = a.pooled * a.pooled;
  const long long total = (long long)a.m * bins;
  const int nvalid = a.num_valid ? *a.num_valid : a.m;
  for (long long wid = warp_global; wid < total; wid += (long long)gridDim.x * 8) {
    const int roi = (int)(wid / bins);
    const int bin = (int)(wid - (long long)roi * bins);
    const int ph = bin / a.pooled, pw = bin - ph * a.pooled;
    const size_t obase = ((size_t)roi * bins + bin) * a.c;
    if (roi >= nvalid) {
      if (!BWD) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        for (int cg = lane; cg * 4 < a.c; cg += 32) store_vec4<T>(reinterpret_cast<T*>(a.out) + obase + cg * 4, z);
      }
      continue;
    }
    const float* box = a.rois + (size_t)roi * 4;
    const int lvl = roi_level(box, a);
    const int height = a.fh[lvl], width = a.fw[lvl];
    const float sc = a.scale[lvl];
    const int b = a.roi_batch[roi];
    const float roi_start_w = box[0] * sc - 0.5f, roi_start_h = box[1] * sc - 0.5f;
    const float roi_end_w = box[2] * sc - 0.5f, roi_end_h = box[3] * sc - 0.5f;
    const float roi_width = roi_end_w - roi_start_w, roi_height = roi_end_h - roi_start_h;
    const float bin_h = roi_height / (float)a.pooled, bin_w = roi_width / (float)a.pooled;
    const int grid_h = (int)ceilf(roi_height / (float)a.pooled);
    const int grid_w = (int)ceilf(roi_width / (float)a.pooled);
    const float count = fmaxf((float)(grid_h * grid_w), 1.f);
    const size_t fbase = (size_t)b * height * width * a.c;

    for (int cg = lane; cg * 4 < a.c; cg += 32) {
      const int c0 = cg * 4;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      float g[4];
      if (BWD) {
        load_vec<T, 4>(reinterpret_cast<const T*>(a.dout) + obase + c0, g);
#pragma unroll
        for (int k = 0; k < 4; ++k) g[k] /= count;
      }
      for (int iy = 0; iy < grid_h; ++iy) {
        const float y = roi_start_h + ph * bin_h + ((float)iy + .5f) * bin_h / (float)grid_h;
        for (int ix = 0; ix < grid_w; ++ix) {
          const float x = roi_start_w + pw * bin_w + ((float)ix + .5f) * bin_w / (float)grid_w;
          const Tap t = make_tap(y, x, height, width);
          if (!t.valid) continue;
          const size_t o1 = fbase + ((size_t)t.y_low * width + t.x_low) * a.c + c0;
          const size_t o2 = fbase + ((size_t)t.y_low * width + t.x_high) * a.c + c0;
          const size_t o3 = fbase + ((size_t)t.y_high * width + t.x_low) * a.c + c0;
          const size_t o4 = fbase + ((size_t)t.y_high * width + t.x_high) * a.c + c0;
          if (!BWD) {
            const T* f = reinterpret_cast<const T*>(a.feat[lvl]);
            float v1[4], v2[4], v3[4], v4[4];
            load_vec<T, 4>(f + o1, v1);
            load_vec<T, 4>(f + o2, v2);
            load_vec<T, 4>(f + o3, v3);
            load_vec<T, 4>(f + o4, v4);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] += t.w1 * v1[k] + t.w2 * v2[k] + t.w3 * v3[k] + t.w4 * v4[k];
          } else {
            float* d = a.dfeat[lvl];
            atomicAdd(reinterpret_cast<float4*>(d + o1), make_float4(g[0] * t.w1, g[1] * t.w1, g[2] * t.w1, g[3] * t.w1));
            atomicAdd(reinterpret_cast<float4*>(d + o2), make_float4(g[0] * t.w2, g[1] * t.w2, g[2] * t.w2, g[3] * t.w2));
            atomicAdd(reinterpret_cast<float4*>(d + o3), make_float4(g[0] * t.w3, g[1] * t.w3, g[2] * t.w3, g[3] * t.w3));
            atomicAdd(reinterpret_cast<float4*>(d + o4), make_float4(g[0] * t.w4, g[1] * t.w4, g[2] * t.w4, g[3] * t.w4));
          }
        }
      }
      if (!BWD) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] /= count;
        store_vec4<T>(reinterpret_cast<T*>(a.out) + obase + c0, acc);
      }
    }
  }
}

// Forward, channels % 8 == 0: one warp per (RoI, bin), 8 channels (16 B of bf16) per lane, so a 256-channel pixel is
// ONE 512-byte warp load per bilinear tap.  Same per-channel arithmetic order as the generic kernel above.
template <typename T>
__global__ void __launch_bounds__(256) roi_align_fwd8_kernel(const RoiArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int bins = a.pooled * a.pooled;
  const long long total = (long long)a.m * bins;
  const int nvalid = a.num_valid ? *a.num_valid : a.m;
  for (long long wid = warp_global; wid < total; wid += (long long)gridDim.x * 8) {
    const int roi = (int)(wid / bins);
    const int bin = (int)(wid - (long long)roi * bins);
    const int ph = bin / a.pooled, pw = bin - ph * a.pooled;
    const size_t obase = ((size_t)roi * bins + bin) * a.c;
    if (roi >= nvalid) {
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int c0 = lane * 8; c0 < a.c; c0 += 256) store_vec8<T>(reinterpret_cast<T*>(a.out) + obase + c0, z);
      continue;
    }
    const float* box = a.rois + (size_t)roi * 4;
    const int lvl = roi_level(box, a);
    const int height = a.fh[lvl], width = a.fw[lvl];
    const float sc = a.scale[lvl];
    const int b = a.roi_batch[roi];
    const float roi_start_w = box[0] * sc - 0.5f, roi_start_h = box[1] * sc - 0.5f;
    const float roi_end_w = box[2] * sc - 0.5f, roi_end_h = box[3] * sc - 0.5f;
    const float roi_width = roi_end_w - roi_start_w, roi_height = roi_end_h - roi_start_h;
    const float bin_h = roi_height / (float)a.pooled, bin_w = roi_width / (float)a.pooled;
    const int grid_h = (int)ceilf(roi_height / (float)a.pooled);
    const int grid_w = (int)ceilf(roi_width / (float)a.pooled);
    const float count = fmaxf((float)(grid_h * grid_w), 1.f);
    const T* f = reinterpret_cast<const T*>(a.feat[lvl]) + (size_t)b * height * width * a.c;
    for (int c0 = lane * 8; c0 < a.c; c0 += 256) {
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
      for (int iy = 0; iy < grid_h; ++iy) {
        const float y = roi_start_h + ph * bin_h + ((float)iy + .5f) * bin_h / (float)grid_h;
        for (int ix = 0; ix < grid_w; ++ix) {
          const float x = roi_start_w + pw * bin_w + ((float)ix + .5f) * bin_w / (float)grid_w;
          const Tap t = make_tap(y, x, height, width);
          if (!t.valid) continue;
          float v1[8], v2[8], v3[8], v4[8];
          load_vec8<T>(f + ((size_t)t.y_low * width + t.x_low) * a.c + c0, v1);
          load_vec8<T>(f + ((size_t)t.y_low * width + t.x_high) * a.c + c0, v2);
          load_vec8<T>(f + ((size_t)t.y_high * width + t.x_low) * a.c + c0, v3);
          load_vec8<T>(f + ((size_t)t.y_high * width + t.x_high) * a.c + c0, v4);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += t.w1 * v1[k] + t.w2 * v2[k] + t.w3 * v3[k] + t.w4 * v4[k];
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] /= count;
      store_vec8<T>(reinterpret_cast<T*>(a.out) + obase + c0, acc);
    }
  }
}

// Backward, pixel-centric.  Bilinear sampling + bin averaging is separable:
//   d feat[y, x, c] = sum_{ph, pw} Wy[ph][y] * Wx[pw][x] * dout[ph, pw, c] / count
// with Wy[ph][y] = sum over the bin's valid y samples of the bilinear weight landing on row y (likewise Wx).  One
// block per RoI: the 49 x C output gradient is staged in shared memory (pre-divided by count), the two sparse
// weight tables are built there, then each warp takes one footprint pixel at a time and issues ONE 128-bit
// reduction per 4 channels per UNIQUE pixel — the sample-centric scatter issues one per sample tap, about 4x more.
constexpr int kMaxAxis = 1024;  // longest footprint along one axis (feature-map extent)

template <typename T>
__global__ void __launch_bounds__(256) roi_align_bwd_pixel_kernel(const RoiArgs a) {
  extern __shared__ float s_dyn[];
  pdl_wait();
  pdl_launch_dependents();
  const int P = a.pooled, bins = P * P;
  float* s_g = s_dyn;                       // [bins][C]
  float* s_wy = s_g + (size_t)bins * a.c;   // [P][len_y]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvalid = a.num_valid ? min(*a.num_valid, a.m) : a.m;
  for (int roi = blockIdx.x; roi < nvalid; roi += gridDim.x) {
    const float* box = a.rois + (size_t)roi * 4;
    const int lvl = roi_level(box, a);
    const int height = a.fh[lvl], width = a.fw[lvl];
    const float sc = a.scale[lvl];
    const int b = a.roi_batch[roi];
    const float roi_start_w = box[0] * sc - 0.5f, roi_start_h = box[1] * sc - 0.5f;
    const float roi_end_w = box[2] * sc - 0.5f, roi_end_h = box[3] * sc - 0.5f;
    const float roi_width = roi_end_w - roi_start_w, roi_height = roi_end_h - roi_start_h;
    const float bin_h = roi_height / (float)P, bin_w = roi_width / (float)P;
    const int grid_h = (int)ceilf(roi_height / (float)P);
    const int grid_w = (int)ceilf(roi_width / (float)P);
    if (grid_h <= 0 || grid_w <= 0) continue;   // no samples: zero gradient (block-uniform)
    const float inv_count = 1.f / (float)(grid_h * grid_w);
    // footprint: sample coordinates grow monotonically with (bin, sample)
    const float y_first = roi_start_h + .5f * bin_h / (float)grid_h;
    const float y_last = roi_start_h + (P - 1) * bin_h + ((float)(grid_h - 1) + .5f) * bin_h / (float)grid_h;
    const float x_first = roi_start_w + .5f * bin_w / (float)grid_w;
    const float x_last = roi_start_w + (P - 1) * bin_w + ((float)(grid_w - 1) + .5f) * bin_w / (float)grid_w;
    const int ymin = y_first <= 0.f ? 0 : min((int)y_first, height - 1);
    const int ymax = y_last <= 0.f ? 0 : min((int)y_last + 1, height - 1);
    const int xmin = x_first <= 0.f ? 0 : min((int)x_first, width - 1);
    const int xmax = x_last <= 0.f ? 0 : min((int)x_last + 1, width - 1);
    const int len_y = ymax - ymin + 1, len_x = xmax - xmin + 1;
    float* s_wx = s_wy + (size_t)P * len_y;
    __syncthreads();  // previous RoI's tables / gradient tile are no longer read
    for (int i = threadIdx.x; i < P * (len_y + len_x); i += 256) s_wy[i] = 0.f;
    {
      const T* g = reinterpret_cast<const T*>(a.dout) + (size_t)roi * bins * a.c;
      for (int i = threadIdx.x * 8; i < bins * a.c; i += 256 * 8) {
        float f[8];
        load_vec8<T>(g + i, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) s_g[i + k] = f[k] * inv_count;
      }
    }
    __syncthreads();
    // sparse per-axis weights (torchvision bilinear_interpolate index / weight rules, validity per axis)
    for (int i = threadIdx.x; i < P * grid_h; i += 256) {
      const int ph = i / grid_h, iy = i - ph * grid_h;
      float y = roi_start_h + ph * bin_h + ((float)iy + .5f) * bin_h / (float)grid_h;
      if (y < -1.0f || y > (float)height) continue;
      if (y <= 0.f) y = 0.f;
      int lo = (int)y, hi;
      if (lo >= height - 1) { hi = lo = height - 1; y = (float)lo; } else { hi = lo + 1; }
      const float l = y - lo, h = 1.f - l;
      if (lo >= ymin && lo <= ymax) atomicAdd(&s_wy[ph * len_y + lo - ymin], h);
      if (hi >= ymin && hi <= ymax) atomicAdd(&s_wy[ph * len_y + hi - ymin], l);
    }
    for (int i = threadIdx.x; i < P * grid_w; i += 256) {
      const int pw = i / grid_w, ix = i - pw * grid_w;
      float x = roi_start_w + pw * bin_w + ((float)ix + .5f) * bin_w / (float)grid_w;
      if (x < -1.0f || x > (float)width) continue;
      if (x <= 0.f) x = 0.f;
      int lo = (int)x, hi;
      if (lo >= width - 1) { hi = lo = width - 1; x = (float)lo; } else { hi = lo + 1; }
      const float l = x - lo, h = 1.f - l;
      if (lo >= xmin && lo <= xmax) atomicAdd(&s_wx[pw * len_x + lo - xmin], h);
      if (hi >= xmin && hi <= xmax) atomicAdd(&s_wx[pw * len_x + hi - xmin], l);
    }
    __syncthreads();
    float* d = a.dfeat[lvl] + (size_t)b * height * width * a.c;
    for (int pix = warp; pix < len_y * len_x; pix += 8) {
      const int yy = pix / len_x, xx = pix - yy * len_x;
      float wy[8], wx[8];
      bool any_y = false, any_x = false;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        wy[p] = p < P ? s_wy[p * len_y + yy] : 0.f;
        wx[p] = p < P ? s_wx[p * len_x + xx] : 0.f;
        any_y |= wy[p] != 0.f;
        any_x |= wx[p] != 0.f;
      }
      if (!(any_y && any_x)) continue;
      for (int c0 = lane * 8; c0 < a.c; c0 += 256) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
        for (int ph = 0; ph < 8; ++ph) {
          if (ph >= P || wy[ph] == 0.f) continue;
#pragma unroll
          for (int pw = 0; pw < 8; ++pw) {
            if (pw >= P || wx[pw] == 0.f) continue;
            const float w = wy[ph] * wx[pw];
            const float4 g0 = *reinterpret_cast<const float4*>(s_g + (size_t)(ph * P + pw) * a.c + c0);
            const float4 g1 = *reinterpret_cast<const float4*>(s_g + (size_t)(ph * P + pw) * a.c + c0 + 4);
            acc[0] += w * g0.x; acc[1] += w * g0.y; acc[2] += w * g0.z; acc[3] += w * g0.w;
            acc[4] += w * g1.x; acc[5] += w * g1.y; acc[6] += w * g1.z; acc[7] += w * g1.w;
          }
        }
        float* dp = d + ((size_t)(ymin + yy) * width + (xmin + xx)) * a.c + c0;
        atomicAdd(reinterpret_cast<float4*>(dp), make_float4(acc[0], acc[1], acc[2], acc[3]));
        atomicAdd(reinterpret_cast<float4*>(dp + 4), make_float4(acc[4], acc[5], acc[6], acc[7]));
      }
    }
  }
}

int fill_args(const aldi_roialign_params* p, RoiArgs* a, const char* who) {
  ALDI_CHECK_ARG(p && p->rois && p->roi_batch, "%s: null pointer", who);
  ALDI_CHECK_ARG(p->channels > 0 && p->channels % 4 == 0, "%s: channels must be a multiple of 4", who);
  ALDI_CHECK_ARG(p->num_levels >= 1 && p->num_levels <= 4, "%s: 1..4 levels supported", who);
  for (int i = 0; i < 4; ++i) {
    int j = i < p->num_levels ? i : p->num_levels - 1;
    a->feat[i] = p->feat[j];
    a->dfeat[i] = p->dfeat[j];
    a->fh[i] = p->feat_h[j];
    a->fw[i] = p->feat_w[j];
    a->scale[i] = p->scale[j];
  }
  a->c = p->channels;
  a->pooled = p->pooled;
  a->rois = p->rois;
  a->roi_batch = p->roi_batch;
  a->num_valid = p->num_valid;
  a->m = p->num_rois;
  a->out = p->out;
  a->dout = p->dout;
  a->min_level = p->min_level;
  a->max_level = p->min_level + p->num_levels - 1;
  a->canonical_size = p->canonical_box_size;
  a->canonical_level = p->canonical_level;
  return ALDI_OK;
}

}  // namespace

extern "C" int aldi_roi_align_forward(const aldi_roialign_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RoiArgs a;
  int rc = fill_args(p, &a, "aldi_roi_align_forward");
  if (rc) return rc;
  ALDI_CHECK_ARG(p->out, "aldi_roi_align_forward: null out");
  if (p->num_rois == 0) return ALDI_OK;
  long long warps = (long long)p->num_rois * p->pooled * p->pooled;
  long long blocks = (warps + 7) / 8;
  long long cap = (long long)aldi_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  const bool vec8 = p->channels % 8 == 0;
  for (int i = 0; vec8 && i < p->num_levels; ++i)
    if (reinterpret_cast<uintptr_t>(p->feat[i]) & 15) return (aldi_set_error("aldi_roi_align_forward: feat[%d] not 16-byte aligned", i), ALDI_ERR_INVALID);
  if (vec8 && p->dtype == ALDI_DTYPE_BF16)
    aldi_launch_pdl(roi_align_fwd8_kernel<__nv_bfloat16>, dim3((unsigned)blocks), dim3(256), 0, stream, a);
  else if (vec8)
    aldi_launch_pdl(roi_align_fwd8_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, stream, a);
  else if (p->dtype == ALDI_DTYPE_BF16)
    roi_align_kernel<__nv_bfloat16, false><<<(unsigned)blocks, 256, 0, stream>>>(a);
  else
    roi_align_kernel<float, false><<<(unsigned)blocks, 256, 0, stream>>>(a);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_roi_align_forward");
  return ALDI_OK;
}

extern "C" int aldi_roi_align_backward(const aldi_roialign_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RoiArgs a;
  int rc = fill_args(p, &a, "aldi_roi_align_backward");
  if (rc) return rc;
  ALDI_CHECK_ARG(p->dout, "aldi_roi_align_backward: null dout");
  for (int i = 0; i < p->num_levels; ++i)
    ALDI_CHECK_ARG(p->dfeat[i] && (reinterpret_cast<uintptr_t>(p->dfeat[i]) & 15) == 0,
                   "aldi_roi_align_backward: dfeat[%d] null or not 16-byte aligned", i);
  if (p->num_rois == 0) return ALDI_OK;
  int max_h = 0, max_w = 0;
  for (int i = 0; i < p->num_levels; ++i) {
    if (p->feat_h[i] > max_h) max_h = p->feat_h[i];
    if (p->feat_w[i] > max_w) max_w = p->feat_w[i];
  }
  static const char* legacy = getenv("ALDI_ROI_BWD_LEGACY");  // A/B knob: the sample-centric scatter
  if (!legacy && p->channels % 8 == 0 && p->pooled <= 8 && max_h <= kMaxAxis && max_w <= kMaxAxis) {
    const size_t smem = ((size_t)p->pooled * p->pooled * p->channels + (size_t)p->pooled * (max_h + max_w)) * sizeof(float);
    if (smem <= 200 * 1024) {
      auto kern = p->dtype == ALDI_DTYPE_BF16 ? roi_align_bwd_pixel_kernel<__nv_bfloat16> : roi_align_bwd_pixel_kernel<float>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) {
        aldi_set_error("aldi_roi_align_backward: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return ALDI_ERR_CUDA;
      }
      int per_sm = (int)((220 * 1024) / (smem + 1024));
      if (per_sm > 8) per_sm = 8;
      if (per_sm < 1) per_sm = 1;
      long long blocks = (long long)aldi_num_sms() * per_sm;
      if (blocks > p->num_rois) blocks = p->num_rois;
      aldi_launch_pdl(kern, dim3((unsigned)blocks), dim3(256), smem, stream, a);
      ALDI_COUNT_LAUNCH();
      ALDI_CUDA_LAUNCH_CHECK("aldi_roi_align_backward");
      return ALDI_OK;
    }
  }
  long long warps = (long long)p->num_rois * p->pooled * p->pooled;
  long long blocks = (warps + 7) / 8;
  long long cap = (long long)aldi_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (p->dtype == ALDI_DTYPE_BF16)
    roi_align_kernel<__nv_bfloat16, true><<<(unsigned)blocks, 256, 0, stream>>>(a);
  else
    roi_align_kernel<float, true><<<(unsigned)blocks, 256, 0, stream>>>(a);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_roi_align_backward");
  return ALDI_OK;
}
