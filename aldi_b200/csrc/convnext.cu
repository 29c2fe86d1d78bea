// ConvNeXt building blocks on channels-last activations (BASELINE configs[4]): everything of
// aldi/backbone.py:189-338 that is not a dense GEMM (those run on conv_tc / wgrad_tc as 1x1 layers).
//
//   aldi_layernorm_*        LayerNorm over the channel axis per pixel: both the channels_last F.layer_norm and the
//                           hand-written channels_first variant of aldi/backbone.py:321-346 are this on NHWC
//   aldi_dwconv7_*          depthwise 7x7, padding 3 (ConvNextBlock.dwconv, :205): forward, data gradient (the same
//                           stencil with the taps reversed) and weight gradient
//   aldi_gelu_*             exact (erf) GELU between the two pointwise layers (:208) and its derivative
//   aldi_layerscale_*       x = input + drop_path(gamma * u) (:222-227) and its gradients (d u, d gamma)
//   aldi_space_to_depth / aldi_depth_to_space   the 2x2 / 4x4 stride-k "patchify" convolutions (:249-258) become
//                           1x1 GEMMs over rows of k*k*C values; the inverse scatters the data gradient back
//   aldi_patchify_image     uint8 NCHW image -> normalised 4x4x3 patches (48 of 64 channels), the stem's GEMM operand
//   aldi_adamw_step         torch.optim.AdamW over the flat buffers (aldi/trainer.py:205-206)
// All HBM-bound streams; fp32 arithmetic, activation dtype T in {float (parity mode), bf16}.
#include <algorithm>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "sm100.cuh"
#include "tmap.h"
#include "../../include/aldi_b200.h"

using namespace sm100;

#define ALDI_CUDA_CHECK(expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      aldi_set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                \
      return ALDI_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

namespace {

template <typename T> struct V8;   // 8 elements per thread
template <> struct V8<float> {
  static __device__ __forceinline__ void load(const float* p, float* f) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
};
template <> struct V8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* f) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* f) {
    uint4 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = q;
  }
};

int blocks_for(long long work, int per_block, int per_sm) {
  long long b = (work + per_block - 1) / per_block;
  const long long cap = (long long)aldi_num_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row (pixel); stats[row] = (mean, rstd)
template <typename T>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
              long long rows, int c, int stride, T* __restrict__ y, float2* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, wstep = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = w0; r < rows; r += wstep) {
    const T* xr = x + r * stride;
    float s = 0.f;
    for (int i = lane; i < c; i += 32) s += to_f32<T>(xr[i]);
    const float mean = warp_sum(s) / (float)c;
    float v = 0.f;
    for (int i = lane; i < c; i += 32) { const float d = to_f32<T>(xr[i]) - mean; v += d * d; }
    const float rstd = rsqrtf(warp_sum(v) / (float)c + eps);
    T* yr = y + r * stride;
    for (int i = lane; i < c; i += 32) yr[i] = from_f32<T>((to_f32<T>(xr[i]) - mean) * rstd * gamma[i] + beta[i]);
    for (int i = c + lane; i < stride; i += 32) yr[i] = from_f32<T>(0.f);
    if (lane == 0 && stats) stats[r] = make_float2(mean, rstd);
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; dgamma += sum dy * xhat; dbeta += sum dy.
// One warp per row; CPL = channels per lane (compile time, so the per-lane d-gamma / d-beta partials live in registers);
// a block's partials meet in shared memory, then one atomic per channel per block.
template <typename T, int CPL>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float2* __restrict__ stats,
              const T* __restrict__ dy, long long rows, int c, int stride, T* __restrict__ dx, int accumulate,
              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float s_red[];   // [2][c]
  const int lane = threadIdx.x & 31;
  const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, wstep = ((long long)gridDim.x * blockDim.x) >> 5;
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  float pg[CPL], pb[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) pg[k] = pb[k] = 0.f;
  for (long long r = w0; r < rows; r += wstep) {
    const T* xr = x + r * stride;
    const T* dr = dy + r * stride;
    const float2 st = stats[r];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int i = lane + 32 * k;
      if (i < c) {
        const float g = to_f32<T>(dr[i]) * __ldg(gamma + i);
        sg += g;
        sgx += g * (to_f32<T>(xr[i]) - st.x) * st.y;
      }
    }
    const float mg = warp_sum(sg) / (float)c, mgx = warp_sum(sgx) / (float)c;
    T* dxr = dx + r * stride;
    // second sweep re-reads the row (3 KB, L1-resident) instead of holding it in registers next to the partials
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int i = lane + 32 * k;
      if (i < c) {
        const float d = to_f32<T>(dr[i]), xh = (to_f32<T>(xr[i]) - st.x) * st.y;
        float v = st.y * (d * __ldg(gamma + i) - mg - xh * mgx);
        if (accumulate) v += to_f32<T>(dxr[i]);
        dxr[i] = from_f32<T>(v);
        pg[k] += d * xh;
        pb[k] += d;
      }
    }
    if (!accumulate)
      for (int i = c + lane; i < stride; i += 32) dxr[i] = from_f32<T>(0.f);
  }
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int i = lane + 32 * k;
    if (i < c) { atomicAdd(&s_red[i], pg[k]); atomicAdd(&s_red[c + i], pb[k]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    if (s_red[i] != 0.f) atomicAdd(dgamma + i, s_red[i]);
    if (s_red[c + i] != 0.f) atomicAdd(dbeta + i, s_red[c + i]);
  }
}

// Vector variants (c and stride multiples of 8, c <= 2048): a lane owns 8-channel vectors j = lane + 32 k, moved as one
// 16-byte (bf16) access instead of eight 2-byte ones; the forward keeps the row in registers (one read, one write) and
// normalises RPW rows per warp iteration so that narrow rows still keep several loads in flight.
template <typename T, int VPL, int RPW>
__global__ void __launch_bounds__(256, 2)
ln_fwd_vec_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                  long long rows, int c, int stride, T* __restrict__ y, float2* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int nvec = c >> 3, svec = stride >> 3;
  const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, wstep = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r0 = w0 * RPW; r0 < rows; r0 += wstep * RPW) {
    float v[RPW][VPL][8];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int j = lane + 32 * k;
        if (j < nvec && r0 + rr < rows) V8<T>::load(x + (r0 + rr) * stride + 8 * j, v[rr][k]);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[rr][k][e] = 0.f;
        }
      }
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
      const long long r = r0 + rr;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[rr][k][e];
      const float mean = warp_sum(s) / (float)c;
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k)
        if (lane + 32 * k < nvec) {
#pragma unroll
          for (int e = 0; e < 8; ++e) { const float d = v[rr][k][e] - mean; q += d * d; }
        }
      const float rstd = rsqrtf(warp_sum(q) / (float)c + eps);
      if (r >= rows) continue;                            // warp-uniform
      T* yr = y + r * stride;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int j = lane + 32 * k;
        if (j < nvec) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + 8 * j)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + 8 * j + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + 8 * j)), b1 = __ldg(reinterpret_cast<const float4*>(beta + 8 * j + 4));
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = (v[rr][k][e] - mean) * rstd * gg[e] + bb[e];
          V8<T>::store(yr + 8 * j, o);
        }
      }
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int j = nvec + lane; j < svec; j += 32) V8<T>::store(yr + 8 * j, z);
      if (lane == 0 && stats) stats[r] = make_float2(mean, rstd);
    }
  }
}

// backward: d-gamma / d-beta partials of the lane's own channels stay in registers across all its rows.  Rows of up
// to 1024 channels are held in registers as well (RPW rows per warp iteration: one read of x and dy); wider rows take
// a second sweep that hits L1, as in the scalar kernel.
template <typename T, int VPL, int RPW>
__global__ void __launch_bounds__(256, VPL <= 3 ? 2 : 1)
ln_bwd_vec_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float2* __restrict__ stats,
                  const T* __restrict__ dy, long long rows, int c, int stride, T* __restrict__ dx, int accumulate,
                  float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float s_red[];   // [2][c]
  constexpr bool KEEP = VPL * RPW <= 4;
  static_assert(KEEP || RPW == 1, "wide rows: one row per iteration");
  const int lane = threadIdx.x & 31;
  const int nvec = c >> 3, svec = stride >> 3;
  const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, wstep = ((long long)gridDim.x * blockDim.x) >> 5;
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  float pg[VPL][8], pb[VPL][8];
#pragma unroll
  for (int k = 0; k < VPL; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) pg[k][e] = pb[k][e] = 0.f;
  const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long r0 = w0 * RPW; r0 < rows; r0 += wstep * RPW) {
    float xv[KEEP ? RPW : 1][KEEP ? VPL : 1][8], dv[KEEP ? RPW : 1][KEEP ? VPL : 1][8];
    if constexpr (KEEP) {
#pragma unroll
      for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const int j = lane + 32 * k;
          if (j < nvec && r0 + rr < rows) {
            V8<T>::load(x + (r0 + rr) * stride + 8 * j, xv[rr][k]);
            V8<T>::load(dy + (r0 + rr) * stride + 8 * j, dv[rr][k]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) xv[rr][k][e] = dv[rr][k][e] = 0.f;
          }
        }
    }
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
      const long long r = r0 + rr;
      const bool live = r < rows;                         // warp-uniform
      const T* xr = x + r * stride;
      const T* dr = dy + r * stride;
      const float2 st = live ? stats[r] : make_float2(0.f, 0.f);
      float sg = 0.f, sgx = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int j = lane + 32 * k;
        if (j < nvec && live) {
          float xs[8], ds[8];
          if constexpr (!KEEP) {
            V8<T>::load(xr + 8 * j, xs);
            V8<T>::load(dr + 8 * j, ds);
          }
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + 8 * j)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + 8 * j + 4));
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float xe = KEEP ? xv[KEEP ? rr : 0][KEEP ? k : 0][e] : xs[e], de = KEEP ? dv[KEEP ? rr : 0][KEEP ? k : 0][e] : ds[e];
            const float g = de * gg[e];
            sg += g;
            sgx += g * (xe - st.x) * st.y;
          }
        }
      }
      const float mg = warp_sum(sg) / (float)c, mgx = warp_sum(sgx) / (float)c;
      if (!live) continue;
      T* dxr = dx + r * stride;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int j = lane + 32 * k;
        if (j < nvec) {
          float xs[8], ds[8], o[8];
          if constexpr (!KEEP) {
            V8<T>::load(xr + 8 * j, xs);
            V8<T>::load(dr + 8 * j, ds);
          }
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + 8 * j)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + 8 * j + 4));
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          if (accumulate) V8<T>::load(dxr + 8 * j, o);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float xe = KEEP ? xv[KEEP ? rr : 0][KEEP ? k : 0][e] : xs[e], de = KEEP ? dv[KEEP ? rr : 0][KEEP ? k : 0][e] : ds[e];
            const float xh = (xe - st.x) * st.y;
            const float val = st.y * (de * gg[e] - mg - xh * mgx);
            o[e] = accumulate ? o[e] + val : val;
            pg[k][e] += de * xh;
            pb[k][e] += de;
          }
          V8<T>::store(dxr + 8 * j, o);
        }
      }
      if (!accumulate)
        for (int j = nvec + lane; j < svec; j += 32) V8<T>::store(dxr + 8 * j, z);
    }
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int j = lane + 32 * k;
    if (j < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { atomicAdd(&s_red[8 * j + e], pg[k][e]); atomicAdd(&s_red[c + 8 * j + e], pb[k][e]); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    if (s_red[i] != 0.f) atomicAdd(dgamma + i, s_red[i]);
    if (s_red[c + i] != 0.f) atomicAdd(dbeta + i, s_red[c + i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// depthwise 7x7, padding 3.  block = 256 threads: 8 channel groups of 8 (64 channels, blockIdx.y) x 32 pixels.
// FLIP: data gradient (tap (6-r, 6-s)); weights [c][49] fp32.
template <typename T, bool FLIP>
__global__ void __launch_bounds__(256)
dw7_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int n, int h, int wd, int c,
           int stride, T* __restrict__ y, int accumulate) {
  // each thread: 8 channels x 4 consecutive output pixels of one row.  Per filter row it loads the 10 input columns
  // the 4 outputs share and the row's 7 x 8 weights once: 2.8x fewer global loads and 4x fewer shared-memory reads
  // than one pixel per thread.
  __shared__ float s_w[49][64];
  const int c0 = blockIdx.y * 64;
  for (int i = threadIdx.x; i < 49 * 64; i += 256) {
    const int t = i / 64, cc = i % 64;
    s_w[t][cc] = (c0 + cc < c) ? w[(size_t)(c0 + cc) * 49 + (FLIP ? 48 - t : t)] : 0.f;
  }
  __syncthreads();
  const int cg = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int ch = c0 + cg * 8;
  if (ch >= stride) return;
  const int wq = (wd + 3) >> 2;                       // 4-pixel groups per row
  const long long total = (long long)n * h * wq;
  for (long long p = (long long)blockIdx.x * 32 + pl; p < total; p += (long long)gridDim.x * 32) {
    const int xq = (int)(p % wq);
    const long long q = p / wq;
    const int py = (int)(q % h), img = (int)(q / h);
    const int px0 = xq * 4;
    float acc[4][8];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[o][k] = (bias && ch + k < c) ? bias[ch + k] : 0.f;
    for (int r = 0; r < 7; ++r) {
      const int yy = py + r - 3;
      if (yy < 0 || yy >= h) continue;
      float wr[7][8];
#pragma unroll
      for (int s7 = 0; s7 < 7; ++s7) {
        const float4 w0 = *reinterpret_cast<const float4*>(&s_w[r * 7 + s7][cg * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&s_w[r * 7 + s7][cg * 8 + 4]);
        wr[s7][0] = w0.x; wr[s7][1] = w0.y; wr[s7][2] = w0.z; wr[s7][3] = w0.w;
        wr[s7][4] = w1.x; wr[s7][5] = w1.y; wr[s7][6] = w1.z; wr[s7][7] = w1.w;
      }
      const T* row = x + (((long long)img * h + yy) * wd) * stride + ch;
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        const int xx = px0 - 3 + j;
        if (xx < 0 || xx >= wd) continue;
        float f[8];
        V8<T>::load(row + (long long)xx * stride, f);
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int s7 = j - o;                       // input column j feeds output o through tap s7
          if (s7 < 0 || s7 > 6) continue;
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[o][k] += f[k] * wr[s7][k];
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      if (px0 + o >= wd) continue;
      T* yp = y + ((((long long)img * h + py) * wd) + px0 + o) * stride + ch;
      if (accumulate) {
        float f[8];
        V8<T>::load(yp, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[o][k] += f[k];
      }
      V8<T>::store(yp, acc[o]);
    }
  }
}

// weight gradient: dw[c][r][s] += sum_pixels dy[p, c] * x[p + (r - 3, s - 3), c].  blockIdx.z = filter row r; a thread owns
// 8 channels and walks a run of 32 consecutive pixels of one image row with a sliding window of 7 x-vectors, so each dy
// and x element is read once per filter row (7x in total instead of 49x).
constexpr int kWgRun = 32;
template <typename T>
__global__ void __launch_bounds__(256)
dw7_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, int n, int h, int wd, int c, int stride,
                 float* __restrict__ dw) {
  __shared__ float s_part[32][64];
  const int c0 = blockIdx.y * 64, r = blockIdx.z;
  const int cg = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int ch = c0 + cg * 8;
  float acc[7][8];
#pragma unroll
  for (int s7 = 0; s7 < 7; ++s7)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[s7][k] = 0.f;
  if (ch < stride) {
    const int runs_per_row = (wd + kWgRun - 1) / kWgRun;
    const long long total = (long long)n * h * runs_per_row;
    for (long long p = (long long)blockIdx.x * 32 + pl; p < total; p += (long long)gridDim.x * 32) {
      const int run = (int)(p % runs_per_row);
      const long long q = p / runs_per_row;
      const int py = (int)(q % h), img = (int)(q / h);
      const int yy = py + r - 3;
      if (yy < 0 || yy >= h) continue;
      const int x0 = run * kWgRun, x1 = min(x0 + kWgRun, wd);
      const T* xrow = x + (((long long)img * h + yy) * wd) * stride + ch;
      const T* drow = dy + (((long long)img * h + py) * wd) * stride + ch;
      float win[7][8];                                  // x at columns px - 3 .. px + 3
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int xx = x0 - 3 + j;
        if (xx >= 0 && xx < wd) V8<T>::load(xrow + (long long)xx * stride, win[j + 1]);
        else {
#pragma unroll
          for (int k = 0; k < 8; ++k) win[j + 1][k] = 0.f;
        }
      }
      for (int px = x0; px < x1; ++px) {
#pragma unroll
        for (int j = 0; j < 6; ++j)
#pragma unroll
          for (int k = 0; k < 8; ++k) win[j][k] = win[j + 1][k];
        const int xx = px + 3;
        if (xx < wd) V8<T>::load(xrow + (long long)xx * stride, win[6]);
        else {
#pragma unroll
          for (int k = 0; k < 8; ++k) win[6][k] = 0.f;
        }
        float d[8];
        V8<T>::load(drow + (long long)px * stride, d);
#pragma unroll
        for (int s7 = 0; s7 < 7; ++s7)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[s7][k] += d[k] * win[s7][k];
      }
    }
  }
  for (int s7 = 0; s7 < 7; ++s7) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) s_part[pl][cg * 8 + k] = acc[s7][k];
    __syncthreads();
    if (threadIdx.x < 64) {
      float v = 0.f;
      for (int i = 0; i < 32; ++i) v += s_part[i][threadIdx.x];
      if (c0 + threadIdx.x < c && v != 0.f) atomicAdd(dw + (size_t)(c0 + threadIdx.x) * 49 + r * 7 + s7, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 depthwise 7x7 on shared-memory tiles (the default for bf16; the streaming kernels above remain for fp32 parity
// mode).  The streaming forward re-reads every input element ~17x through L2 (7 filter rows x overlapping columns, no
// reuse between output rows) and the streaming weight gradient 7x: both are L2-bound at a few per cent of the HBM
// roofline.  Here a CTA stages a (16+6) x (16+6) pixel x 64 channel halo tile once (cp.async, zero-filled outside the
// map, double-buffered against the previous tile's math; ONE TMA box load per tile issued by one thread when the map
// is at least a box wide -- per-thread cp.async had the warps stall on the load/store queue, ncu lg_throttle 1.3 per
// issue -- and per-thread cp.async for smaller maps) and a thread owns TWO channels -- one 32-bit word per pixel,
// so a warp reads one conflict-free 128-byte wavefront per pixel -- with those channels' 49 taps (forward / data
// gradient) or 49 partial sums (weight gradient) held in REGISTERS.  A warp owns two tile rows; a unit of work is
// 2 rows x 8 pixels (forward: 1568 FMAs per 112 shared loads) or 1 row x 8 pixels (weight gradient: 784 per 106).
constexpr int kDwT = 16, kDwHalo = kDwT + 6;
constexpr int kDwXBytes = kDwHalo * kDwHalo * 128, kDwYBytes = kDwT * kDwT * 128;
constexpr int kDwWgScratch = 8 * 7 * 64 * (int)sizeof(float);
constexpr int kDwTail = 16 + 128;                       // two mbarriers + slack to align the tiles to 128 bytes
constexpr int kDwWBytes = 64 * 49 * (int)sizeof(float);   // one channel block's taps, staged coalesced
constexpr int kDwFwdSmem = 2 * kDwXBytes + kDwWBytes + kDwTail;
constexpr int kDwWgSmem = 2 * (kDwXBytes + kDwYBytes) + kDwWgScratch + kDwTail;
__device__ __forceinline__ unsigned char* dw_align128(unsigned char* p) {
  return p + ((128u - (smem_u32(p) & 127u)) & 127u);
}

__device__ __forceinline__ uint32_t dw_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dw_cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dw_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void dw_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float2 dw_unpack(uint32_t v) {      // bf16x2 -> (low channel, high channel)
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}

// rows x cols pixels from (gy0, gx0) of image img, channels [c0, c0 + 64) -> dst + (row * cols + col) * 128 bytes;
// pixels outside the map and channel chunks beyond the row stride arrive as zeros (cp.async src-size 0)
__device__ __forceinline__ void dw_load_tile(const __nv_bfloat16* __restrict__ src, uint32_t dst, int img, int gy0, int gx0,
                                             int rows, int cols, int h, int wd, int c0, int stride) {
  const int chunks = rows * cols * 8;
  for (int i = threadIdx.x; i < chunks; i += 256) {
    const int k = i & 7, px = i >> 3;
    const int row = px / cols, col = px - row * cols;
    const int gy = gy0 + row, gx = gx0 + col;
    const bool ok = gy >= 0 && gy < h && gx >= 0 && gx < wd && c0 + 8 * k < stride;
    const __nv_bfloat16* p = ok ? src + (((long long)img * h + gy) * wd + gx) * stride + c0 + 8 * k : src;
    dw_cp_async16(dst + (uint32_t)i * 16, p, ok ? 16 : 0);
  }
}

struct DwTile { int cb, img, y0, x0; };
__device__ __forceinline__ DwTile dw_tile(long long t, int n, int tiles_y, int tiles_x) {
  DwTile d;
  d.x0 = (int)(t % tiles_x) * kDwT;
  t /= tiles_x;
  d.y0 = (int)(t % tiles_y) * kDwT;
  t /= tiles_y;
  d.img = (int)(t % n);
  d.cb = (int)(t / n);
  return d;
}

template <bool FLIP, bool TMA>
__global__ void __launch_bounds__(256, 1)
dw7_tile_kernel(const __grid_constant__ CUtensorMap tmX, const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                const float* __restrict__ bias, int n, int h, int wd, int c, int stride, __nv_bfloat16* __restrict__ y,
                int accumulate) {
  extern __shared__ unsigned char dw_smem_raw[];
  unsigned char* dw_smem = dw_align128(dw_smem_raw);
  float* s_w = reinterpret_cast<float*>(dw_smem + 2 * kDwXBytes);                  // [64][49]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dw_smem + 2 * kDwXBytes + kDwWBytes);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_x = (wd + kDwT - 1) / kDwT, tiles_y = (h + kDwT - 1) / kDwT;
  const long long total = (long long)((stride + 63) / 64) * n * tiles_y * tiles_x;
  // a CONTIGUOUS range of tiles per CTA (channel block slowest): the taps in registers change at most once or twice
  const long long t0 = total * blockIdx.x / gridDim.x, t1 = total * (blockIdx.x + 1) / gridDim.x;
  if (t0 >= t1) return;
  const uint32_t sbase = smem_u32(dw_smem);
  if (TMA) {
    if (threadIdx.x == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      fence_barrier_init();
    }
    __syncthreads();
  }
  {
    const DwTile d = dw_tile(t0, n, tiles_y, tiles_x);
    if (TMA) {
      if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], kDwXBytes);
        tma_load_4d(dw_smem, &tmX, &bars[0], d.cb * 64, d.x0 - 3, d.y0 - 3, d.img);
      }
    } else {
      dw_load_tile(x, sbase, d.img, d.y0 - 3, d.x0 - 3, kDwHalo, kDwHalo, h, wd, d.cb * 64, stride);
      dw_cp_commit();
    }
  }
  float2 wreg[49];
  float2 breg = make_float2(0.f, 0.f);
  int cur_cb = -1;
  int it = 0;
  for (long long t = t0; t < t1; ++t, ++it) {
    const DwTile d = dw_tile(t, n, tiles_y, tiles_x);
    if (t + 1 < t1) {
      const DwTile e = dw_tile(t + 1, n, tiles_y, tiles_x);
      if (TMA) {
        if (threadIdx.x == 0) {                           // the buffer was released by the barrier that ended trip it - 1
          mbar_expect_tx(&bars[(it + 1) & 1], kDwXBytes);
          tma_load_4d(dw_smem + ((it + 1) & 1) * kDwXBytes, &tmX, &bars[(it + 1) & 1], e.cb * 64, e.x0 - 3, e.y0 - 3, e.img);
        }
      } else {
        dw_load_tile(x, sbase + ((it + 1) & 1) * kDwXBytes, e.img, e.y0 - 3, e.x0 - 3, kDwHalo, kDwHalo, h, wd, e.cb * 64, stride);
      }
    }
    if (!TMA) dw_cp_commit();
    const int ch = d.cb * 64 + 2 * lane;
    if (d.cb != cur_cb) {                                 // block-uniform
      cur_cb = d.cb;
      const int c0 = d.cb * 64;
      for (int i = threadIdx.x; i < 64 * 49; i += 256)    // contiguous in global memory: coalesced
        s_w[i] = (c0 + i / 49 < c) ? w[(size_t)c0 * 49 + i] : 0.f;
      __syncthreads();
#pragma unroll
      for (int tp = 0; tp < 49; ++tp) {
        const int src = FLIP ? 48 - tp : tp;
        wreg[tp] = make_float2(s_w[(2 * lane) * 49 + src], s_w[(2 * lane + 1) * 49 + src]);
      }
      breg.x = (bias && ch < c) ? bias[ch] : 0.f;
      breg.y = (bias && ch + 1 < c) ? bias[ch + 1] : 0.f;
      // (the next write of s_w is behind at least one end-of-trip barrier)
    }
    if (TMA) {
      mbar_wait(&bars[it & 1], (it >> 1) & 1);
    } else {
      dw_cp_wait<1>();
      __syncthreads();
    }
    const uint32_t* tile = reinterpret_cast<const uint32_t*>(dw_smem + (it & 1) * kDwXBytes);
#pragma unroll 1
    for (int u = 0; u < 2; ++u) {
      const int cx = u * 8;
      float2 acc[2][8];
#pragma unroll
      for (int o = 0; o < 2; ++o)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[o][q] = breg;
#pragma unroll
      for (int i = 0; i < 8; ++i) {                       // halo rows 2*warp + i feed output rows 2*warp + {0, 1}
        const uint32_t* rowp = tile + ((2 * warp + i) * kDwHalo + cx) * 32 + lane;
        float2 f[14];
#pragma unroll
        for (int j = 0; j < 14; ++j) f[j] = dw_unpack(rowp[j * 32]);
        // tap-major order: consecutive FMAs go to 16 different accumulators (x 2 channels), never back to back on one
#pragma unroll
        for (int s7 = 0; s7 < 7; ++s7)
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            const int r = i - o;
            if (r < 0 || r > 6) continue;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              acc[o][q].x = fmaf(f[q + s7].x, wreg[r * 7 + s7].x, acc[o][q].x);
              acc[o][q].y = fmaf(f[q + s7].y, wreg[r * 7 + s7].y, acc[o][q].y);
            }
          }
      }
      if (ch < stride) {
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int gy = d.y0 + 2 * warp + o;
          if (gy >= h) continue;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int gx = d.x0 + cx + q;
            if (gx >= wd) continue;
            __nv_bfloat162* yp = reinterpret_cast<__nv_bfloat162*>(y + (((long long)d.img * h + gy) * wd + gx) * stride + ch);
            float2 v = acc[o][q];
            if (accumulate) {
              const float2 old = __bfloat1622float2(*yp);
              v.x += old.x;
              v.y += old.y;
            }
            *yp = __floats2bfloat162_rn(v.x, v.y);
          }
        }
      }
    }
    __syncthreads();                                      // the buffer is refilled by the next iteration's prefetch
  }
  if (!TMA) dw_cp_wait<0>();
}

// weight gradient on the same tiles: x halo tile + dy tile in shared memory, 49 x 2 partial sums per thread in
// registers over a CONTIGUOUS range of tiles; one cross-warp reduction + one atomicAdd per (channel, tap) per channel
// block the range touches
template <bool TMA>
__global__ void __launch_bounds__(256, 1)
dw7_wgrad_tile_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                      const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int n, int h, int wd, int c,
                      int stride, float* __restrict__ dw) {
  extern __shared__ unsigned char dw_smem_raw[];
  unsigned char* dw_smem = dw_align128(dw_smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_x = (wd + kDwT - 1) / kDwT, tiles_y = (h + kDwT - 1) / kDwT;
  const long long total = (long long)((stride + 63) / 64) * n * tiles_y * tiles_x;
  const long long t0 = total * blockIdx.x / gridDim.x, t1 = total * (blockIdx.x + 1) / gridDim.x;
  if (t0 >= t1) return;
  constexpr int kStage = kDwXBytes + kDwYBytes;
  const uint32_t sbase = dw_smem_addr(dw_smem);
  float* scratch = reinterpret_cast<float*>(dw_smem + 2 * kStage);          // [8 warps][7][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dw_smem + 2 * kStage + kDwWgScratch);
  if (TMA) {
    if (threadIdx.x == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      fence_barrier_init();
    }
    __syncthreads();
  }
  {
    const DwTile d = dw_tile(t0, n, tiles_y, tiles_x);
    if (TMA) {
      if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], kStage);
        tma_load_4d(dw_smem, &tmX, &bars[0], d.cb * 64, d.x0 - 3, d.y0 - 3, d.img);
        tma_load_4d(dw_smem + kDwXBytes, &tmDY, &bars[0], d.cb * 64, d.x0, d.y0, d.img);
      }
    } else {
      dw_load_tile(x, sbase, d.img, d.y0 - 3, d.x0 - 3, kDwHalo, kDwHalo, h, wd, d.cb * 64, stride);
      dw_load_tile(dy, sbase + kDwXBytes, d.img, d.y0, d.x0, kDwT, kDwT, h, wd, d.cb * 64, stride);
      dw_cp_commit();
    }
  }
  float2 acc[49];
#pragma unroll
  for (int tp = 0; tp < 49; ++tp) acc[tp] = make_float2(0.f, 0.f);
  int cur_cb = dw_tile(t0, n, tiles_y, tiles_x).cb;

  auto flush = [&](int cb) {
    const int c0 = cb * 64;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      __syncthreads();
#pragma unroll
      for (int s7 = 0; s7 < 7; ++s7)
        *reinterpret_cast<float2*>(scratch + (warp * 7 + s7) * 64 + 2 * lane) = acc[r * 7 + s7];
      __syncthreads();
      for (int i = threadIdx.x; i < 7 * 64; i += 256) {
        const int s7 = i >> 6, cc = i & 63;
        float v = 0.f;
#pragma unroll
        for (int wp = 0; wp < 8; ++wp) v += scratch[(wp * 7 + s7) * 64 + cc];
        if (c0 + cc < c && v != 0.f) atomicAdd(dw + (size_t)(c0 + cc) * 49 + r * 7 + s7, v);
      }
    }
#pragma unroll
    for (int tp = 0; tp < 49; ++tp) acc[tp] = make_float2(0.f, 0.f);
  };

  int it = 0;
  for (long long t = t0; t < t1; ++t, ++it) {
    const DwTile d = dw_tile(t, n, tiles_y, tiles_x);
    if (t + 1 < t1) {
      const DwTile e = dw_tile(t + 1, n, tiles_y, tiles_x);
      if (TMA) {
        if (threadIdx.x == 0) {
          unsigned char* dstp = dw_smem + ((it + 1) & 1) * kStage;
          mbar_expect_tx(&bars[(it + 1) & 1], kStage);
          tma_load_4d(dstp, &tmX, &bars[(it + 1) & 1], e.cb * 64, e.x0 - 3, e.y0 - 3, e.img);
          tma_load_4d(dstp + kDwXBytes, &tmDY, &bars[(it + 1) & 1], e.cb * 64, e.x0, e.y0, e.img);
        }
      } else {
        const uint32_t dst = sbase + ((it + 1) & 1) * kStage;
        dw_load_tile(x, dst, e.img, e.y0 - 3, e.x0 - 3, kDwHalo, kDwHalo, h, wd, e.cb * 64, stride);
        dw_load_tile(dy, dst + kDwXBytes, e.img, e.y0, e.x0, kDwT, kDwT, h, wd, e.cb * 64, stride);
      }
    }
    if (!TMA) dw_cp_commit();
    if (d.cb != cur_cb) {
      flush(cur_cb);
      cur_cb = d.cb;
    }
    if (TMA) {
      mbar_wait(&bars[it & 1], (it >> 1) & 1);
    } else {
      dw_cp_wait<1>();
      __syncthreads();
    }
    const uint32_t* xt = reinterpret_cast<const uint32_t*>(dw_smem + (it & 1) * kStage);
    const uint32_t* dt = reinterpret_cast<const uint32_t*>(dw_smem + (it & 1) * kStage + kDwXBytes);
#pragma unroll 1
    for (int u = 0; u < 4; ++u) {                         // output row 2*warp + (u >> 1), columns (u & 1) * 8 ..
      const int ro = 2 * warp + (u >> 1), cx = (u & 1) * 8;
      float2 g[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) g[q] = dw_unpack(dt[(ro * kDwT + cx + q) * 32 + lane]);
#pragma unroll
      for (int r = 0; r < 7; ++r) {
        const uint32_t* rowp = xt + ((ro + r) * kDwHalo + cx) * 32 + lane;
        float2 f[14];
#pragma unroll
        for (int j = 0; j < 14; ++j) f[j] = dw_unpack(rowp[j * 32]);
        // pixel-major order: consecutive FMAs go to 7 different partial sums (x 2 channels)
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
          for (int s7 = 0; s7 < 7; ++s7) {
            acc[r * 7 + s7].x = fmaf(g[q].x, f[q + s7].x, acc[r * 7 + s7].x);
            acc[r * 7 + s7].y = fmaf(g[q].y, f[q + s7].y, acc[r * 7 + s7].y);
          }
      }
    }
    __syncthreads();
  }
  if (!TMA) dw_cp_wait<0>();
  flush(cur_cb);
}

// (C, W, H, N) view of a channels-last bf16 map with a 64-channel x bw x bh box, linear in shared memory
int dw_make_map(CUtensorMap* tm, const void* base, int n, int h, int wd, int stride, int bw, int bh) {
  const uint64_t dims[4] = {(uint64_t)stride, (uint64_t)wd, (uint64_t)h, (uint64_t)n};
  const uint64_t strides[3] = {(uint64_t)stride * 2, (uint64_t)wd * stride * 2, (uint64_t)h * wd * stride * 2};
  const uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
  return aldi_make_tmap_bf16_sw(tm, base, 4, dims, strides, box, 0);
}
bool dw_tma_ok(const void* a, const void* b, int h, int wd, int stride) {
  static const bool off = getenv("ALDI_DW7_NO_TMA") != nullptr;
  return !off && h >= kDwHalo && wd >= kDwHalo && stride >= 64 && ((uintptr_t)a & 15) == 0 && (!b || ((uintptr_t)b & 15) == 0);
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_d(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * expf(-0.5f * x * x);
}
// bf16 activations: Phi(x) from the Abramowitz-Stegun 7.1.26 rational form of erf (|error| <= 1.5e-7, two orders below
// bf16 rounding), evaluated on the TAIL Phi(-|x|) = poly(t) e^{-x^2/2} / 2 so that negative x has no 1 + erf cancellation;
// the exponential is shared with the density term of the derivative.  ~14 instructions per element instead of erff's
// ~25 (+ expf), which had the kernel instruction-bound at the same level as its HBM floor.  fp32 (parity mode) keeps erff.
__device__ __forceinline__ float dw_ex2(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float dw_rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ void gelu_phi_fast(float x, float& Phi, float& e) {
  e = dw_ex2(x * x * (-0.5f * 1.44269504088896340736f));       // exp(-x^2 / 2): two MUFU ops per element in total
  const float t = dw_rcp(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float tail = p * t * e;
  Phi = x >= 0.f ? 1.f - tail : tail;
}
template <typename T, bool BWD> __device__ __forceinline__ float gelu_apply(float h, float g) {
  if constexpr (sizeof(T) == 2) {
    float Phi, e;
    gelu_phi_fast(h, Phi, e);
    return BWD ? g * fmaf(h * 0.39894228040143267794f, e, Phi) : h * Phi;
  } else {
    return BWD ? g * gelu_d(h) : gelu_f(h);
  }
}
template <typename T, bool BWD>
__global__ void __launch_bounds__(256) gelu_kernel(const T* __restrict__ h, const T* __restrict__ da, T* __restrict__ out, size_t n8) {
  // two vectors per thread per trip (the second one gridDim * blockDim further on): twice the loads in flight
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += 2 * step) {
    const size_t i2 = i + step;
    const bool two = i2 < n8;
    float f[8], g[8], f2[8], g2[8];
    V8<T>::load(h + i * 8, f);
    if (two) V8<T>::load(h + i2 * 8, f2);
    if (BWD) {
      V8<T>::load(da + i * 8, g);
      if (two) V8<T>::load(da + i2 * 8, g2);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = gelu_apply<T, BWD>(f[k], BWD ? g[k] : 0.f);
    V8<T>::store(out + i * 8, f);
    if (two) {
#pragma unroll
      for (int k = 0; k < 8; ++k) f2[k] = gelu_apply<T, BWD>(f2[k], BWD ? g2[k] : 0.f);
      V8<T>::store(out + i2 * 8, f2);
    }
  }
}

// out = input + gamma[c] * u * keep[n]           (keep = DropPath mask / keep_prob per sample, NULL = 1)
template <typename T>
__global__ void __launch_bounds__(256)
layerscale_fwd_kernel(const T* __restrict__ u, const T* __restrict__ input, const float* __restrict__ gamma,
                      const float* __restrict__ keep, long long rows, long long rows_per_image, int c, int stride,
                      T* __restrict__ out) {
  // (row, channel vector, image) advance incrementally: the 64-bit divisions of a flat index cost more instructions
  // per 16-byte vector than the arithmetic itself
  const int cv = stride / 8;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long r = id / cv;
  int col = (int)(id - r * cv);
  const long long dr = nthreads / cv;
  const int dc = (int)(nthreads - dr * cv);
  long long img = r / rows_per_image, rem = r - img * rows_per_image;
  while (r < rows) {
    const int ch = col * 8;
    const float m = keep ? keep[img] : 1.f;
    float a[8], b[8];
    V8<T>::load(u + r * stride + ch, a);
    V8<T>::load(input + r * stride + ch, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) b[k] += (ch + k < c) ? (gamma ? gamma[ch + k] : 1.f) * a[k] * m : 0.f;
    V8<T>::store(out + r * stride + ch, b);
    long long step = dr;
    col += dc;
    if (col >= cv) { col -= cv; ++step; }
    r += step;
    rem += step;
    while (rem >= rows_per_image) { rem -= rows_per_image; ++img; }
  }
}

// du = gamma * dy * keep;  dgamma[c] += sum dy * u * keep.  block = 8 channel groups x 32 row lanes, blockIdx.y = 64 channels
template <typename T>
__global__ void __launch_bounds__(256)
layerscale_bwd_kernel(const T* __restrict__ u, const T* __restrict__ dy, const float* __restrict__ gamma,
                      const float* __restrict__ keep, long long rows, long long rows_per_image, int c, int stride,
                      T* __restrict__ du, float* __restrict__ dgamma) {
  __shared__ float s_part[32][64];
  const int c0 = blockIdx.y * 64;
  const int cg = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int ch = c0 + cg * 8;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (ch < stride) {
    const long long r_first = (long long)blockIdx.x * 32 + pl, r_step = (long long)gridDim.x * 32;
    long long img = r_first / rows_per_image, rem = r_first - img * rows_per_image;
    for (long long r = r_first; r < rows; r += r_step) {
      const float m = keep ? keep[img] : 1.f;
      rem += r_step;
      while (rem >= rows_per_image) { rem -= rows_per_image; ++img; }
      float a[8], d[8], o[8];
      V8<T>::load(u + r * stride + ch, a);
      V8<T>::load(dy + r * stride + ch, d);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool ok = ch + k < c;
        acc[k] += ok ? d[k] * a[k] * m : 0.f;
        o[k] = ok ? (gamma ? gamma[ch + k] : 1.f) * d[k] * m : 0.f;
      }
      V8<T>::store(du + r * stride + ch, o);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) s_part[pl][cg * 8 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 64 && dgamma) {
    float v = 0.f;
    for (int i = 0; i < 32; ++i) v += s_part[i][threadIdx.x];
    if (c0 + threadIdx.x < c && v != 0.f) atomicAdd(dgamma + c0 + threadIdx.x, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// out[n, y, x, (dy*b + dx)*c + ch] = in[n, b*y + dy, b*x + dx, ch]; INVERSE scatters back (data gradient of the patchify conv)
template <typename T, bool INVERSE>
__global__ void __launch_bounds__(256)
s2d_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int ho, int wo, int b, int c, int in_stride, int out_stride) {
  const int kc = b * b * c;
  const size_t total = (size_t)n * ho * wo * out_stride;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % out_stride);
    size_t r = i / out_stride;
    const int x = (int)(r % wo);
    r /= wo;
    const int y = (int)(r % ho), img = (int)(r / ho);
    if (k >= kc) {
      if (!INVERSE) out[i] = from_f32<T>(0.f);
      continue;
    }
    const int ch = k % c, dd = k / c, dx = dd % b, dy = dd / b;
    const size_t fine = ((((size_t)img * ho * b + (size_t)y * b + dy) * wo * b) + (size_t)x * b + dx) * in_stride + ch;
    if (INVERSE) out[fine] = in[i];   // here `in` is the (coarse, b*b*c) gradient and `out` the fine map
    else out[i] = in[fine];
  }
}

// 8-channel vectors (c, both strides multiples of 8; fewer than 2^31 vectors): one 16-byte (bf16) move and one index
// decomposition per vector instead of per element
template <typename T> __device__ __forceinline__ void copy8(const T* src, T* dst);
template <> __device__ __forceinline__ void copy8<__nv_bfloat16>(const __nv_bfloat16* src, __nv_bfloat16* dst) {
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
}
template <> __device__ __forceinline__ void copy8<float>(const float* src, float* dst) {
  reinterpret_cast<float4*>(dst)[0] = reinterpret_cast<const float4*>(src)[0];
  reinterpret_cast<float4*>(dst)[1] = reinterpret_cast<const float4*>(src)[1];
}
template <typename T, bool INVERSE>
__global__ void __launch_bounds__(256)
s2d_vec_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int ho, int wo, int b, int c, int in_stride, int out_stride) {
  const unsigned kc = b * b * c, cvo = out_stride / 8;
  const unsigned total = (unsigned)n * ho * wo * cvo;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned k = (i % cvo) * 8;
    unsigned r = i / cvo;
    const unsigned x = r % wo;
    r /= wo;
    const unsigned y = r % ho, img = r / ho;
    if (k >= kc) {
      if (!INVERSE) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        V8<T>::store(out + (size_t)i * 8, z);
      }
      continue;
    }
    const unsigned ch = k % c, dd = k / c, dx = dd % b, dy = dd / b;
    const size_t fine = ((((size_t)img * ho * b + (size_t)y * b + dy) * wo * b) + (size_t)x * b + dx) * in_stride + ch;
    if (INVERSE) copy8<T>(in + (size_t)i * 8, out + fine);
    else copy8<T>(in + fine, out + (size_t)i * 8);
  }
}

struct Norm3 { float mean[3], stdv[3]; };
// images (N, 3, H, W) uint8 -> patches (N, H/b, W/b, pad64(b*b*3)) with value (px - mean[c]) / std[c], order (dy, dx, c)
template <typename T>
__global__ void __launch_bounds__(256)
patchify_kernel(const unsigned char* __restrict__ img, const int* __restrict__ sizes, T* __restrict__ out, int n, int hp, int wp,
                int b, int out_stride, Norm3 nm) {
  const int ho = hp / b, wo = wp / b, kc = b * b * 3;
  const size_t total = (size_t)n * ho * wo * out_stride;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % out_stride);
    size_t r = i / out_stride;
    const int x = (int)(r % wo);
    r /= wo;
    const int y = (int)(r % ho), im = (int)(r / ho);
    float v = 0.f;
    if (k < kc) {
      const int ch = k % 3, dd = k / 3, dx = dd % b, dy = dd / b;
      const int yy = y * b + dy, xx = x * b + dx;
      // pixels beyond the image's own size are padding of the batch canvas: zeros AFTER normalisation (ImageList)
      if (yy < sizes[2 * im] && xx < sizes[2 * im + 1])
        v = ((float)img[(((size_t)im * 3 + ch) * hp + yy) * wp + xx] - nm.mean[ch]) / nm.stdv[ch];
    }
    out[i] = from_f32<T>(v);
  }
}

// torch.optim.AdamW (decoupled weight decay, bias-corrected)
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g, size_t n, float lr,
             float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pi;
  }
}

}  // namespace

#define DISPATCH_T(dtype, CALL_F32, CALL_BF16)                        \
  do {                                                                \
    if ((dtype) == ALDI_DTYPE_BF16) { CALL_BF16; } else { CALL_F32; } \
  } while (0)

extern "C" int aldi_layernorm_forward(const void* x, const float* gamma, const float* beta, float eps, long long rows, int c,
                                      int stride, int dtype, void* y, float* stats, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(x && gamma && beta && y && rows > 0 && c > 0 && stride >= c, "aldi_layernorm_forward: bad args");
  static const bool legacy = getenv("ALDI_LN_LEGACY") != nullptr;
  const bool aligned16 = (((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0;
  if (!legacy && aligned16 && c % 8 == 0 && stride % 8 == 0 && c <= 2048) {
    const int nvec = c / 8;
#define LN_FWD(TT, VPL, RPW)                                                                                          \
  ln_fwd_vec_kernel<TT, VPL, RPW><<<blocks_for(rows, 8 * RPW, 8), 256, 0, stream>>>((const TT*)x, gamma, beta, eps, rows, c, \
                                                                                   stride, (TT*)y, (float2*)stats)
#define LN_FWD_T(TT)                          \
  do {                                        \
    if (nvec <= 32) LN_FWD(TT, 1, 4);         \
    else if (nvec <= 64) LN_FWD(TT, 2, 2);    \
    else if (nvec <= 96) LN_FWD(TT, 3, 1);    \
    else if (nvec <= 128) LN_FWD(TT, 4, 1);   \
    else if (nvec <= 192) LN_FWD(TT, 6, 1);   \
    else LN_FWD(TT, 8, 1);                    \
  } while (0)
    if (dtype == ALDI_DTYPE_BF16) LN_FWD_T(__nv_bfloat16);
    else LN_FWD_T(float);
#undef LN_FWD_T
#undef LN_FWD
    ALDI_COUNT_LAUNCH();
    ALDI_CUDA_LAUNCH_CHECK("aldi_layernorm_forward");
    return ALDI_OK;
  }
  const int grid = blocks_for(rows, 8, 8);
  DISPATCH_T(dtype,
             (ln_fwd_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, gamma, beta, eps, rows, c, stride, (float*)y, (float2*)stats)),
             (ln_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, gamma, beta, eps, rows, c, stride,
                                                                    (__nv_bfloat16*)y, (float2*)stats)));
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_layernorm_forward");
  return ALDI_OK;
}

extern "C" int aldi_layernorm_backward(const void* x, const float* gamma, const float* stats, const void* dy, long long rows,
                                       int c, int stride, int dtype, void* dx, int accumulate, float* dgamma, float* dbeta,
                                       void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(x && gamma && stats && dy && dx && dgamma && dbeta && rows > 0 && c > 0 && stride >= c,
                 "aldi_layernorm_backward: bad args");
  ALDI_CHECK_ARG(c <= 2048, "aldi_layernorm_backward: at most 2048 channels");
  const int grid = blocks_for(rows, 8 * 16, 4);
  const size_t smem = (size_t)2 * c * sizeof(float);
  static const bool legacy = getenv("ALDI_LN_LEGACY") != nullptr;
  const bool aligned16 = (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)gamma) & 15) == 0;
  if (!legacy && aligned16 && c % 8 == 0 && stride % 8 == 0) {
    const int nvec = c / 8;
    // enough CTAs to fill the GPU even for a few thousand rows, few enough that the per-CTA d-gamma / d-beta atomics
    // (2 c each) stay cheap: >= 4 rows per warp, at most two CTAs per SM
    const int vgrid = blocks_for(rows, 8 * 4, 2);
#define LN_BWDV(TT, VPL, RPW)                                                                                                  \
  ln_bwd_vec_kernel<TT, VPL, RPW><<<vgrid, 256, smem, stream>>>((const TT*)x, gamma, (const float2*)stats, (const TT*)dy, rows, c,    \
                                                               stride, (TT*)dx, accumulate, dgamma, dbeta)
#define LN_BWDV_T(TT)                         \
  do {                                        \
    if (nvec <= 32) LN_BWDV(TT, 1, 4);        \
    else if (nvec <= 64) LN_BWDV(TT, 2, 1);   \
    else if (nvec <= 96) LN_BWDV(TT, 3, 1);   \
    else if (nvec <= 128) LN_BWDV(TT, 4, 1);  \
    else if (nvec <= 192) LN_BWDV(TT, 6, 1);  \
    else LN_BWDV(TT, 8, 1);                   \
  } while (0)
    if (dtype == ALDI_DTYPE_BF16) LN_BWDV_T(__nv_bfloat16);
    else LN_BWDV_T(float);
#undef LN_BWDV_T
#undef LN_BWDV
    ALDI_COUNT_LAUNCH();
    ALDI_CUDA_LAUNCH_CHECK("aldi_layernorm_backward");
    return ALDI_OK;
  }
#define LN_BWD(TT, CPL)                                                                                                    \
  ln_bwd_kernel<TT, CPL><<<grid, 256, smem, stream>>>((const TT*)x, gamma, (const float2*)stats, (const TT*)dy, rows, c, stride, \
                                                      (TT*)dx, accumulate, dgamma, dbeta)
#define LN_BWD_T(TT)                          \
  do {                                        \
    if (c <= 256) LN_BWD(TT, 8);              \
    else if (c <= 512) LN_BWD(TT, 16);        \
    else if (c <= 1024) LN_BWD(TT, 32);       \
    else LN_BWD(TT, 64);                      \
  } while (0)
  if (dtype == ALDI_DTYPE_BF16) LN_BWD_T(__nv_bfloat16);
  else LN_BWD_T(float);
#undef LN_BWD_T
#undef LN_BWD
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_layernorm_backward");
  return ALDI_OK;
}

extern "C" int aldi_dwconv7(const void* x, const float* w, const float* bias, int n, int h, int wd, int c, int stride, int dtype,
                            int flip, void* y, int accumulate, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(x && w && y && n > 0 && h > 0 && wd > 0 && c > 0 && stride >= c && stride % 8 == 0, "aldi_dwconv7: bad args");
  static const bool legacy = getenv("ALDI_DW7_LEGACY") != nullptr;
  if (dtype == ALDI_DTYPE_BF16 && !legacy) {
    static bool attr = false;
    if (!attr) {
      ALDI_CUDA_CHECK(cudaFuncSetAttribute(dw7_tile_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwFwdSmem));
      ALDI_CUDA_CHECK(cudaFuncSetAttribute(dw7_tile_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwFwdSmem));
      ALDI_CUDA_CHECK(cudaFuncSetAttribute(dw7_tile_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwFwdSmem));
      ALDI_CUDA_CHECK(cudaFuncSetAttribute(dw7_tile_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwFwdSmem));
      attr = true;
    }
    const long long tiles = (long long)((stride + 63) / 64) * n * ((h + kDwT - 1) / kDwT) * ((wd + kDwT - 1) / kDwT);
    const int tgrid = (int)std::min<long long>(tiles, aldi_num_sms());
    const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
    __nv_bfloat16* yb = (__nv_bfloat16*)y;
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    if (dw_tma_ok(x, nullptr, h, wd, stride)) {
      int rc = dw_make_map(&tm, x, n, h, wd, stride, kDwHalo, kDwHalo);
      if (rc != ALDI_OK) return rc;
      if (flip) dw7_tile_kernel<true, true><<<tgrid, 256, kDwFwdSmem, stream>>>(tm, xb, w, bias, n, h, wd, c, stride, yb, accumulate);
      else dw7_tile_kernel<false, true><<<tgrid, 256, kDwFwdSmem, stream>>>(tm, xb, w, bias, n, h, wd, c, stride, yb, accumulate);
    } else {
      if (flip) dw7_tile_kernel<true, false><<<tgrid, 256, kDwFwdSmem, stream>>>(tm, xb, w, bias, n, h, wd, c, stride, yb, accumulate);
      else dw7_tile_kernel<false, false><<<tgrid, 256, kDwFwdSmem, stream>>>(tm, xb, w, bias, n, h, wd, c, stride, yb, accumulate);
    }
    ALDI_COUNT_LAUNCH();
    ALDI_CUDA_LAUNCH_CHECK("aldi_dwconv7");
    return ALDI_OK;
  }
  const dim3 grid(blocks_for((long long)n * h * ((wd + 3) / 4), 32, 16), (stride + 63) / 64);
  if (dtype == ALDI_DTYPE_BF16) {
    if (flip) dw7_kernel<__nv_bfloat16, true><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, w, bias, n, h, wd, c, stride, (__nv_bfloat16*)y, accumulate);
    else dw7_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, w, bias, n, h, wd, c, stride, (__nv_bfloat16*)y, accumulate);
  } else {
    if (flip) dw7_kernel<float, true><<<grid, 256, 0, stream>>>((const float*)x, w, bias, n, h, wd, c, stride, (float*)y, accumulate);
    else dw7_kernel<float, false><<<grid, 256, 0, stream>>>((const float*)x, w, bias, n, h, wd, c, stride, (float*)y, accumulate);
  }
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_dwconv7");
  return ALDI_OK;
}

extern "C" int aldi_dwconv7_wgrad(const void* x, const void* dy, int n, int h, int wd, int c, int stride, int dtype, float* dw,
                                  void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(x && dy && dw && n > 0 && c > 0 && stride >= c && stride % 8 == 0, "aldi_dwconv7_wgrad: bad args");
  static const bool legacy = getenv("ALDI_DW7_LEGACY") != nullptr;
  if (dtype == ALDI_DTYPE_BF16 && !legacy) {
    static bool attr = false;
    if (!attr) {
      ALDI_CUDA_CHECK(cudaFuncSetAttribute(dw7_wgrad_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwWgSmem));
      ALDI_CUDA_CHECK(cudaFuncSetAttribute(dw7_wgrad_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwWgSmem));
      attr = true;
    }
    const long long tiles = (long long)((stride + 63) / 64) * n * ((h + kDwT - 1) / kDwT) * ((wd + kDwT - 1) / kDwT);
    const int tgrid = (int)std::min<long long>(tiles, aldi_num_sms());
    const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
    const __nv_bfloat16* db = (const __nv_bfloat16*)dy;
    CUtensorMap tmx, tmd;
    memset(&tmx, 0, sizeof(tmx));
    memset(&tmd, 0, sizeof(tmd));
    if (dw_tma_ok(x, dy, h, wd, stride)) {
      int rc = dw_make_map(&tmx, x, n, h, wd, stride, kDwHalo, kDwHalo);
      if (rc == ALDI_OK) rc = dw_make_map(&tmd, dy, n, h, wd, stride, kDwT, kDwT);
      if (rc != ALDI_OK) return rc;
      dw7_wgrad_tile_kernel<true><<<tgrid, 256, kDwWgSmem, stream>>>(tmx, tmd, xb, db, n, h, wd, c, stride, dw);
    } else {
      dw7_wgrad_tile_kernel<false><<<tgrid, 256, kDwWgSmem, stream>>>(tmx, tmd, xb, db, n, h, wd, c, stride, dw);
    }
    ALDI_COUNT_LAUNCH();
    ALDI_CUDA_LAUNCH_CHECK("aldi_dwconv7_wgrad");
    return ALDI_OK;
  }
  const dim3 grid(blocks_for((long long)n * h * ((wd + kWgRun - 1) / kWgRun), 32, 4), (stride + 63) / 64, 7);
  DISPATCH_T(dtype, (dw7_wgrad_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, (const float*)dy, n, h, wd, c, stride, dw)),
             (dw7_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, n, h, wd, c,
                                                                        stride, dw)));
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_dwconv7_wgrad");
  return ALDI_OK;
}

extern "C" int aldi_gelu(const void* h, const void* da, void* out, size_t n, int dtype, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(h && out && n % 8 == 0, "aldi_gelu: n must be a multiple of 8");
  if (n == 0) return ALDI_OK;
  const int grid = blocks_for((long long)(n / 8), 256, 8);
  if (dtype == ALDI_DTYPE_BF16) {
    if (da) gelu_kernel<__nv_bfloat16, true><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)h, (const __nv_bfloat16*)da, (__nv_bfloat16*)out, n / 8);
    else gelu_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)h, nullptr, (__nv_bfloat16*)out, n / 8);
  } else {
    if (da) gelu_kernel<float, true><<<grid, 256, 0, stream>>>((const float*)h, (const float*)da, (float*)out, n / 8);
    else gelu_kernel<float, false><<<grid, 256, 0, stream>>>((const float*)h, nullptr, (float*)out, n / 8);
  }
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_gelu");
  return ALDI_OK;
}

extern "C" int aldi_layerscale_forward(const void* u, const void* input, const float* gamma, const float* keep, long long rows,
                                       long long rows_per_image, int c, int stride, int dtype, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(u && input && out && rows > 0 && stride % 8 == 0 && stride >= c && rows_per_image > 0, "aldi_layerscale_forward: bad args");
  const int grid = blocks_for(rows * (stride / 8), 256, 8);
  DISPATCH_T(dtype,
             (layerscale_fwd_kernel<float><<<grid, 256, 0, stream>>>((const float*)u, (const float*)input, gamma, keep, rows, rows_per_image, c,
                                                                     stride, (float*)out)),
             (layerscale_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)u, (const __nv_bfloat16*)input, gamma, keep,
                                                                             rows, rows_per_image, c, stride, (__nv_bfloat16*)out)));
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_layerscale_forward");
  return ALDI_OK;
}

extern "C" int aldi_layerscale_backward(const void* u, const void* dy, const float* gamma, const float* keep, long long rows,
                                        long long rows_per_image, int c, int stride, int dtype, void* du, float* dgamma,
                                        void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(u && dy && du && rows > 0 && stride % 8 == 0 && stride >= c && rows_per_image > 0, "aldi_layerscale_backward: bad args");
  const dim3 grid(blocks_for(rows, 32 * 16, 4), (stride + 63) / 64);
  DISPATCH_T(dtype,
             (layerscale_bwd_kernel<float><<<grid, 256, 0, stream>>>((const float*)u, (const float*)dy, gamma, keep, rows, rows_per_image, c,
                                                                     stride, (float*)du, dgamma)),
             (layerscale_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)u, (const __nv_bfloat16*)dy, gamma, keep,
                                                                             rows, rows_per_image, c, stride, (__nv_bfloat16*)du, dgamma)));
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_layerscale_backward");
  return ALDI_OK;
}

extern "C" int aldi_space_to_depth(const void* in, void* out, int n, int ho, int wo, int block, int c, int in_stride,
                                   int out_stride, int dtype, int inverse, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(in && out && n > 0 && ho > 0 && wo > 0 && block >= 1 && c > 0 && in_stride >= c && out_stride >= block * block * c,
                 "aldi_space_to_depth: bad args");
  const long long nvec = (long long)n * ho * wo * (out_stride / 8);
  if (c % 8 == 0 && in_stride % 8 == 0 && out_stride % 8 == 0 && nvec < (1LL << 31) && ((uintptr_t)in & 15) == 0 &&
      ((uintptr_t)out & 15) == 0) {
    const int vgrid = blocks_for(nvec, 256, 8);
    if (dtype == ALDI_DTYPE_BF16) {
      if (inverse) s2d_vec_kernel<__nv_bfloat16, true><<<vgrid, 256, 0, stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, n, ho, wo, block, c, in_stride, out_stride);
      else s2d_vec_kernel<__nv_bfloat16, false><<<vgrid, 256, 0, stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, n, ho, wo, block, c, in_stride, out_stride);
    } else {
      if (inverse) s2d_vec_kernel<float, true><<<vgrid, 256, 0, stream>>>((const float*)in, (float*)out, n, ho, wo, block, c, in_stride, out_stride);
      else s2d_vec_kernel<float, false><<<vgrid, 256, 0, stream>>>((const float*)in, (float*)out, n, ho, wo, block, c, in_stride, out_stride);
    }
    ALDI_COUNT_LAUNCH();
    ALDI_CUDA_LAUNCH_CHECK("aldi_space_to_depth");
    return ALDI_OK;
  }
  const int grid = blocks_for((long long)n * ho * wo * out_stride, 256, 16);
  if (dtype == ALDI_DTYPE_BF16) {
    if (inverse) s2d_kernel<__nv_bfloat16, true><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, n, ho, wo, block, c, in_stride, out_stride);
    else s2d_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, n, ho, wo, block, c, in_stride, out_stride);
  } else {
    if (inverse) s2d_kernel<float, true><<<grid, 256, 0, stream>>>((const float*)in, (float*)out, n, ho, wo, block, c, in_stride, out_stride);
    else s2d_kernel<float, false><<<grid, 256, 0, stream>>>((const float*)in, (float*)out, n, ho, wo, block, c, in_stride, out_stride);
  }
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_space_to_depth");
  return ALDI_OK;
}

extern "C" int aldi_patchify_image(const unsigned char* images, const int* sizes, void* out, int n, int hp, int wp, int block,
                                   int out_stride, int dtype, const float* h_mean, const float* h_std, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(images && sizes && out && h_mean && h_std && n > 0 && block >= 1 && hp % block == 0 && wp % block == 0 &&
                     out_stride >= block * block * 3, "aldi_patchify_image: bad args");
  Norm3 nm;
  for (int i = 0; i < 3; ++i) { nm.mean[i] = h_mean[i]; nm.stdv[i] = h_std[i]; }
  const int grid = blocks_for((long long)n * (hp / block) * (wp / block) * out_stride, 256, 16);
  DISPATCH_T(dtype, (patchify_kernel<float><<<grid, 256, 0, stream>>>(images, sizes, (float*)out, n, hp, wp, block, out_stride, nm)),
             (patchify_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(images, sizes, (__nv_bfloat16*)out, n, hp, wp, block, out_stride, nm)));
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_patchify_image");
  return ALDI_OK;
}

extern "C" int aldi_adamw_step(float* params, float* exp_avg, float* exp_avg_sq, const float* grads, size_t n, float lr,
                               float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                               void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(params && exp_avg && exp_avg_sq && grads && step >= 1, "aldi_adamw_step: bad args");
  if (n == 0) return ALDI_OK;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adamw_kernel<<<blocks_for((long long)n, 256, 8), 256, 0, stream>>>(params, exp_avg, exp_avg_sq, grads, n, lr, beta1, beta2, eps,
                                                                     weight_decay, bc1, bc2_sqrt, grad_scale);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_adamw_step");
  return ALDI_OK;
}
