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
#include "common.cuh"
#include "../../include/aldi_b200.h"

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
// Each warp keeps per-lane partial dgamma / dbeta for channels lane, lane+32, ... (c <= 32 * kMaxPerLane).
constexpr int kMaxPerLane = 64;
template <typename T>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float2* __restrict__ stats,
              const T* __restrict__ dy, long long rows, int c, int stride, T* __restrict__ dx, int accumulate,
              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float s_red[];   // [2][c]
  const int lane = threadIdx.x & 31;
  const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, wstep = ((long long)gridDim.x * blockDim.x) >> 5;
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  float pg[kMaxPerLane / 4], pb[kMaxPerLane / 4];   // register partials for the first 16 strides; the rest go to smem
#pragma unroll
  for (int k = 0; k < kMaxPerLane / 4; ++k) pg[k] = pb[k] = 0.f;
  for (long long r = w0; r < rows; r += wstep) {
    const T* xr = x + r * stride;
    const T* dr = dy + r * stride;
    const float2 st = stats[r];
    float sg = 0.f, sgx = 0.f;
    for (int i = lane; i < c; i += 32) {
      const float xh = (to_f32<T>(xr[i]) - st.x) * st.y, g = to_f32<T>(dr[i]) * gamma[i];
      sg += g;
      sgx += g * xh;
    }
    const float mg = warp_sum(sg) / (float)c, mgx = warp_sum(sgx) / (float)c;
    T* dxr = dx + r * stride;
    int k = 0;
    for (int i = lane; i < c; i += 32, ++k) {
      const float d = to_f32<T>(dr[i]);
      const float xh = (to_f32<T>(xr[i]) - st.x) * st.y, g = d * gamma[i];
      float v = st.y * (g - mg - xh * mgx);
      if (accumulate) v += to_f32<T>(dxr[i]);
      dxr[i] = from_f32<T>(v);
      if (k < kMaxPerLane / 4) { pg[k] += d * xh; pb[k] += d; }
      else { atomicAdd(&s_red[i], d * xh); atomicAdd(&s_red[c + i], d); }
    }
    if (!accumulate)
      for (int i = c + lane; i < stride; i += 32) dxr[i] = from_f32<T>(0.f);
  }
  int k = 0;
  for (int i = lane; i < c && k < kMaxPerLane / 4; i += 32, ++k) { atomicAdd(&s_red[i], pg[k]); atomicAdd(&s_red[c + i], pb[k]); }
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
  const long long total = (long long)n * h * wd;
  for (long long p = (long long)blockIdx.x * 32 + pl; p < total; p += (long long)gridDim.x * 32) {
    const int px = (int)(p % wd);
    const long long q = p / wd;
    const int py = (int)(q % h), img = (int)(q / h);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = (bias && ch + k < c) ? bias[ch + k] : 0.f;
    for (int r = 0; r < 7; ++r) {
      const int yy = py + r - 3;
      if (yy < 0 || yy >= h) continue;
      for (int s = 0; s < 7; ++s) {
        const int xx = px + s - 3;
        if (xx < 0 || xx >= wd) continue;
        float f[8];
        V8<T>::load(x + (((long long)img * h + yy) * wd + xx) * stride + ch, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k] * s_w[r * 7 + s][cg * 8 + k];
      }
    }
    T* yp = y + p * stride + ch;
    if (accumulate) {
      float f[8];
      V8<T>::load(yp, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
    V8<T>::store(yp, acc);
  }
}

// weight gradient: dw[c][t] += sum_pixels dy[p, c] * x[p + t - 3, c].  block = (8 channel groups of 8) x 32 pixel lanes,
// blockIdx.y = 64-channel block, blockIdx.z = tap; pixels grid-strided over blockIdx.x.
template <typename T>
__global__ void __launch_bounds__(256)
dw7_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, int n, int h, int wd, int c, int stride,
                 float* __restrict__ dw) {
  __shared__ float s_part[32][64];
  const int c0 = blockIdx.y * 64, t = blockIdx.z, r = t / 7, s = t % 7;
  const int cg = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int ch = c0 + cg * 8;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (ch < stride) {
    const long long total = (long long)n * h * wd;
    for (long long p = (long long)blockIdx.x * 32 + pl; p < total; p += (long long)gridDim.x * 32) {
      const int px = (int)(p % wd);
      const long long q = p / wd;
      const int py = (int)(q % h), img = (int)(q / h);
      const int yy = py + r - 3, xx = px + s - 3;
      if (yy < 0 || yy >= h || xx < 0 || xx >= wd) continue;
      float a[8], b[8];
      V8<T>::load(dy + p * stride + ch, a);
      V8<T>::load(x + (((long long)img * h + yy) * wd + xx) * stride + ch, b);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += a[k] * b[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) s_part[pl][cg * 8 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 64) {
    float v = 0.f;
    for (int i = 0; i < 32; ++i) v += s_part[i][threadIdx.x];
    if (c0 + threadIdx.x < c && v != 0.f) atomicAdd(dw + (size_t)(c0 + threadIdx.x) * 49 + t, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_d(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * expf(-0.5f * x * x);
}
template <typename T, bool BWD>
__global__ void __launch_bounds__(256) gelu_kernel(const T* __restrict__ h, const T* __restrict__ da, T* __restrict__ out, size_t n8) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float f[8], g[8];
    V8<T>::load(h + i * 8, f);
    if (BWD) V8<T>::load(da + i * 8, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = BWD ? g[k] * gelu_d(f[k]) : gelu_f(f[k]);
    V8<T>::store(out + i * 8, f);
  }
}

// out = input + gamma[c] * u * keep[n]           (keep = DropPath mask / keep_prob per sample, NULL = 1)
template <typename T>
__global__ void __launch_bounds__(256)
layerscale_fwd_kernel(const T* __restrict__ u, const T* __restrict__ input, const float* __restrict__ gamma,
                      const float* __restrict__ keep, long long rows, long long rows_per_image, int c, int stride,
                      T* __restrict__ out) {
  const int cv = stride / 8;
  const size_t total = (size_t)rows * cv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cv) * 8;
    const long long r = (long long)(i / cv);
    const float m = keep ? keep[r / rows_per_image] : 1.f;
    float a[8], b[8];
    V8<T>::load(u + r * stride + ch, a);
    V8<T>::load(input + r * stride + ch, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) b[k] += (ch + k < c) ? (gamma ? gamma[ch + k] : 1.f) * a[k] * m : 0.f;
    V8<T>::store(out + r * stride + ch, b);
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
    for (long long r = (long long)blockIdx.x * 32 + pl; r < rows; r += (long long)gridDim.x * 32) {
      const float m = keep ? keep[r / rows_per_image] : 1.f;
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
  ALDI_CHECK_ARG(c <= 32 * kMaxPerLane * 4, "aldi_layernorm_backward: at most %d channels", 32 * kMaxPerLane * 4);
  const int grid = blocks_for(rows, 8 * 16, 4);
  const size_t smem = (size_t)2 * c * sizeof(float);
  DISPATCH_T(dtype,
             (ln_bwd_kernel<float><<<grid, 256, smem, stream>>>((const float*)x, gamma, (const float2*)stats, (const float*)dy, rows, c,
                                                                stride, (float*)dx, accumulate, dgamma, dbeta)),
             (ln_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>((const __nv_bfloat16*)x, gamma, (const float2*)stats,
                                                                        (const __nv_bfloat16*)dy, rows, c, stride,
                                                                        (__nv_bfloat16*)dx, accumulate, dgamma, dbeta)));
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_layernorm_backward");
  return ALDI_OK;
}

extern "C" int aldi_dwconv7(const void* x, const float* w, const float* bias, int n, int h, int wd, int c, int stride, int dtype,
                            int flip, void* y, int accumulate, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(x && w && y && n > 0 && h > 0 && wd > 0 && c > 0 && stride >= c && stride % 8 == 0, "aldi_dwconv7: bad args");
  const dim3 grid(blocks_for((long long)n * h * wd, 32, 16), (stride + 63) / 64);
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
  const dim3 grid(blocks_for((long long)n * h * wd, 32 * 64, 2), (stride + 63) / 64, 49);
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
