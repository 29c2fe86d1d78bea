// HBM-bound glue kernels of the detector trunk: image normalisation, stem im2col, max-pool, FPN
// top-down backward, bias gradients, FrozenBN folding.  Coalesced channels-last accesses, grids capped at a
// multiple of the SM count (grid-stride loops).
#include "common.cuh"
#include "../../include/aldi_b200.h"

namespace {

int grid_for(size_t work, int threads) {
  size_t blocks = (work + threads - 1) / threads;
  size_t cap = (size_t)aldi_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

struct Norm3 {
  float mean[3], stdv[3];
};

// ---- GeneralizedRCNN.preprocess_image (D2 rcnn.py): (x - mean) / std, zero pad to the batch canvas ----
// in: uint8 (N,3,Hin,Win) planar;  out: fp32 (N,Hp,Wp,4) channels-last, 4th channel zero.
__global__ void __launch_bounds__(256)
preprocess_kernel(const uint8_t* __restrict__ in, const int* __restrict__ sizes, float* __restrict__ out, int n,
                  int hin, int win, int hp, int wp, Norm3 nm) {
  const size_t total = (size_t)n * hp * wp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % wp);
    size_t r = i / wp;
    const int y = (int)(r % hp);
    const int b = (int)(r / hp);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < sizes[2 * b] && x < sizes[2 * b + 1]) {
      const uint8_t* p = in + ((size_t)b * 3 * hin + y) * win + x;
      v.x = __fdiv_rn(__fsub_rn((float)p[0], nm.mean[0]), nm.stdv[0]);
      v.y = __fdiv_rn(__fsub_rn((float)p[(size_t)hin * win], nm.mean[1]), nm.stdv[1]);
      v.z = __fdiv_rn(__fsub_rn((float)p[(size_t)2 * hin * win], nm.mean[2]), nm.stdv[2]);
    }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

// ---- stem 7x7/2 conv as a GEMM: fused normalise + im2col of the uint8 image -------------------------
// out: bf16 (N, Ho, Wo, 192): k = (r*7 + s)*3 + c for the 147 real taps, zero for k >= 147.
// One thread produces 8 consecutive k (16 B store); zero padding = conv padding and canvas padding.
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const uint8_t* __restrict__ in, const int* __restrict__ sizes, __nv_bfloat16* __restrict__ out,
                   int n, int hin, int win, int ho, int wo, Norm3 nm) {
  const size_t total = (size_t)n * ho * wo * 24;  // 24 vectors of 8 per output pixel
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kv = (int)(i % 24);
    size_t r = i / 24;
    const int ox = (int)(r % wo);
    r /= wo;
    const int oy = (int)(r % ho);
    const int b = (int)(r / ho);
    const int vh = sizes[2 * b], vw = sizes[2 * b + 1];
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kv * 8 + e;
      float v = 0.f;
      if (k < 147) {
        const int tap = k / 3, c = k - tap * 3;
        const int rr = tap / 7, ss = tap - rr * 7;
        const int iy = oy * 2 + rr - 3, ix = ox * 2 + ss - 3;
        if (iy >= 0 && iy < vh && ix >= 0 && ix < vw)
          v = __fdiv_rn(__fsub_rn((float)__ldg(in + (((size_t)b * 3 + c) * hin + iy) * win + ix), nm.mean[c]),
                        nm.stdv[c]);
      }
      f[e] = v;
    }
    uint4 q;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    reinterpret_cast<uint4*>(out)[i] = q;
  }
}

// ---- max_pool2d(kernel 3, stride 2, padding 1) on channels-last ------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int h, int w, int c, int ho, int wo) {
  const size_t total = (size_t)n * ho * wo * c;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    size_t r = i / c;
    const int ox = (int)(r % wo);
    r /= wo;
    const int oy = (int)(r % ho);
    const int b = (int)(r / ho);
    float m = -INFINITY;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int iy = oy * 2 + dy;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int ix = ox * 2 + dx;
        if (ix < 0 || ix >= w) continue;
        m = fmaxf(m, to_f32<T>(in[(((size_t)b * h + iy) * w + ix) * c + ch]));
      }
    }
    out[i] = from_f32<T>(m);
  }
}

// ---- dst[n,h,w,c] += sum_{i,j in 0..1} src[n,2h+i,2w+j,c]  (backward of nearest-2x upsample + add) --
template <typename T>
__global__ void __launch_bounds__(256)
sum2x2_kernel(const T* __restrict__ src, T* __restrict__ dst, int n, int h, int w, int c) {
  const size_t total = (size_t)n * h * w * c;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    size_t r = i / c;
    const int x = (int)(r % w);
    r /= w;
    const int y = (int)(r % h);
    const int b = (int)(r / h);
    const size_t base = (((size_t)b * 2 * h + 2 * y) * 2 * w + 2 * x) * c + ch;
    const size_t row = (size_t)2 * w * c;
    float s = to_f32<T>(src[base]) + to_f32<T>(src[base + c]) + to_f32<T>(src[base + row]) +
              to_f32<T>(src[base + row + c]);
    dst[i] = from_f32<T>(to_f32<T>(dst[i]) + s);
  }
}

// ---- dst += src (fp32 source, e.g. RoIAlign's atomic gradient buffer) -------------------------------
template <typename T>
__global__ void __launch_bounds__(256) add_f32_kernel(T* __restrict__ dst, const float* __restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = from_f32<T>(to_f32<T>(dst[i]) + src[i]);
}

// ---- out[c] += scale * sum_rows x[row, c]   (bias gradients) ----------------------------------------
// grid (chunks of rows, channel groups of 32); block 32 x 8: lanes over channels (coalesced), 8 row lanes.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, long long row_stride, long long rows, int c, float scale,
              float* __restrict__ out) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int ch = blockIdx.y * 32 + lane;
  float s = 0.f;
  if (ch < c) {
    for (long long r = (long long)blockIdx.x * 8 + ry; r < rows; r += (long long)gridDim.x * 8)
      s += to_f32<T>(x[r * row_stride + ch]);
  }
  part[ry][lane] = s;
  __syncthreads();
  if (ry == 0 && ch < c) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][lane];
    atomicAdd(out + ch, t * scale);
  }
}

// ---- FrozenBatchNorm2d folded to y = x*scale + shift (D2 layers/batch_norm.py) ----------------------
__global__ void frozenbn_fold_kernel(const float* w, const float* b, const float* mean, const float* var, float eps,
                                     float* scale, float* shift, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float s = __fmul_rn(w[i], __frsqrt_rn(__fadd_rn(var[i], eps)));
    scale[i] = s;
    shift[i] = __fsub_rn(b[i], __fmul_rn(mean[i], s));
  }
}

}  // namespace

extern "C" int aldi_preprocess(const uint8_t* images, const int* sizes, float* out, int n, int hin, int win, int hp,
                               int wp, const float* h_mean, const float* h_std, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(images && sizes && out && h_mean && h_std, "aldi_preprocess: null pointer");
  ALDI_CHECK_ARG(n > 0 && hp >= hin && wp >= win, "aldi_preprocess: bad sizes");
  Norm3 nm;
  for (int i = 0; i < 3; ++i) { nm.mean[i] = h_mean[i]; nm.stdv[i] = h_std[i]; }
  preprocess_kernel<<<grid_for((size_t)n * hp * wp, 256), 256, 0, stream>>>(images, sizes, out, n, hin, win, hp, wp, nm);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_preprocess");
  return ALDI_OK;
}

extern "C" int aldi_stem_im2col(const uint8_t* images, const int* sizes, void* out_bf16, int n, int hin, int win,
                                int ho, int wo, const float* h_mean, const float* h_std, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(images && sizes && out_bf16 && h_mean && h_std, "aldi_stem_im2col: null pointer");
  Norm3 nm;
  for (int i = 0; i < 3; ++i) { nm.mean[i] = h_mean[i]; nm.stdv[i] = h_std[i]; }
  stem_im2col_kernel<<<grid_for((size_t)n * ho * wo * 24, 256), 256, 0, stream>>>(
      images, sizes, (__nv_bfloat16*)out_bf16, n, hin, win, ho, wo, nm);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_stem_im2col");
  return ALDI_OK;
}

extern "C" int aldi_maxpool3x3s2(const void* in, void* out, int dtype, int n, int h, int w, int c, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(in && out, "aldi_maxpool3x3s2: null pointer");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const int grid = grid_for((size_t)n * ho * wo * c, 256);
  if (dtype == ALDI_DTYPE_BF16)
    maxpool_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, n, h, w, c, ho, wo);
  else
    maxpool_kernel<float><<<grid, 256, 0, stream>>>((const float*)in, (float*)out, n, h, w, c, ho, wo);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_maxpool3x3s2");
  return ALDI_OK;
}

extern "C" int aldi_sum2x2_accum(const void* fine, void* coarse, int dtype, int n, int h, int w, int c, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(fine && coarse, "aldi_sum2x2_accum: null pointer");
  const int grid = grid_for((size_t)n * h * w * c, 256);
  if (dtype == ALDI_DTYPE_BF16)
    sum2x2_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)fine, (__nv_bfloat16*)coarse, n, h, w, c);
  else
    sum2x2_kernel<float><<<grid, 256, 0, stream>>>((const float*)fine, (float*)coarse, n, h, w, c);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_sum2x2_accum");
  return ALDI_OK;
}

extern "C" int aldi_add_f32(void* dst, int dtype, const float* src, size_t n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(dst && src, "aldi_add_f32: null pointer");
  if (n == 0) return ALDI_OK;
  const int grid = grid_for(n, 256);
  if (dtype == ALDI_DTYPE_BF16)
    add_f32_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((__nv_bfloat16*)dst, src, n);
  else
    add_f32_kernel<float><<<grid, 256, 0, stream>>>((float*)dst, src, n);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_add_f32");
  return ALDI_OK;
}

extern "C" int aldi_colsum(const void* x, int dtype, long long rows, long long row_stride, int c, float scale,
                           float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(x && out && c > 0, "aldi_colsum: bad args");
  if (rows <= 0) return ALDI_OK;
  long long gx = (rows + 8 * 64 - 1) / (8 * 64);
  long long cap = (long long)aldi_num_sms() * 8;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)((c + 31) / 32));
  if (dtype == ALDI_DTYPE_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, row_stride, rows, c, scale, out);
  else
    colsum_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, row_stride, rows, c, scale, out);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_colsum");
  return ALDI_OK;
}

extern "C" int aldi_frozenbn_fold(const float* weight, const float* bias, const float* mean, const float* var,
                                  float eps, float* scale, float* shift, int n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(weight && bias && mean && var && scale && shift && n > 0, "aldi_frozenbn_fold: bad args");
  frozenbn_fold_kernel<<<(n + 255) / 256, 256, 0, stream>>>(weight, bias, mean, var, eps, scale, shift, n);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_frozenbn_fold");
  return ALDI_OK;
}
