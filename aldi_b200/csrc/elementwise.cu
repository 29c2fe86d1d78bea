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
// One block = 4 x 64 output pixels: the (13 x 133 x 3) input patch is read once, coalesced, into shared
// memory (normalised, zero outside the valid image = conv padding + canvas padding); the 256 x 384 B of
// output rows are then written as consecutive 16-byte vectors (fully coalesced, write-bound kernel).
constexpr int ST_TH = 4, ST_TW = 64, ST_PH = 2 * ST_TH + 5, ST_PW = 2 * ST_TW + 5;
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const uint8_t* __restrict__ in, const int* __restrict__ sizes, __nv_bfloat16* __restrict__ out,
                   int n, int hin, int win, int ho, int wo, Norm3 nm) {
  __shared__ float patch[3][ST_PH][ST_PW + 1];
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * ST_TH, ox0 = blockIdx.x * ST_TW;
  const int vh = sizes[2 * b], vw = sizes[2 * b + 1];
  const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
  for (int i = threadIdx.x; i < 3 * ST_PH * ST_PW; i += 256) {
    const int px = i % ST_PW;
    const int r = i / ST_PW;
    const int py = r % ST_PH, c = r / ST_PH;
    const int iy = iy0 + py, ix = ix0 + px;
    float v = 0.f;
    if (iy >= 0 && iy < vh && ix >= 0 && ix < vw)
      v = __fdiv_rn(__fsub_rn((float)__ldg(in + (((size_t)b * 3 + c) * hin + iy) * win + ix), nm.mean[c]), nm.stdv[c]);
    patch[c][py][px] = v;
  }
  __syncthreads();
  for (int v = threadIdx.x; v < ST_TH * ST_TW * 24; v += 256) {
    const int kv = v % 24;
    const int pix = v / 24;
    const int lx = pix % ST_TW, ly = pix / ST_TW;
    const int oy = oy0 + ly, ox = ox0 + lx;
    if (oy >= ho || ox >= wo) continue;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kv * 8 + e;
      float val = 0.f;
      if (k < 147) {
        const int tap = k / 3, c = k - tap * 3;
        const int rr = tap / 7, ss = tap - rr * 7;
        val = patch[c][ly * 2 + rr][lx * 2 + ss];
      }
      f[e] = val;
    }
    uint4 q;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    reinterpret_cast<uint4*>(out)[(((size_t)b * ho + oy) * wo + ox) * 24 + kv] = q;
  }
}

// ---- stem 7x7/2 conv without im2col: fused normalise + 2x2 space-to-depth of the uint8 image ---------------
// out: bf16 (N, Ho, Wo + 4, 16): pixel (i, jp) holds the 2x2 input block at rows 2i..2i+1, columns 2(jp-2)..2(jp-2)+1
// as channel (dy*2+dx)*4 + c (c == 3 and everything outside the valid image are zero; two zero columns on the left,
// two on the right).  With r+1 = 2a+dy, s+1 = 2b+dx the 7x7 stride-2 pad-3 conv becomes a 4x4 stride-1 conv over
// this map, and the four column taps b of one output pixel are 64 CONTIGUOUS values starting at column jp = ox: the
// conv kernel reads them as one 128-byte TMA row of an overlapping-row view (row stride 16 elements), row taps a by
// TMA coordinate, zero rows above/below by TMA out-of-bounds fill.
template <typename T>
__global__ void __launch_bounds__(256)
stem_s2d_kernel(const uint8_t* __restrict__ in, const int* __restrict__ sizes, T* __restrict__ out, int n,
                int hin, int win, int ho, int wp, Norm3 nm) {
  const size_t total = (size_t)n * ho * wp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int jp = (int)(i % wp);
    size_t r = i / wp;
    const int y = (int)(r % ho);
    const int b = (int)(r / ho);
    const int vh = sizes[2 * b], vw = sizes[2 * b + 1];
    const int x0 = 2 * (jp - 2), y0 = 2 * y;
    float f[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] = 0.f;
    if (x0 >= 0 && x0 < win) {
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int iy = y0 + dy;
        if (iy >= vh) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const uint8_t* p = in + (((size_t)b * 3 + c) * hin + iy) * win + x0;
          const uchar2 px = (x0 + 1 < win) ? *reinterpret_cast<const uchar2*>(p) : make_uchar2(*p, 0);
          if (x0 < vw) f[(dy * 2 + 0) * 4 + c] = __fdiv_rn(__fsub_rn((float)px.x, nm.mean[c]), nm.stdv[c]);
          if (x0 + 1 < vw) f[(dy * 2 + 1) * 4 + c] = __fdiv_rn(__fsub_rn((float)px.y, nm.mean[c]), nm.stdv[c]);
        }
      }
    }
    if constexpr (sizeof(T) == 4) {
      float4* o = reinterpret_cast<float4*>(out) + i * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = make_float4(f[4 * e], f[4 * e + 1], f[4 * e + 2], f[4 * e + 3]);
    } else {
      uint4 q[2];
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(q);
#pragma unroll
      for (int e = 0; e < 8; ++e) h2[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
      reinterpret_cast<uint4*>(out)[i * 2] = q[0];
      reinterpret_cast<uint4*>(out)[i * 2 + 1] = q[1];
    }
  }
}

// ---- split-bf16 parity mode ("bf16x3" / "bf16x6"): an fp32 tensor as a sum of 2 or 3 bf16 tensors ------------------
// part[0] = bf16(x), part[1] = bf16(x - part[0]), part[2] = bf16(x - part[0] - part[1]); the tcgen05 kernels then run
// once per kept product term (hi*hi, lo*hi, hi*lo, ...) into one fp32 accumulation: the SAME tensor-core main loops as
// the bf16 step, at 2^-16 (two parts) or fp32-level (three parts) relative accuracy, so the whole step can be held to
// the oracle's 1e-3 bar.  Source: strided channels-last view; parts: contiguous (n, h, w, c), `part_stride` elements apart.
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ x, long long sn, long long sh, long long sw, int n, int h, int w, int c,
                  __nv_bfloat16* __restrict__ out, long long part_stride, int parts) {
  const size_t total = (size_t)n * h * w * c;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    size_t r = i / c;
    const int xw = (int)(r % w);
    r /= w;
    const int y = (int)(r % h);
    const int b = (int)(r / h);
    float v = __ldg(x + (long long)b * sn + (long long)y * sh + (long long)xw * sw + ch);
    for (int p = 0; p < parts; ++p) {
      const __nv_bfloat16 q = __float2bfloat16_rn(v);
      out[(long long)p * part_stride + (long long)i] = q;
      v = __fsub_rn(v, __bfloat162float(q));
    }
  }
}

// The fused epilogue of aldi_conv_tc restated on an fp32 accumulation (the sum of the split-mode launches):
//   v = raw*scale[co] + bias[co] (+ residual) ; relu ; (* (mask > 0)) ; (+= out)
__global__ void __launch_bounds__(256)
conv_epilogue_f32_kernel(const float* __restrict__ raw, int n, int ho, int wo, int cp, const float* __restrict__ scale,
                         const float* __restrict__ bias, const float* __restrict__ residual, int res_mode, long long res_sn,
                         long long res_sh, long long res_sw, const float* __restrict__ mask, long long mask_sn,
                         long long mask_sh, long long mask_sw, int relu, int accumulate, float* __restrict__ out,
                         long long out_sn, long long out_sh, long long out_sw, int cout_store) {
  const size_t total = (size_t)n * ho * wo * cout_store;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout_store);
    size_t r = i / cout_store;
    const int xw = (int)(r % wo);
    r /= wo;
    const int y = (int)(r % ho);
    const int b = (int)(r / ho);
    float v = raw[(((size_t)b * ho + y) * wo + xw) * cp + co];
    if (scale) v = __fmul_rn(v, __ldg(scale + co));
    if (bias) v = __fadd_rn(v, __ldg(bias + co));
    if (res_mode == 1) v += residual[(long long)b * res_sn + (long long)y * res_sh + (long long)xw * res_sw + co];
    else if (res_mode == 2) v += residual[(long long)b * res_sn + (long long)(y >> 1) * res_sh + (long long)(xw >> 1) * res_sw + co];
    if (relu) v = fmaxf(v, 0.f);
    if (mask && !(mask[(long long)b * mask_sn + (long long)y * mask_sh + (long long)xw * mask_sw + co] > 0.f)) v = 0.f;
    float* o = out + (long long)b * out_sn + (long long)y * out_sh + (long long)xw * out_sw + co;
    *o = accumulate ? *o + v : v;
  }
}

// ---- max_pool2d(kernel 3, stride 2, padding 1) on channels-last, 16 bytes of channels per thread ------
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float* f) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* f) {
    uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* f) {
    uint4 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = q;
  }
};

template <typename T>
__global__ void __launch_bounds__(256)
maxpool_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int h, int w, int c, int ho, int wo) {
  constexpr int V = Vec16<T>::N;
  const int cv = c / V;
  const size_t total = (size_t)n * ho * wo * cv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cv) * V;
    size_t r = i / cv;
    const int ox = (int)(r % wo);
    r /= wo;
    const int oy = (int)(r % ho);
    const int b = (int)(r / ho);
    float m[V];
#pragma unroll
    for (int k = 0; k < V; ++k) m[k] = -INFINITY;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int iy = oy * 2 + dy;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int ix = ox * 2 + dx;
        if (ix < 0 || ix >= w) continue;
        float f[V];
        Vec16<T>::load(in + (((size_t)b * h + iy) * w + ix) * c + ch, f);
#pragma unroll
        for (int k = 0; k < V; ++k) m[k] = fmaxf(m[k], f[k]);
      }
    }
    Vec16<T>::store(out + (((size_t)b * ho + oy) * wo + ox) * c + ch, m);
  }
}

// ---- dst[n,h,w,c] += sum_{i,j in 0..1} src[n,2h+i,2w+j,c]  (backward of nearest-2x upsample + add) --
template <typename T>
__global__ void __launch_bounds__(256)
sum2x2_kernel(const T* __restrict__ src, T* __restrict__ dst, int n, int h, int w, int c) {
  constexpr int V = Vec16<T>::N;
  const int cv = c / V;
  const size_t total = (size_t)n * h * w * cv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cv) * V;
    size_t r = i / cv;
    const int x = (int)(r % w);
    r /= w;
    const int y = (int)(r % h);
    const int b = (int)(r / h);
    const size_t base = (((size_t)b * 2 * h + 2 * y) * 2 * w + 2 * x) * c + ch;
    const size_t row = (size_t)2 * w * c;
    float a0[V], a1[V], a2[V], a3[V], d[V];
    Vec16<T>::load(src + base, a0);
    Vec16<T>::load(src + base + c, a1);
    Vec16<T>::load(src + base + row, a2);
    Vec16<T>::load(src + base + row + c, a3);
    T* dp = dst + (((size_t)b * h + y) * w + x) * c + ch;
    Vec16<T>::load(dp, d);
#pragma unroll
    for (int k = 0; k < V; ++k) d[k] += a0[k] + a1[k] + a2[k] + a3[k];
    Vec16<T>::store(dp, d);
  }
}

// ---- dst += src (fp32 source, e.g. RoIAlign's atomic gradient buffer); n multiple of 8 --------------
template <typename T, bool ASSIGN>
__global__ void __launch_bounds__(256) add_f32_kernel(T* __restrict__ dst, const float* __restrict__ src, size_t n) {
  constexpr int V = Vec16<T>::N;
  const size_t nv = n / V;
  pdl_wait();
  pdl_launch_dependents();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x) {
    float d[V], s[V];
    if (!ASSIGN) Vec16<T>::load(dst + i * V, d);
#pragma unroll
    for (int k = 0; k < V; k += 4) Vec16<float>::load(src + i * V + k, s + k);
#pragma unroll
    for (int k = 0; k < V; ++k) d[k] = ASSIGN ? s[k] : d[k] + s[k];
    Vec16<T>::store(dst + i * V, d);
  }
  for (size_t i = nv * V + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = from_f32<T>(ASSIGN ? src[i] : to_f32<T>(dst[i]) + src[i]);
}

// ---- out[c] += scale * sum_rows x[row, c]   (bias gradients) ----------------------------------------
// grid (chunks of rows, channel groups of 32); block 32 x 8: lanes over channels (coalesced), 8 row lanes.
// out[ch] += scale * sum over (image, row) of x[img*img_stride + row*row_stride + ch]: 16-byte loads, one thread
// per 16 bytes of a row, 256 / (threads per row) rows per pass, register accumulation, one shared-memory reduction
// and one atomicAdd per channel per block.
constexpr int kColsumThreads = 512;
template <typename T>
__global__ void __launch_bounds__(kColsumThreads, 2)
colsum_kernel(const T* __restrict__ x, int n_img, long long rows, long long img_stride, long long row_stride, int c,
              float scale, float* __restrict__ out) {
  constexpr int V = Vec16<T>::N;
  __shared__ float part[kColsumThreads][V + 1];
  const int tpr = (c + V - 1) / V;            // threads per row
  const int rpp = kColsumThreads / tpr;       // rows per pass
  const int rg = threadIdx.x / tpr, tc = threadIdx.x - rg * tpr;
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  pdl_wait();
  pdl_launch_dependents();
  if (rg < rpp) {
    const long long total = (long long)n_img * rows;
    const long long step = (long long)gridDim.x * rpp;
    const T* base = x + (long long)tc * V;
    long long r = (long long)blockIdx.x * rpp + rg;
    // four independent 16-byte loads in flight per thread, 1024 threads per SM: a pure HBM stream.  Few, fat blocks:
    // the per-block atomics below all land on the same handful of cache lines and serialise in the L2.
    for (; r + 3 * step < total; r += 4 * step) {
      float f[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long ru = r + u * step;
        const long long img = (n_img == 1) ? 0 : ru / rows;
        Vec16<T>::load(base + img * img_stride + (ru - img * rows) * row_stride, f[u]);
      }
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += (f[0][k] + f[1][k]) + (f[2][k] + f[3][k]);
    }
    for (; r < total; r += step) {
      const long long img = (n_img == 1) ? 0 : r / rows;
      float f[V];
      Vec16<T>::load(base + img * img_stride + (r - img * rows) * row_stride, f);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += f[k];
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) part[threadIdx.x][k] = acc[k];
  __syncthreads();
  // thread (q, tc) of the first 4*tpr threads sums channel quad q of column group tc over the row groups
  constexpr int Q = V / 4;
  if (threadIdx.x < Q * tpr) {
    const int q = threadIdx.x / tpr, t2 = threadIdx.x - q * tpr;
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int g = 0; g < rpp; ++g) {
#pragma unroll
      for (int k = 0; k < 4; ++k) t[k] += part[g * tpr + t2][q * 4 + k];
    }
    const int ch = t2 * V + q * 4;
    if (ch + 4 <= c && ((reinterpret_cast<uintptr_t>(out + ch) & 15) == 0)) {
      atomicAdd(reinterpret_cast<float4*>(out + ch), make_float4(t[0] * scale, t[1] * scale, t[2] * scale, t[3] * scale));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ch + k < c) atomicAdd(out + ch + k, t[k] * scale);
    }
  }
}

// ---- backward of AdaptiveAvgPool2d(1) behind a ReLU (aldi/align.py:103-118 ConvDiscriminator) ---------------------
// dh[n, p, ch] = h[n, p, ch] > 0 ? dgap[n, ch] * scale : 0 and ndh = -dh (operand of the gradient-reversed dgrad)
template <typename T>
__global__ void __launch_bounds__(256)
gap_backward_kernel(const T* __restrict__ h, const float* __restrict__ dgap, int n, long long pix, int c_p, int c,
                    float scale, T* __restrict__ dh, T* __restrict__ ndh) {
  constexpr int V = Vec16<T>::N;
  const int cv = c_p / V;
  const size_t total = (size_t)n * pix * cv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cv) * V;
    const int img = (int)(i / ((size_t)pix * cv));
    float f[V], d[V], nd[V];
    Vec16<T>::load(h + i * V, f);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float g = (ch + k < c && f[k] > 0.f) ? __ldg(dgap + (size_t)img * c + ch + k) * scale : 0.f;
      d[k] = g;
      nd[k] = -g;
    }
    Vec16<T>::store(dh + i * V, d);
    Vec16<T>::store(ndh + i * V, nd);
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
  dim3 grid((wo + ST_TW - 1) / ST_TW, (ho + ST_TH - 1) / ST_TH, n);
  stem_im2col_kernel<<<grid, 256, 0, stream>>>(images, sizes, (__nv_bfloat16*)out_bf16, n, hin, win, ho, wo, nm);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_stem_im2col");
  return ALDI_OK;
}

extern "C" int aldi_stem_s2d(const uint8_t* images, const int* sizes, void* out_bf16, int n, int hin, int win,
                             const float* h_mean, const float* h_std, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(images && sizes && out_bf16 && h_mean && h_std, "aldi_stem_s2d: null pointer");
  ALDI_CHECK_ARG(n > 0 && hin % 2 == 0 && win % 2 == 0, "aldi_stem_s2d: canvas must have even height and width");
  Norm3 nm;
  for (int i = 0; i < 3; ++i) { nm.mean[i] = h_mean[i]; nm.stdv[i] = h_std[i]; }
  const int ho = hin / 2, wp = win / 2 + 4;
  stem_s2d_kernel<__nv_bfloat16><<<grid_for((size_t)n * ho * wp, 256), 256, 0, stream>>>(images, sizes, (__nv_bfloat16*)out_bf16,
                                                                                        n, hin, win, ho, wp, nm);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_stem_s2d");
  return ALDI_OK;
}

extern "C" int aldi_stem_s2d_f32(const uint8_t* images, const int* sizes, float* out, int n, int hin, int win,
                                 const float* h_mean, const float* h_std, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(images && sizes && out && h_mean && h_std, "aldi_stem_s2d_f32: null pointer");
  ALDI_CHECK_ARG(n > 0 && hin % 2 == 0 && win % 2 == 0, "aldi_stem_s2d_f32: canvas must have even height and width");
  Norm3 nm;
  for (int i = 0; i < 3; ++i) { nm.mean[i] = h_mean[i]; nm.stdv[i] = h_std[i]; }
  const int ho = hin / 2, wp = win / 2 + 4;
  stem_s2d_kernel<float><<<grid_for((size_t)n * ho * wp, 256), 256, 0, stream>>>(images, sizes, out, n, hin, win, ho, wp, nm);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_stem_s2d_f32");
  return ALDI_OK;
}

extern "C" int aldi_split_bf16(const float* x, int n, int h, int w, int c, long long sn, long long sh, long long sw,
                               void* out, long long part_stride, int parts, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(x && out, "aldi_split_bf16: null pointer");
  ALDI_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && parts >= 1 && parts <= 3, "aldi_split_bf16: bad extents / parts");
  split_bf16_kernel<<<grid_for((size_t)n * h * w * c, 256), 256, 0, stream>>>(x, sn, sh, sw, n, h, w, c, (__nv_bfloat16*)out,
                                                                             part_stride, parts);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_split_bf16");
  return ALDI_OK;
}

extern "C" int aldi_conv_epilogue_f32(const float* raw, const aldi_conv_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(raw && p && p->out, "aldi_conv_epilogue_f32: null pointer");
  ALDI_CHECK_ARG(p->out_dtype == ALDI_DTYPE_F32, "aldi_conv_epilogue_f32: fp32 output only");
  ALDI_CHECK_ARG(p->cout_store > 0 && p->cout_store <= p->cout_p, "aldi_conv_epilogue_f32: bad cout_store");
  ALDI_CHECK_ARG(!p->res_mode || p->residual, "aldi_conv_epilogue_f32: res_mode without residual");
  conv_epilogue_f32_kernel<<<grid_for((size_t)p->n * p->ho * p->wo * p->cout_store, 256), 256, 0, stream>>>(
      raw, p->n, p->ho, p->wo, p->cout_p, p->scale, p->bias, (const float*)p->residual, p->res_mode, p->res_sn, p->res_sh,
      p->res_sw, (const float*)p->mask, p->mask_sn, p->mask_sh, p->mask_sw, p->relu, p->accumulate, (float*)p->out,
      p->out_sn, p->out_sh, p->out_sw, p->cout_store);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_conv_epilogue_f32");
  return ALDI_OK;
}

extern "C" int aldi_maxpool3x3s2(const void* in, void* out, int dtype, int n, int h, int w, int c, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(in && out, "aldi_maxpool3x3s2: null pointer");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  ALDI_CHECK_ARG(c % 8 == 0, "aldi_maxpool3x3s2: channels must be a multiple of 8");
  const int grid = grid_for((size_t)n * ho * wo * c / 4, 256);
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
  ALDI_CHECK_ARG(c % 8 == 0, "aldi_sum2x2_accum: channels must be a multiple of 8");
  const int grid = grid_for((size_t)n * h * w * c / 4, 256);
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
  const int grid = grid_for(n / 4 + 1, 256);
  if (dtype == ALDI_DTYPE_BF16)
    aldi_launch_pdl(add_f32_kernel<__nv_bfloat16, false>, dim3(grid), dim3(256), 0, stream, (__nv_bfloat16*)dst, src, n);
  else
    aldi_launch_pdl(add_f32_kernel<float, false>, dim3(grid), dim3(256), 0, stream, (float*)dst, src, n);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_add_f32");
  return ALDI_OK;
}

extern "C" int aldi_cast_f32(void* dst, int dtype, const float* src, size_t n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(dst && src, "aldi_cast_f32: null pointer");
  if (n == 0) return ALDI_OK;
  const int grid = grid_for(n / 4 + 1, 256);
  if (dtype == ALDI_DTYPE_BF16)
    aldi_launch_pdl(add_f32_kernel<__nv_bfloat16, true>, dim3(grid), dim3(256), 0, stream, (__nv_bfloat16*)dst, src, n);
  else
    aldi_launch_pdl(add_f32_kernel<float, true>, dim3(grid), dim3(256), 0, stream, (float*)dst, src, n);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_cast_f32");
  return ALDI_OK;
}

extern "C" int aldi_colsum(const void* x, int dtype, int n_img, long long rows, long long img_stride,
                           long long row_stride, int c, float scale, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(x && out && c > 0 && n_img > 0, "aldi_colsum: bad args");
  const int v = dtype == ALDI_DTYPE_BF16 ? 8 : 4;
  ALDI_CHECK_ARG((c + v - 1) / v <= kColsumThreads, "aldi_colsum: at most %d channels", kColsumThreads * v);
  ALDI_CHECK_ARG(row_stride % v == 0 && img_stride % v == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
                 "aldi_colsum: rows must be 16-byte aligned (and hold ceil(c/%d)*%d readable channels)", v, v);
  if (rows <= 0) return ALDI_OK;
  const int rpp = kColsumThreads / ((c + v - 1) / v);
  long long gx = ((long long)n_img * rows + (long long)rpp * 16 - 1) / ((long long)rpp * 16);
  long long cap = (long long)aldi_num_sms() * 2;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  if (dtype == ALDI_DTYPE_BF16)
    aldi_launch_pdl(colsum_kernel<__nv_bfloat16>, dim3((unsigned)gx), dim3(kColsumThreads), 0, stream, (const __nv_bfloat16*)x, n_img,
                    rows, img_stride, row_stride, c, scale, out);
  else
    aldi_launch_pdl(colsum_kernel<float>, dim3((unsigned)gx), dim3(kColsumThreads), 0, stream, (const float*)x, n_img, rows,
                    img_stride, row_stride, c, scale, out);
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

extern "C" int aldi_gap_backward(const void* h, const float* dgap, int dtype, int n, long long pix, int c_p, int c,
                                 float scale, void* dh, void* ndh, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(h && dgap && dh && ndh && n > 0 && pix > 0 && c > 0 && c <= c_p, "aldi_gap_backward: bad args");
  const int v = dtype == ALDI_DTYPE_BF16 ? 8 : 4;
  ALDI_CHECK_ARG(c_p % v == 0, "aldi_gap_backward: padded channels must be a multiple of %d", v);
  const int grid = grid_for((size_t)n * pix * (c_p / v), 256);
  if (dtype == ALDI_DTYPE_BF16)
    gap_backward_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)h, dgap, n, pix, c_p, c, scale,
                                                                  (__nv_bfloat16*)dh, (__nv_bfloat16*)ndh);
  else
    gap_backward_kernel<float><<<grid, 256, 0, stream>>>((const float*)h, dgap, n, pix, c_p, c, scale, (float*)dh,
                                                         (float*)ndh);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_gap_backward");
  return ALDI_OK;
}
