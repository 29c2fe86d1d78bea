// ViTDet backbone kernels (BASELINE configs[2]; aldi/backbone.py:21-64 -> detectron2 modeling/backbone/vit.py, utils.py):
//   aldi_window_partition     window_partition / window_unpartition (zero padding tokens are written, never masked)
//   aldi_add_rows_bcast       x += get_abs_pos(pos_embed)            aldi_sum_over_batch: its gradient
//   aldi_bicubic_resize       get_abs_pos: F.interpolate(bicubic, align_corners=False) and its transpose
//   aldi_linear_resize_rows   get_rel_pos: F.interpolate(linear) of a relative-position table and its transpose
//   aldi_maxpool2x2(+_backward)  SimpleFeaturePyramid's scale-0.5 branch
//   aldi_attention_forward / _backward: dispatch (bf16 -> tcgen05 kernels in attn_tc.cu) and the CUDA-core fp32 kernels
//                             (parity mode; also the cross-check of the tensor-core path)
// Activations channels-last, dtype T in {float, bf16}; arithmetic fp32.
#include <math.h>

#include "common.cuh"
#include "../../include/aldi_b200.h"

int aldi_attention_forward_tc(const aldi_attn_params* p, cudaStream_t stream);
int aldi_attention_backward_tc(const aldi_attn_params* p, cudaStream_t stream);

namespace {

inline int blocks_for(long long items, int threads, int per_thread) {
  long long b = (items + (long long)threads * per_thread - 1) / ((long long)threads * per_thread);
  const long long cap = (long long)aldi_num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
window_partition_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int h, int w, int ws, int stride, int inverse) {
  // vectors of 8 elements; one thread per (window token, 8-channel group) of the WINDOW tensor
  const int nwh = (h + ws - 1) / ws, nww = (w + ws - 1) / ws;
  const int vec = stride / 8;
  const long long total = (long long)n * nwh * nww * ws * ws * vec;
  constexpr int kBytes = 8 * sizeof(T);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    long long r = i / vec;
    const int tx = (int)(r % ws); r /= ws;
    const int ty = (int)(r % ws); r /= ws;
    const int wx = (int)(r % nww); r /= nww;
    const int wy = (int)(r % nwh);
    const int img = (int)(r / nwh);
    const int y = wy * ws + ty, x = wx * ws + tx;
    const bool inside = y < h && x < w;
    const long long wi = (i / vec) * stride + v * 8;
    const long long xi = (((long long)img * h + y) * w + x) * stride + v * 8;
    if (!inverse) {
      if (inside) {
        if (kBytes == 16) *reinterpret_cast<uint4*>(out + wi) = *reinterpret_cast<const uint4*>(in + xi);
        else { *reinterpret_cast<uint4*>(out + wi) = *reinterpret_cast<const uint4*>(in + xi);
               *(reinterpret_cast<uint4*>(out + wi) + 1) = *(reinterpret_cast<const uint4*>(in + xi) + 1); }
      } else {
        const uint4 z = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(out + wi) = z;
        if (kBytes == 32) *(reinterpret_cast<uint4*>(out + wi) + 1) = z;
      }
    } else if (inside) {
      *reinterpret_cast<uint4*>(out + xi) = *reinterpret_cast<const uint4*>(in + wi);
      if (kBytes == 32) *(reinterpret_cast<uint4*>(out + xi) + 1) = *(reinterpret_cast<const uint4*>(in + wi) + 1);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
add_rows_bcast_kernel(T* __restrict__ x, const float* __restrict__ pos, long long per_image, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    x[i] = from_f32<T>(to_f32<T>(x[i]) + pos[i % per_image]);
}

template <typename T>
__global__ void __launch_bounds__(256)
sum_over_batch_kernel(const T* __restrict__ dx, int n, long long per_image, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < n; ++k) s += to_f32<T>(dx[(long long)k * per_image + i]);
    out[i] += s;
  }
}

// ATen UpSampleBicubic2d: cubic convolution coefficients with A = -0.75
__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }
__device__ __forceinline__ void cubic_coeffs(float t, float* c) {
  const float A = -0.75f;
  c[0] = cubic2(t + 1.f, A);
  c[1] = cubic1(t, A);
  c[2] = cubic1(1.f - t, A);
  c[3] = cubic2(2.f - t, A);
}

__global__ void __launch_bounds__(256)
bicubic_kernel(float* __restrict__ src, int sh, int sw, float* __restrict__ dst, int dh, int dw, int c, int dst_stride,
               int backward) {
  const float scale_h = (float)sh / (float)dh, scale_w = (float)sw / (float)dw;
  const long long total = (long long)dh * dw * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const int ox = (int)((i / c) % dw), oy = (int)(i / ((long long)c * dw));
    // area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=true): no clamp at zero
    const float fy = scale_h * ((float)oy + 0.5f) - 0.5f, fx = scale_w * ((float)ox + 0.5f) - 0.5f;
    const int iy = (int)floorf(fy), ix = (int)floorf(fx);
    float cy[4], cx[4];
    cubic_coeffs(fy - (float)iy, cy);
    cubic_coeffs(fx - (float)ix, cx);
    float* d = dst + ((long long)oy * dw + ox) * dst_stride + ch;
    if (!backward) {
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int yy = min(max(iy - 1 + a, 0), sh - 1);
        float row = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int xx = min(max(ix - 1 + b, 0), sw - 1);
          row += src[((long long)yy * sw + xx) * c + ch] * cx[b];
        }
        acc += row * cy[a];
      }
      *d = acc;
    } else {
      const float g = *d;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int yy = min(max(iy - 1 + a, 0), sh - 1);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int xx = min(max(ix - 1 + b, 0), sw - 1);
          atomicAdd(src + ((long long)yy * sw + xx) * c + ch, g * cy[a] * cx[b]);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
linear_rows_kernel(float* __restrict__ src, int rows_in, float* __restrict__ dst, int rows_out, int c, int backward) {
  const float scale = (float)rows_in / (float)rows_out;
  const long long total = (long long)rows_out * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c), o = (int)(i / c);
    float f = scale * ((float)o + 0.5f) - 0.5f;     // area_pixel_compute_source_index, align_corners=false: clamped at 0
    if (f < 0.f) f = 0.f;
    const int i0 = min((int)f, rows_in - 1), i1 = min(i0 + 1, rows_in - 1);
    const float l1 = f - (float)i0, l0 = 1.f - l1;
    if (!backward) {
      dst[i] = l0 * src[(long long)i0 * c + ch] + l1 * src[(long long)i1 * c + ch];
    } else {
      const float g = dst[i];
      atomicAdd(src + (long long)i0 * c + ch, l0 * g);
      atomicAdd(src + (long long)i1 * c + ch, l1 * g);
    }
  }
}

template <typename T, bool kBackward>
__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ out, int n, int h, int w, int stride) {
  const int ho = h / 2, wo = w / 2;
  const long long total = (long long)n * ho * wo * stride;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % stride);
    long long r = i / stride;
    const int ox = (int)(r % wo); r /= wo;
    const int oy = (int)(r % ho);
    const int img = (int)(r / ho);
    const long long base = (((long long)img * h + 2 * oy) * w + 2 * ox) * stride + ch;
    const long long offs[4] = {0, (long long)stride, (long long)w * stride, (long long)w * stride + stride};
    float best = to_f32<T>(x[base]);
    int arg = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float v = to_f32<T>(x[base + offs[k]]);
      if (v > best) { best = v; arg = k; }     // ATen max_pool2d: the first maximum in scan order keeps the index
    }
    if (!kBackward) {
      out[i] = from_f32<T>(best);
    } else {
      const T g = dy[i];
#pragma unroll
      for (int k = 0; k < 4; ++k) out[base + offs[k]] = (k == arg) ? g : from_f32<T>(0.f);
    }
  }
}

// GEMM layout rel[(b, q, h), c] <-> key-major rel_h[(b, h, kh), q], rel_w[(b, h, kw), q]; one block = 16 queries of one
// (b, h): rows cross shared memory so that both sides move contiguous runs (columns on the GEMM side, queries on the other)
constexpr int kRtQ = 16;
__global__ void __launch_bounds__(256)
relpos_transpose_kernel(float* __restrict__ rel, int rp, float* __restrict__ rel_h, float* __restrict__ rel_w, int gh, int gw,
                        int heads, int backward) {
  extern __shared__ float tile[];                 // [kRtQ][rp + 1]
  const int tn = gh * gw, q0 = blockIdx.x * kRtQ, h = blockIdx.y, b = blockIdx.z;
  const int nq = min(kRtQ, tn - q0), ld = rp + 1;
  const long long bh = (long long)b * heads + h;
  const int nh = 2 * gh - 1;
  if (!backward) {
    for (int i = threadIdx.x; i < nq * rp; i += blockDim.x) {
      const int qi = i / rp, c = i - qi * rp;
      tile[qi * ld + c] = rel[(((long long)b * tn + q0 + qi) * heads + h) * rp + c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (gh + gw) * kRtQ; i += blockDim.x) {
      const int k = i / kRtQ, qi = i % kRtQ;
      if (qi >= nq) continue;
      const int q = q0 + qi, qh = q / gw, qw = q - qh * gw;
      if (k < gh) rel_h[(bh * gh + k) * tn + q] = tile[qi * ld + gh - 1 + qh - k];
      else rel_w[(bh * gw + (k - gh)) * tn + q] = tile[qi * ld + nh + gw - 1 + qw - (k - gh)];
    }
  } else {
    for (int i = threadIdx.x; i < nq * rp; i += blockDim.x) tile[(i / rp) * ld + (i % rp)] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < (gh + gw) * kRtQ; i += blockDim.x) {
      const int k = i / kRtQ, qi = i % kRtQ;
      if (qi >= nq) continue;
      const int q = q0 + qi, qh = q / gw, qw = q - qh * gw;
      if (k < gh) tile[qi * ld + gh - 1 + qh - k] = rel_h[(bh * gh + k) * tn + q];
      else tile[qi * ld + nh + gw - 1 + qw - (k - gh)] = rel_w[(bh * gw + (k - gh)) * tn + q];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nq * rp; i += blockDim.x) {
      const int qi = i / rp, c = i - qi * rp;
      rel[(((long long)b * tn + q0 + qi) * heads + h) * rp + c] = tile[qi * ld + c];
    }
  }
}

// da == NULL: out = max(x, 0); else out = da * (x > 0)   (the ReLU after the LayerNorm of ViTDet's 4-conv box head)
template <typename T>
__global__ void __launch_bounds__(256)
relu_kernel(const T* __restrict__ x, const T* __restrict__ da, T* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = to_f32<T>(x[i]);
    out[i] = da ? (v > 0.f ? da[i] : from_f32<T>(0.f)) : from_f32<T>(fmaxf(v, 0.f));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// CUDA-core attention (fp32 arithmetic), head dim 64.  Tokens in raster order; keys in tiles of 32 staged in shared memory.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kHd = 64;
constexpr int kKeyTile = 32;

struct AttnArgs {
  const void* qkv; const void* out; const void* dout;
  void* out_w; void* dqkv;
  const float* rel_h; const float* rel_w; float* drel_h; float* drel_w; float* lse; float* delta;
  long long row_stride, batch_stride, out_stride, out_batch_stride;
  int gh, gw, heads, dim;
  float scale;
};

template <typename T>
__device__ __forceinline__ void load_row64(const T* p, float* f) {
#pragma unroll
  for (int d = 0; d < kHd; ++d) f[d] = to_f32<T>(p[d]);
}

template <typename T>
__device__ __forceinline__ void stage_tile(const T* base, long long row_stride, int k0, int tn, int col, float (*dst)[kHd],
                                           int tid, int nthreads, int rows) {
  for (int i = tid; i < rows * kHd; i += nthreads) {
    const int key = i / kHd, d = i % kHd;
    const int kk = k0 + key;
    dst[key][d] = kk < tn ? to_f32<T>(base[(long long)kk * row_stride + col + d]) : 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(128)
attn_fwd_simt(const AttnArgs a) {
  __shared__ float sK[kKeyTile][kHd];
  __shared__ float sV[kKeyTile][kHd];
  const int tid = threadIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tn = a.gh * a.gw;
  const int q = blockIdx.x * 128 + tid;
  const bool valid = q < tn;
  const int qc = valid ? q : tn - 1;
  const T* base = reinterpret_cast<const T*>(a.qkv) + (long long)b * a.batch_stride;
  float qv[kHd], o[kHd];
  load_row64<T>(base + (long long)qc * a.row_stride + h * kHd, qv);
#pragma unroll
  for (int d = 0; d < kHd; ++d) { qv[d] *= a.scale; o[d] = 0.f; }
  // key-major relative-position products: rel_h[(b, h, kh), q], rel_w[(b, h, kw), q]
  const long long bh = (long long)b * a.heads + h;
  const float* rh = a.rel_h ? a.rel_h + bh * a.gh * tn + qc : nullptr;
  const float* rw = a.rel_h ? a.rel_w + bh * a.gw * tn + qc : nullptr;
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < tn; k0 += kKeyTile) {
    __syncthreads();
    stage_tile<T>(base, a.row_stride, k0, tn, a.dim + h * kHd, sK, tid, 128, kKeyTile);
    stage_tile<T>(base, a.row_stride, k0, tn, 2 * a.dim + h * kHd, sV, tid, 128, kKeyTile);
    __syncthreads();
    float s[kKeyTile];
    float mt = -INFINITY;
#pragma unroll
    for (int j = 0; j < kKeyTile; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < kHd; ++d) acc += qv[d] * sK[j][d];
      const int kk = k0 + j;
      if (kk < tn) {
        if (rh) {
          const int kh = kk / a.gw, kw = kk - kh * a.gw;
          acc += rh[(long long)kh * tn] + rw[(long long)kw * tn];
        }
      } else {
        acc = -INFINITY;
      }
      s[j] = acc;
      mt = fmaxf(mt, acc);
    }
    const float mn = fmaxf(m, mt);
    const float alpha = __expf(m - mn);
    l *= alpha;
#pragma unroll
    for (int d = 0; d < kHd; ++d) o[d] *= alpha;
#pragma unroll
    for (int j = 0; j < kKeyTile; ++j) {
      const float p = __expf(s[j] - mn);
      l += p;
#pragma unroll
      for (int d = 0; d < kHd; ++d) o[d] += p * sV[j][d];
    }
    m = mn;
  }
  if (valid) {
    const float inv = 1.f / l;
    T* orow = reinterpret_cast<T*>(a.out_w) + (long long)b * a.out_batch_stride + (long long)q * a.out_stride + h * kHd;
#pragma unroll
    for (int d = 0; d < kHd; ++d) orow[d] = from_f32<T>(o[d] * inv);
    a.lse[((long long)b * a.heads + h) * tn + q] = m + logf(l);
  }
}

// thread per query: dq, drelpos row, delta
template <typename T>
__global__ void __launch_bounds__(128)
attn_bwd_dq_simt(const AttnArgs a) {
  __shared__ float sK[kKeyTile][kHd];
  __shared__ float sV[kKeyTile][kHd];
  const int tid = threadIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tn = a.gh * a.gw;
  const int q = blockIdx.x * 128 + tid;
  const bool valid = q < tn;
  const int qc = valid ? q : tn - 1;
  const T* base = reinterpret_cast<const T*>(a.qkv) + (long long)b * a.batch_stride;
  float qv[kHd], dov[kHd], dq[kHd];
  load_row64<T>(base + (long long)qc * a.row_stride + h * kHd, qv);
  const long long orow = (long long)b * a.out_batch_stride + (long long)qc * a.out_stride + h * kHd;
  load_row64<T>(reinterpret_cast<const T*>(a.dout) + orow, dov);
  float delta = 0.f;
  {
    const T* op = reinterpret_cast<const T*>(a.out) + orow;
#pragma unroll
    for (int d = 0; d < kHd; ++d) delta += dov[d] * to_f32<T>(op[d]);
  }
#pragma unroll
  for (int d = 0; d < kHd; ++d) { qv[d] *= a.scale; dq[d] = 0.f; }
  const long long stat = ((long long)b * a.heads + h) * tn + qc;
  const float lse = a.lse[stat];
  if (valid) a.delta[stat] = delta;
  const long long bh = (long long)b * a.heads + h;
  const float* rh = a.rel_h ? a.rel_h + bh * a.gh * tn + qc : nullptr;
  const float* rw = a.rel_h ? a.rel_w + bh * a.gw * tn + qc : nullptr;
  float* dh = (a.rel_h && valid) ? a.drel_h + bh * a.gh * tn + qc : nullptr;
  float* dw = (a.rel_h && valid) ? a.drel_w + bh * a.gw * tn + qc : nullptr;
  if (dh) {       // this thread owns column q of both gradient arrays
    for (int k = 0; k < a.gh; ++k) dh[(long long)k * tn] = 0.f;
    for (int k = 0; k < a.gw; ++k) dw[(long long)k * tn] = 0.f;
  }
  for (int k0 = 0; k0 < tn; k0 += kKeyTile) {
    __syncthreads();
    stage_tile<T>(base, a.row_stride, k0, tn, a.dim + h * kHd, sK, tid, 128, kKeyTile);
    stage_tile<T>(base, a.row_stride, k0, tn, 2 * a.dim + h * kHd, sV, tid, 128, kKeyTile);
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < kKeyTile; ++j) {
      const int kk = k0 + j;
      if (kk >= tn) break;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < kHd; ++d) { s += qv[d] * sK[j][d]; dp += dov[d] * sV[j][d]; }
      long long ch = 0, cw = 0;
      if (rh) {
        const int kh = kk / a.gw, kw = kk - kh * a.gw;
        ch = (long long)kh * tn;
        cw = (long long)kw * tn;
        s += rh[ch] + rw[cw];
      }
      const float p = __expf(s - lse);
      const float ds = p * (dp - delta);
#pragma unroll
      for (int d = 0; d < kHd; ++d) dq[d] += ds * sK[j][d];
      if (dh) { dh[ch] += ds; dw[cw] += ds; }     // this thread owns the column: plain read-modify-write
    }
  }
  if (valid) {
    T* dqrow = reinterpret_cast<T*>(a.dqkv) + (long long)b * a.batch_stride + (long long)q * a.row_stride + h * kHd;
#pragma unroll
    for (int d = 0; d < kHd; ++d) dqrow[d] = from_f32<T>(dq[d] * a.scale);
  }
}

// two threads per key (each owns 32 of the 64 channels): dk, dv; queries in tiles of 32 staged in shared memory
template <typename T>
__global__ void __launch_bounds__(128)
attn_bwd_dkv_simt(const AttnArgs a) {
  __shared__ float sQ[kKeyTile][kHd];
  __shared__ float sDO[kKeyTile][kHd];
  __shared__ float sLse[kKeyTile], sDelta[kKeyTile];
  const int tid = threadIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tn = a.gh * a.gw;
  const int key = blockIdx.x * 64 + (tid >> 1), half = tid & 1;
  const bool valid = key < tn;
  const int kc = valid ? key : tn - 1;
  const int kh = kc / a.gw, kw = kc % a.gw;
  const T* base = reinterpret_cast<const T*>(a.qkv) + (long long)b * a.batch_stride;
  const T* dob = reinterpret_cast<const T*>(a.dout) + (long long)b * a.out_batch_stride;
  float kv[32], vv[32], dk[32], dv[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    kv[d] = to_f32<T>(base[(long long)kc * a.row_stride + a.dim + h * kHd + half * 32 + d]);
    vv[d] = to_f32<T>(base[(long long)kc * a.row_stride + 2 * a.dim + h * kHd + half * 32 + d]);
    dk[d] = 0.f;
    dv[d] = 0.f;
  }
  for (int q0 = 0; q0 < tn; q0 += kKeyTile) {
    __syncthreads();
    stage_tile<T>(base, a.row_stride, q0, tn, h * kHd, sQ, tid, 128, kKeyTile);
    stage_tile<T>(dob, a.out_stride, q0, tn, h * kHd, sDO, tid, 128, kKeyTile);
    if (tid < kKeyTile) {
      const int qq = q0 + tid;
      const long long stat = ((long long)b * a.heads + h) * tn + (qq < tn ? qq : tn - 1);
      sLse[tid] = a.lse[stat];
      sDelta[tid] = a.delta[stat];
    }
    __syncthreads();
    const int lim = min(kKeyTile, tn - q0);
    for (int i = 0; i < lim; ++i) {
      const int qq = q0 + i;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) { s += sQ[i][half * 32 + d] * kv[d]; dp += sDO[i][half * 32 + d] * vv[d]; }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      dp += __shfl_xor_sync(0xffffffffu, dp, 1);
      s *= a.scale;
      if (a.rel_h) {
        const long long bh = (long long)b * a.heads + h;
        s += a.rel_h[(bh * a.gh + kh) * tn + qq] + a.rel_w[(bh * a.gw + kw) * tn + qq];
      }
      const float p = __expf(s - sLse[i]);
      const float ds = p * (dp - sDelta[i]);
#pragma unroll
      for (int d = 0; d < 32; ++d) { dv[d] += p * sDO[i][half * 32 + d]; dk[d] += ds * sQ[i][half * 32 + d]; }
    }
  }
  if (valid) {
    T* drow = reinterpret_cast<T*>(a.dqkv) + (long long)b * a.batch_stride + (long long)key * a.row_stride + h * kHd + half * 32;
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      drow[a.dim + d] = from_f32<T>(dk[d] * a.scale);
      drow[2 * a.dim + d] = from_f32<T>(dv[d]);
    }
  }
}

AttnArgs make_args(const aldi_attn_params* p) {
  AttnArgs a;
  a.qkv = p->qkv; a.out = p->out; a.out_w = p->out; a.dout = p->dout; a.dqkv = p->dqkv;
  a.rel_h = p->rel_h; a.rel_w = p->rel_w; a.drel_h = p->drel_h; a.drel_w = p->drel_w; a.lse = p->lse; a.delta = p->delta;
  a.row_stride = p->row_stride; a.batch_stride = p->batch_stride;
  a.out_stride = p->out_stride; a.out_batch_stride = p->out_batch_stride;
  a.gh = p->gh; a.gw = p->gw; a.heads = p->heads; a.dim = p->heads * kHd;
  a.scale = p->scale;
  return a;
}

int check_attn(const aldi_attn_params* p, bool backward, const char* who) {
  ALDI_CHECK_ARG(p && p->qkv && p->out && p->lse, "%s: null pointer", who);
  ALDI_CHECK_ARG(p->batch > 0 && p->gh > 0 && p->gw > 0 && p->heads > 0, "%s: empty problem", who);
  ALDI_CHECK_ARG(p->row_stride >= 3LL * p->heads * 64 && p->out_stride >= (long long)p->heads * 64, "%s: row strides too small", who);
  ALDI_CHECK_ARG(p->dtype == ALDI_F32 || p->dtype == ALDI_BF16, "%s: dtype %d", who, p->dtype);
  ALDI_CHECK_ARG((p->rel_h == nullptr) == (p->rel_w == nullptr), "%s: rel_h and rel_w come together", who);
  if (backward) {
    ALDI_CHECK_ARG(p->dout && p->dqkv && p->delta, "%s: null gradient pointer", who);
    ALDI_CHECK_ARG(!p->rel_h || (p->drel_h && p->drel_w), "%s: rel_h / rel_w without drel_h / drel_w", who);
  }
  return ALDI_OK;
}

}  // namespace

#define VIT_DISPATCH(dtype, F32CALL, BF16CALL, who)            \
  do {                                                         \
    if ((dtype) == ALDI_F32) { F32CALL; }                      \
    else if ((dtype) == ALDI_BF16) { BF16CALL; }               \
    else { aldi_set_error("%s: dtype %d", who, (int)(dtype)); return ALDI_ERR_INVALID; } \
    ALDI_COUNT_LAUNCH();                                       \
    ALDI_CUDA_LAUNCH_CHECK(who);                               \
  } while (0)

extern "C" int aldi_window_partition(const void* in, void* out, int n, int h, int w, int ws, int stride, int dtype, int inverse,
                                     void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(in && out && n > 0 && h > 0 && w > 0 && ws > 0 && stride > 0 && stride % 8 == 0, "aldi_window_partition: bad args");
  const long long items = (long long)n * ((h + ws - 1) / ws) * ((w + ws - 1) / ws) * ws * ws * (stride / 8);
  const int grid = blocks_for(items, 256, 4);
  VIT_DISPATCH(dtype,
               (window_partition_kernel<float><<<grid, 256, 0, stream>>>((const float*)in, (float*)out, n, h, w, ws, stride, inverse)),
               (window_partition_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, n, h, w,
                                                                               ws, stride, inverse)),
               "aldi_window_partition");
  return ALDI_OK;
}

extern "C" int aldi_add_rows_bcast(void* x, const float* pos, int n, long long rows_per_image, int stride, int dtype, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(x && pos && n > 0 && rows_per_image > 0 && stride > 0, "aldi_add_rows_bcast: bad args");
  const long long per = rows_per_image * stride, total = per * n;
  const int grid = blocks_for(total, 256, 8);
  VIT_DISPATCH(dtype, (add_rows_bcast_kernel<float><<<grid, 256, 0, stream>>>((float*)x, pos, per, total)),
               (add_rows_bcast_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((__nv_bfloat16*)x, pos, per, total)),
               "aldi_add_rows_bcast");
  return ALDI_OK;
}

extern "C" int aldi_sum_over_batch(const void* dx, int n, long long rows_per_image, int stride, int dtype, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(dx && out && n > 0 && rows_per_image > 0 && stride > 0, "aldi_sum_over_batch: bad args");
  const long long per = rows_per_image * stride;
  const int grid = blocks_for(per, 256, 2);
  VIT_DISPATCH(dtype, (sum_over_batch_kernel<float><<<grid, 256, 0, stream>>>((const float*)dx, n, per, out)),
               (sum_over_batch_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)dx, n, per, out)),
               "aldi_sum_over_batch");
  return ALDI_OK;
}

extern "C" int aldi_bicubic_resize(float* src, int sh, int sw, float* dst, int dh, int dw, int c, int dst_stride, int backward,
                                   void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(src && dst && sh > 0 && sw > 0 && dh > 0 && dw > 0 && c > 0 && dst_stride >= c, "aldi_bicubic_resize: bad args");
  bicubic_kernel<<<blocks_for((long long)dh * dw * c, 256, 2), 256, 0, stream>>>(src, sh, sw, dst, dh, dw, c, dst_stride, backward);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_bicubic_resize");
  return ALDI_OK;
}

extern "C" int aldi_linear_resize_rows(float* src, int rows_in, float* dst, int rows_out, int c, int backward, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(src && dst && rows_in > 0 && rows_out > 0 && c > 0, "aldi_linear_resize_rows: bad args");
  linear_rows_kernel<<<blocks_for((long long)rows_out * c, 256, 2), 256, 0, stream>>>(src, rows_in, dst, rows_out, c, backward);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_linear_resize_rows");
  return ALDI_OK;
}

extern "C" int aldi_maxpool2x2(const void* x, void* out, int dtype, int n, int h, int w, int stride, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(x && out && n > 0 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0 && stride > 0, "aldi_maxpool2x2: bad args");
  const int grid = blocks_for((long long)n * (h / 2) * (w / 2) * stride, 256, 4);
  VIT_DISPATCH(dtype,
               (maxpool2x2_kernel<float, false><<<grid, 256, 0, stream>>>((const float*)x, nullptr, (float*)out, n, h, w, stride)),
               (maxpool2x2_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, nullptr, (__nv_bfloat16*)out,
                                                                                n, h, w, stride)),
               "aldi_maxpool2x2");
  return ALDI_OK;
}

extern "C" int aldi_maxpool2x2_backward(const void* x, const void* dy, void* dx, int dtype, int n, int h, int w, int stride,
                                        void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(x && dy && dx && n > 0 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0 && stride > 0,
                 "aldi_maxpool2x2_backward: bad args");
  const int grid = blocks_for((long long)n * (h / 2) * (w / 2) * stride, 256, 4);
  VIT_DISPATCH(dtype,
               (maxpool2x2_kernel<float, true><<<grid, 256, 0, stream>>>((const float*)x, (const float*)dy, (float*)dx, n, h, w, stride)),
               (maxpool2x2_kernel<__nv_bfloat16, true><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy,
                                                                               (__nv_bfloat16*)dx, n, h, w, stride)),
               "aldi_maxpool2x2_backward");
  return ALDI_OK;
}

extern "C" int aldi_relu(const void* x, const void* da, void* out, size_t n, int dtype, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(x && out, "aldi_relu: null pointer");
  if (n == 0) return ALDI_OK;
  const int grid = blocks_for((long long)n, 256, 8);
  VIT_DISPATCH(dtype, (relu_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, (const float*)da, (float*)out, n)),
               (relu_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)da,
                                                                    (__nv_bfloat16*)out, n)),
               "aldi_relu");
  return ALDI_OK;
}

extern "C" int aldi_relpos_transpose(float* rel, int rp_stride, float* rel_h, float* rel_w, int batch, int gh, int gw, int heads,
                                     int backward, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(rel && rel_h && rel_w && batch > 0 && gh > 0 && gw > 0 && heads > 0, "aldi_relpos_transpose: bad args");
  ALDI_CHECK_ARG(rp_stride >= 2 * gh - 1 + 2 * gw - 1, "aldi_relpos_transpose: rp_stride %d < %d table columns", rp_stride,
                 2 * gh - 1 + 2 * gw - 1);
  ALDI_CHECK_ARG(batch <= 65535 && heads <= 65535, "aldi_relpos_transpose: batch / heads exceed the grid limits");
  const size_t smem = (size_t)kRtQ * (rp_stride + 1) * sizeof(float);
  ALDI_CHECK_ARG(smem <= 48 * 1024, "aldi_relpos_transpose: rp_stride %d too wide", rp_stride);
  const dim3 grid((gh * gw + kRtQ - 1) / kRtQ, heads, batch);
  relpos_transpose_kernel<<<grid, 256, smem, stream>>>(rel, rp_stride, rel_h, rel_w, gh, gw, heads, backward);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_relpos_transpose");
  return ALDI_OK;
}

extern "C" int aldi_attention_forward(const aldi_attn_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = check_attn(p, false, "aldi_attention_forward");
  if (rc) return rc;
  if (p->dtype == ALDI_BF16 && !p->impl) return aldi_attention_forward_tc(p, stream);
  const AttnArgs a = make_args(p);
  const dim3 grid((p->gh * p->gw + 127) / 128, p->heads, p->batch);
  VIT_DISPATCH(p->dtype, (attn_fwd_simt<float><<<grid, 128, 0, stream>>>(a)),
               (attn_fwd_simt<__nv_bfloat16><<<grid, 128, 0, stream>>>(a)), "aldi_attention_forward");
  return ALDI_OK;
}

extern "C" int aldi_attention_backward(const aldi_attn_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = check_attn(p, true, "aldi_attention_backward");
  if (rc) return rc;
  if (p->dtype == ALDI_BF16 && !p->impl) return aldi_attention_backward_tc(p, stream);
  const AttnArgs a = make_args(p);
  const int tn = p->gh * p->gw;
  const dim3 gq((tn + 127) / 128, p->heads, p->batch), gk((tn + 63) / 64, p->heads, p->batch);
  VIT_DISPATCH(p->dtype, (attn_bwd_dq_simt<float><<<gq, 128, 0, stream>>>(a)),
               (attn_bwd_dq_simt<__nv_bfloat16><<<gq, 128, 0, stream>>>(a)), "aldi_attention_backward(dq)");
  VIT_DISPATCH(p->dtype, (attn_bwd_dkv_simt<float><<<gk, 128, 0, stream>>>(a)),
               (attn_bwd_dkv_simt<__nv_bfloat16><<<gk, 128, 0, stream>>>(a)), "aldi_attention_backward(dkv)");
  return ALDI_OK;
}
