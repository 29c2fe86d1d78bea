// Sparse backward of the RPN head (detectron2 StandardRPNHead: 3x3 conv 256->256 + ReLU, 1x1 -> objectness / deltas).
//
// The gradient the RPN losses send into the head is zero almost everywhere: detectron2's RPN.losses and
// aldi/distill.py:193-229 only touch the anchors picked by `subsample_labels` (256 per image, configs/detectron2/
// Base-RCNN-FPN.yaml) -- at 1024x2048 that is at most ~1000 of the 174,592 locations of an image.  The dense backward
// (what cuDNN does under Detectron2, and what this step did in round 1) still runs the 3x3 256->256 data- and
// weight-gradient convolutions over all five levels: 3.4 ms of a 31 ms step on a tensor that is > 99 % zeros.
//
// Here the non-zero rows of the head-output gradient are compacted, the operands of the two layers are GATHERED for
// those locations only (the ReLU'd hidden row t, and the 3x3 x 256 neighbourhood of the FPN feature), the four GEMMs
// run on the gathered rows with the same tcgen05 kernels (aldi_conv_tc / aldi_wgrad_tc treat them as a 1 x 1 x K x C
// image), and the feature gradient rows are SCATTER-added (fp32 atomics) into the fp32 map that the RoIAlign backward
// accumulates into as well.  Same sums, zeros skipped: parity with the oracle's dense autograd is held by the
// whole-step tests.
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

template <typename T> struct Row8;   // 8 consecutive channels
template <> struct Row8<float> {
  static __device__ __forceinline__ void load(const float* p, float* f) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
};
template <> struct Row8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* f) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
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

// ---- 1. compaction: rows of drpn (n * total_locs rows of `dstride` elements) with a non-zero among the first `ch` ----
template <typename T>
__global__ void __launch_bounds__(256)
compact_kernel(const T* __restrict__ drpn, long long rows, int dstride, int ch, int cap, int* __restrict__ idx,
               int* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  for (long long r0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) - lane; r0 < rows;
       r0 += (long long)gridDim.x * blockDim.x) {
    const long long r = r0 + lane;
    bool nz = false;
    if (r < rows) {
      const T* p = drpn + r * dstride;
      for (int c0 = 0; c0 < ch; c0 += 8) {
        float f[8];
        Row8<T>::load(p + c0, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) nz |= (c0 + k < ch) && (f[k] != 0.f);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, nz);
    if (m) {
      int base = 0;
      if (lane == 0) base = atomicAdd(count, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (nz) {
        const int pos = base + __popc(m & ((1u << lane) - 1));
        if (pos < cap) idx[pos] = (int)r;
      }
    }
  }
}

struct Levels {
  int num_levels, total_locs;
  int h[5], w[5], loc_off[5];
  const void* feat[5];   // FPN feature p_l (n, h, w, C) view: element strides below
  long long f_sn[5], f_sh[5], f_sw[5];
  const void* hid[5];    // hidden activation t_l (n, h, w, C) contiguous
  float* dfeat[5];       // fp32 gradient map the rows scatter into (p6 rows land in p5's map, strided)
  long long d_sn[5], d_sh[5], d_sw[5];
};

__device__ __forceinline__ bool decode(const Levels& L, int row, int* img, int* lvl, int* y, int* x) {
  if (row < 0) return false;
  *img = row / L.total_locs;
  const int loc = row - *img * L.total_locs;
  int l = 0;
  while (l + 1 < L.num_levels && loc >= L.loc_off[l + 1]) ++l;
  const int p = loc - L.loc_off[l];
  *lvl = l;
  *y = p / L.w[l];
  *x = p - *y * L.w[l];
  return true;
}

// ---- 2. gather: per compacted row k: dy_g[k] = drpn row; t_g[k] = hidden row; x_g[k][tap] = 3x3 neighbourhood of the feature
template <typename T>
__global__ void __launch_bounds__(256)
gather_kernel(const Levels L, const int* __restrict__ idx, const int* __restrict__ count, int cap, const T* __restrict__ drpn,
              int dstride, int C, T* __restrict__ dy_g, T* __restrict__ t_g, T* __restrict__ x_g, int* __restrict__ err_flag) {
  const int cv = C / 8;                        // 16-byte (bf16) / 32-byte (fp32) vectors per row
  const int per_row = 8 + cv + 9 * cv;         // dy (64 channels) + t + 9 taps
  const int cnt = min(*count, cap);
  if (blockIdx.x == 0 && threadIdx.x == 0 && *count > cap && err_flag) atomicOr(err_flag, 2);
  const size_t total = (size_t)cap * per_row;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / per_row);
    int v = (int)(i - (size_t)k * per_row);
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int img, l, y, x;
    const bool live = k < cnt && decode(L, idx[k], &img, &l, &y, &x);
    if (v < 8) {
      if (live) Row8<T>::load(drpn + (long long)idx[k] * dstride + v * 8, f);
      Row8<T>::store(dy_g + (size_t)k * 64 + v * 8, f);
      continue;
    }
    v -= 8;
    if (v < cv) {
      if (live) Row8<T>::load(reinterpret_cast<const T*>(L.hid[l]) + (((size_t)img * L.h[l] + y) * L.w[l] + x) * C + v * 8, f);
      Row8<T>::store(t_g + (size_t)k * C + v * 8, f);
      continue;
    }
    v -= cv;
    const int tap = v / cv, c8 = v - tap * cv;
    if (live) {
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy >= 0 && yy < L.h[l] && xx >= 0 && xx < L.w[l])
        Row8<T>::load(reinterpret_cast<const T*>(L.feat[l]) + img * L.f_sn[l] + yy * L.f_sh[l] + xx * L.f_sw[l] + c8 * 8, f);
    }
    Row8<T>::store(x_g + ((size_t)k * 9 + tap) * C + c8 * 8, f);
  }
}

// ---- 4. scatter: dfeat[img, y + r - 1, x + s - 1, :] += dx_g[k][tap = r * 3 + s][:]  (the conv's zero padding drops the rest)
template <typename T>
__global__ void __launch_bounds__(256)
scatter_kernel(const Levels L, const int* __restrict__ idx, const int* __restrict__ count, int cap, int C,
               const T* __restrict__ dx_g) {
  const int cv = C / 8;
  const int cnt = min(*count, cap);
  const size_t total = (size_t)cnt * 9 * cv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    size_t r = i / cv;
    const int tap = (int)(r % 9);
    const int k = (int)(r / 9);
    int img, l, y, x;
    if (!decode(L, idx[k], &img, &l, &y, &x)) continue;
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy < 0 || yy >= L.h[l] || xx < 0 || xx >= L.w[l]) continue;
    float f[8];
    Row8<T>::load(dx_g + ((size_t)k * 9 + tap) * C + c8 * 8, f);
    float* d = L.dfeat[l] + img * L.d_sn[l] + yy * L.d_sh[l] + xx * L.d_sw[l] + c8 * 8;
    atomicAdd(reinterpret_cast<float4*>(d), make_float4(f[0], f[1], f[2], f[3]));
    atomicAdd(reinterpret_cast<float4*>(d) + 1, make_float4(f[4], f[5], f[6], f[7]));
  }
}

int fill_levels(Levels* L, const aldi_rpn_sparse_params* p) {
  const aldi_rpn_levels* lv = p->levels;
  ALDI_CHECK_ARG(lv && lv->num_levels >= 1 && lv->num_levels <= 5, "aldi_rpn_sparse: bad levels");
  L->num_levels = lv->num_levels;
  L->total_locs = lv->total_locs;
  for (int i = 0; i < lv->num_levels; ++i) {
    L->h[i] = lv->h[i]; L->w[i] = lv->w[i]; L->loc_off[i] = lv->loc_off[i];
    L->feat[i] = p->feat[i]; L->f_sn[i] = p->feat_sn[i]; L->f_sh[i] = p->feat_sh[i]; L->f_sw[i] = p->feat_sw[i];
    L->hid[i] = p->hidden[i];
    L->dfeat[i] = p->dfeat[i]; L->d_sn[i] = p->dfeat_sn[i]; L->d_sh[i] = p->dfeat_sh[i]; L->d_sw[i] = p->dfeat_sw[i];
  }
  return ALDI_OK;
}

}  // namespace

extern "C" int aldi_rpn_sparse_compact(const void* drpn, int dtype, int n_images, int total_locs, int dstride, int channels,
                                       int cap, int* idx, int* count, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(drpn && idx && count, "aldi_rpn_sparse_compact: null pointer");
  ALDI_CHECK_ARG(n_images > 0 && total_locs > 0 && dstride >= 8 && dstride % 8 == 0 && channels > 0 && channels <= dstride &&
                     cap > 0 && (long long)n_images * total_locs < (1LL << 31), "aldi_rpn_sparse_compact: bad sizes");
  cudaError_t e = cudaMemsetAsync(idx, 0xff, (size_t)cap * sizeof(int), stream);       // -1 = no row
  if (e == cudaSuccess) e = cudaMemsetAsync(count, 0, sizeof(int), stream);
  if (e != cudaSuccess) {
    aldi_set_error("aldi_rpn_sparse_compact: memset failed: %s", cudaGetErrorString(e));
    return ALDI_ERR_CUDA;
  }
  const long long rows = (long long)n_images * total_locs;
  if (dtype == ALDI_DTYPE_BF16)
    compact_kernel<__nv_bfloat16><<<grid_for((size_t)rows, 256), 256, 0, stream>>>((const __nv_bfloat16*)drpn, rows, dstride,
                                                                                  channels, cap, idx, count);
  else
    compact_kernel<float><<<grid_for((size_t)rows, 256), 256, 0, stream>>>((const float*)drpn, rows, dstride, channels, cap,
                                                                          idx, count);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_rpn_sparse_compact");
  return ALDI_OK;
}

extern "C" int aldi_rpn_sparse_gather(const aldi_rpn_sparse_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(p && p->idx && p->count && p->drpn && p->dy_g && p->t_g && p->x_g, "aldi_rpn_sparse_gather: null pointer");
  ALDI_CHECK_ARG(p->channels % 8 == 0 && p->dstride == 64 && p->cap > 0, "aldi_rpn_sparse_gather: channels % 8, dstride 64");
  Levels L;
  int rc = fill_levels(&L, p);
  if (rc) return rc;
  const int per_row = 8 + 10 * (p->channels / 8);
  const int grid = grid_for((size_t)p->cap * per_row, 256);
  if (p->dtype == ALDI_DTYPE_BF16)
    gather_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(L, p->idx, p->count, p->cap, (const __nv_bfloat16*)p->drpn, p->dstride,
                                                          p->channels, (__nv_bfloat16*)p->dy_g, (__nv_bfloat16*)p->t_g,
                                                          (__nv_bfloat16*)p->x_g, p->err_flag);
  else
    gather_kernel<float><<<grid, 256, 0, stream>>>(L, p->idx, p->count, p->cap, (const float*)p->drpn, p->dstride, p->channels,
                                                  (float*)p->dy_g, (float*)p->t_g, (float*)p->x_g, p->err_flag);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_rpn_sparse_gather");
  return ALDI_OK;
}

extern "C" int aldi_rpn_sparse_scatter(const aldi_rpn_sparse_params* p, const void* dx_g, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(p && p->idx && p->count && dx_g, "aldi_rpn_sparse_scatter: null pointer");
  ALDI_CHECK_ARG(p->channels % 8 == 0 && p->cap > 0, "aldi_rpn_sparse_scatter: channels % 8");
  Levels L;
  int rc = fill_levels(&L, p);
  if (rc) return rc;
  for (int i = 0; i < L.num_levels; ++i)
    ALDI_CHECK_ARG(L.dfeat[i] && (reinterpret_cast<uintptr_t>(L.dfeat[i]) & 15) == 0 && L.d_sw[i] % 4 == 0,
                   "aldi_rpn_sparse_scatter: gradient maps must be 16-byte aligned");
  const int grid = grid_for((size_t)p->cap * 9 * (p->channels / 8), 256);
  if (p->dtype == ALDI_DTYPE_BF16)
    scatter_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(L, p->idx, p->count, p->cap, p->channels, (const __nv_bfloat16*)dx_g);
  else
    scatter_kernel<float><<<grid, 256, 0, stream>>>(L, p->idx, p->count, p->cap, p->channels, (const float*)dx_g);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_rpn_sparse_scatter");
  return ALDI_OK;
}
