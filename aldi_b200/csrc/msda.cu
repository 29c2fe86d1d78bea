// Multi-scale deformable attention, forward and backward (Deformable-DETR, BASELINE configs[3]).
//
// Replaces the reference's MSDeformAttn CUDA extension
// (aldi/detr/libs/DeformableDETRDetectron2/deformable_detr/models/ops/src/cuda/ms_deform_attn_cuda.cu:29-153 and
// ms_deform_im2col_cuda.cuh) behind the same operator contract as functions/ms_deform_attn_func.py:21-38.
// Semantics (pinned by tests/golden/msda_golden.pt through oracle/msda_ref.py): pixel = loc * size - 0.5, bilinear
// taps outside the map read 0, out[n,q,m,:] = sum_{l,p} attn * bilinear(value_l[n,:,m,:]).
//
// Layout is the reference's: value (N, S, M, D) so the D channels of one head at one pixel are contiguous.  The
// work item is one (n, q, m) HEAD of one query; a group of G = min(32, pow2ceil(D)) lanes owns it and strides over
// the channels, so every bilinear tap is one contiguous D*sizeof(T) read (D = 32 fp32: exactly one 128-byte line
// per tap per warp) and 32/G heads share a warp when D is small.  Sampling locations / weights of a head are read
// once by its group's lanes (broadcast).  HBM/L2-gather bound: algorithmic bytes per head =
// L*P*(4 taps * D + 3) * sizeof(T) read + D * sizeof(T) written.
//
// Backward: grad_value is scattered with atomics (taps of different queries collide); grad_attn / grad_loc are
// per-(head, level, point) sums over the channels — a segmented warp-shuffle reduction inside the group, one store
// each, no atomics and no shared memory (the reference reduces through shared memory with a serial thread-0 sum, or
// global atomics when D is not a power of two, ms_deform_im2col_cuda.cuh:300-900).
#include "common.cuh"
#include "../../include/aldi_b200.h"

namespace {

constexpr int kMaxLevels = 8;

template <typename T>
struct MsdaArgs {
  const T* value;
  const T* loc;
  const T* attn;
  int n, s, m, d, lq, l, p;
  int h[kMaxLevels], w[kMaxLevels], start[kMaxLevels];
  T* out;
  const T* grad_out;
  T* grad_value;
  T* grad_loc;
  T* grad_attn;
  int group;       // lanes per head (power of two <= 32)
  long long heads; // n * lq * m
};

template <typename T>
struct Bilinear {
  int off[4];   // element offsets of the 4 taps inside value_l[n, :, m, :] (in units of one pixel row of M*D), -1 = outside
  T wt[4];      // bilinear weights
  T dwy[4], dwx[4];
  bool any;
};

template <typename T>
__device__ __forceinline__ Bilinear<T> make_taps(T lx, T ly, int h, int w) {
  Bilinear<T> b;
  const T x = lx * (T)w - (T)0.5, y = ly * (T)h - (T)0.5;
  b.any = (y > (T)-1) && (x > (T)-1) && (y < (T)h) && (x < (T)w);
  const T fy = floor(y), fx = floor(x);
  const int y0 = (int)fy, x0 = (int)fx;
  const T ry = y - fy, rx = x - fx, qy = (T)1 - ry, qx = (T)1 - rx;
  const bool y0ok = y0 >= 0, y1ok = y0 + 1 <= h - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= w - 1;
  b.off[0] = (b.any && y0ok && x0ok) ? y0 * w + x0 : -1;
  b.off[1] = (b.any && y0ok && x1ok) ? y0 * w + x0 + 1 : -1;
  b.off[2] = (b.any && y1ok && x0ok) ? (y0 + 1) * w + x0 : -1;
  b.off[3] = (b.any && y1ok && x1ok) ? (y0 + 1) * w + x0 + 1 : -1;
  b.wt[0] = qy * qx; b.wt[1] = qy * rx; b.wt[2] = ry * qx; b.wt[3] = ry * rx;
  b.dwy[0] = -qx; b.dwy[1] = -rx; b.dwy[2] = qx; b.dwy[3] = rx;
  b.dwx[0] = -qy; b.dwx[1] = qy; b.dwx[2] = -ry; b.dwx[3] = ry;
  return b;
}

template <typename T>
__global__ void __launch_bounds__(256) msda_forward_kernel(const MsdaArgs<T> a) {
  const int G = a.group;
  const int lane_in_group = threadIdx.x & (G - 1);
  const long long group_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const long long group_stride = (long long)gridDim.x * blockDim.x / G;
  const int md = a.m * a.d;
  for (long long head = group_global; head < a.heads; head += group_stride) {
    const int m = (int)(head % a.m);
    const long long nq = head / a.m;
    const int n = (int)(nq / a.lq);
    const T* locp = a.loc + head * a.l * a.p * 2;
    const T* attp = a.attn + head * a.l * a.p;
    for (int d0 = lane_in_group; d0 < a.d; d0 += G) {
      T acc = (T)0;
      for (int l = 0; l < a.l; ++l) {
        const T* vbase = a.value + ((long long)n * a.s + a.start[l]) * md + m * a.d + d0;
        for (int p = 0; p < a.p; ++p) {
          const Bilinear<T> b = make_taps<T>(locp[(l * a.p + p) * 2], locp[(l * a.p + p) * 2 + 1], a.h[l], a.w[l]);
          if (!b.any) continue;
          T v = (T)0;
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (b.off[t] >= 0) v += b.wt[t] * __ldg(vbase + (long long)b.off[t] * md);
          acc += attp[l * a.p + p] * v;
        }
      }
      a.out[head * a.d + d0] = acc;
    }
  }
}

template <typename T>
__device__ __forceinline__ T group_sum(T v, int G) {
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256) msda_backward_kernel(const MsdaArgs<T> a) {
  const int G = a.group;
  const int lane_in_group = threadIdx.x & (G - 1);
  const long long group_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const long long group_stride = (long long)gridDim.x * blockDim.x / G;
  const int md = a.m * a.d;
  // every lane of a warp runs the same number of iterations (shuffles need all 32 lanes): pad the head loop
  const long long iters = (a.heads + group_stride - 1) / group_stride;
  for (long long it = 0; it < iters; ++it) {
    const long long head = group_global + it * group_stride;
    const bool live = head < a.heads;
    const long long hd = live ? head : 0;
    const int m = (int)(hd % a.m);
    const int n = (int)((hd / a.m) / a.lq);
    const T* locp = a.loc + hd * a.l * a.p * 2;
    const T* attp = a.attn + hd * a.l * a.p;
    for (int l = 0; l < a.l; ++l) {
      const long long vrow = ((long long)n * a.s + a.start[l]) * md + m * a.d;
      for (int p = 0; p < a.p; ++p) {
        const Bilinear<T> b = make_taps<T>(locp[(l * a.p + p) * 2], locp[(l * a.p + p) * 2 + 1], a.h[l], a.w[l]);
        const T aw = attp[l * a.p + p];
        T g_attn = (T)0, g_y = (T)0, g_x = (T)0;
        if (live && b.any) {
          for (int d0 = lane_in_group; d0 < a.d; d0 += G) {
            const T go = a.grad_out[hd * a.d + d0];
            const T gv = go * aw;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (b.off[t] < 0) continue;
              const long long o = vrow + (long long)b.off[t] * md + d0;
              const T v = __ldg(a.value + o);
              g_attn += b.wt[t] * v * go;
              g_y += b.dwy[t] * v * gv;
              g_x += b.dwx[t] * v * gv;
              atomicAdd(a.grad_value + o, b.wt[t] * gv);
            }
          }
        }
        g_attn = group_sum(g_attn, G);
        g_y = group_sum(g_y, G);
        g_x = group_sum(g_x, G);
        if (live && lane_in_group == 0) {
          a.grad_attn[hd * a.l * a.p + l * a.p + p] = g_attn;
          a.grad_loc[(hd * a.l * a.p + l * a.p + p) * 2] = (T)a.w[l] * g_x;
          a.grad_loc[(hd * a.l * a.p + l * a.p + p) * 2 + 1] = (T)a.h[l] * g_y;
        }
      }
    }
  }
}

template <typename T>
int fill(const aldi_msda_params* p, MsdaArgs<T>* a, const char* who) {
  ALDI_CHECK_ARG(p && p->value && p->sampling_loc && p->attn_weight && p->spatial_h && p->spatial_w && p->level_start,
                 "%s: null pointer", who);
  ALDI_CHECK_ARG(p->n > 0 && p->s > 0 && p->m > 0 && p->d > 0 && p->lq > 0 && p->p > 0, "%s: empty dimension", who);
  ALDI_CHECK_ARG(p->l >= 1 && p->l <= kMaxLevels, "%s: 1..%d levels supported, got %d", who, kMaxLevels, p->l);
  long long cover = 0;
  for (int i = 0; i < p->l; ++i) {
    ALDI_CHECK_ARG(p->spatial_h[i] > 0 && p->spatial_w[i] > 0 && p->level_start[i] >= 0 &&
                       (long long)p->level_start[i] + (long long)p->spatial_h[i] * p->spatial_w[i] <= p->s,
                   "%s: level %d (%d x %d at %d) does not fit in S=%d", who, i, p->spatial_h[i], p->spatial_w[i],
                   p->level_start[i], p->s);
    a->h[i] = p->spatial_h[i]; a->w[i] = p->spatial_w[i]; a->start[i] = p->level_start[i];
    cover += (long long)p->spatial_h[i] * p->spatial_w[i];
  }
  (void)cover;
  a->value = (const T*)p->value; a->loc = (const T*)p->sampling_loc; a->attn = (const T*)p->attn_weight;
  a->n = p->n; a->s = p->s; a->m = p->m; a->d = p->d; a->lq = p->lq; a->l = p->l; a->p = p->p;
  a->out = (T*)p->out; a->grad_out = (const T*)p->grad_out;
  a->grad_value = (T*)p->grad_value; a->grad_loc = (T*)p->grad_loc; a->grad_attn = (T*)p->grad_attn;
  int g = 1;
  while (g < 32 && g < p->d) g <<= 1;
  a->group = g;
  a->heads = (long long)p->n * p->lq * p->m;
  return ALDI_OK;
}

int grid_for_heads(long long heads, int group) {
  long long blocks = (heads * group + 255) / 256;
  const long long cap = (long long)aldi_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <typename T>
int run_forward(const aldi_msda_params* p, cudaStream_t stream) {
  MsdaArgs<T> a;
  int rc = fill<T>(p, &a, "aldi_msda_forward");
  if (rc) return rc;
  ALDI_CHECK_ARG(p->out, "aldi_msda_forward: null out");
  msda_forward_kernel<T><<<grid_for_heads(a.heads, a.group), 256, 0, stream>>>(a);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_msda_forward");
  return ALDI_OK;
}

template <typename T>
int run_backward(const aldi_msda_params* p, cudaStream_t stream) {
  MsdaArgs<T> a;
  int rc = fill<T>(p, &a, "aldi_msda_backward");
  if (rc) return rc;
  ALDI_CHECK_ARG(p->grad_out && p->grad_value && p->grad_loc && p->grad_attn, "aldi_msda_backward: null gradient pointer");
  msda_backward_kernel<T><<<grid_for_heads(a.heads, a.group), 256, 0, stream>>>(a);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_msda_backward");
  return ALDI_OK;
}

}  // namespace

extern "C" int aldi_msda_forward(const aldi_msda_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(p, "aldi_msda_forward: null params");
  if (p->dtype == ALDI_DTYPE_F32) return run_forward<float>(p, stream);
  if (p->dtype == ALDI_DTYPE_F64) return run_forward<double>(p, stream);
  aldi_set_error("aldi_msda_forward: dtype %d unsupported (the reference op dispatches float / double only)", p->dtype);
  return ALDI_ERR_UNSUPPORTED;
}

extern "C" int aldi_msda_backward(const aldi_msda_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(p, "aldi_msda_backward: null params");
  if (p->dtype == ALDI_DTYPE_F32) return run_backward<float>(p, stream);
  if (p->dtype == ALDI_DTYPE_F64) return run_backward<double>(p, stream);
  aldi_set_error("aldi_msda_backward: dtype %d unsupported (the reference op dispatches float / double only)", p->dtype);
  return ALDI_ERR_UNSUPPORTED;
}
