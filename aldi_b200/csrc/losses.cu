// Fused loss kernels: each produces its scalar losses AND the gradient w.r.t. the head outputs in one pass
// (warp-shuffle + block reduction, one atomic per block per scalar).
//
//   aldi_rpn_loss          detectron2 RPN.losses (rpn.py): BCE-with-logits (sum) on sampled anchors and L1
//                          (smooth-L1 beta=0) on positive anchors, both / (batch_size_per_image * N)
//   aldi_roi_loss          FastRCNNOutputLayers.losses (fast_rcnn.py): CE (mean) + class-specific L1 / #rows
//   aldi_distill_rpn_loss  aldi/distill.py:193-229 get_rpn_losses, INCLUDING its index quirk: masks
//                          are in (n, level, h, w, a) order but are applied to head outputs flattened in
//                          (level, n, a, h, w) / (level, n, a*4+k, h, w) order
//   aldi_distill_roi_loss  aldi/distill.py:231-278 get_roih_losses: soft-target CE (or KL batchmean) and
//                          argmax-gated L1 / #rows
//
// Head outputs are fp32: rpn_out (N, total_locs, ch_stride) with A objectness logits then 4A deltas per
// location; pred (M, pred_stride) with K+1 class logits then 4K box deltas per RoI.  Gradients are written
// densely into channel-padded buffers of the activation dtype (the operand of the next dgrad / wgrad GEMM).
#include "common.cuh"
#include "../../include/aldi_b200.h"

namespace {

__device__ __forceinline__ float bce_with_logits(float x, float y) {
  return fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

template <typename T>
__device__ __forceinline__ void put_grad(T* p, float v, int accumulate) {
  if (accumulate) v += to_f32<T>(*p);
  *p = from_f32<T>(v);
}

__device__ __forceinline__ void level_of_loc(const aldi_rpn_levels& L, int loc, int& lvl, int& hw) {
  lvl = 0;
  while (lvl < L.num_levels - 1 && loc >= L.loc_off[lvl + 1]) ++lvl;
  hw = loc - L.loc_off[lvl];
}
__device__ __forceinline__ void anchor_at(const aldi_rpn_levels& L, int lvl, int hw, int a, float* out) {
  const int w = hw % L.w[lvl], h = hw / L.w[lvl];
  const float sx = (float)(w * L.stride[lvl]), sy = (float)(h * L.stride[lvl]);
  const float* c = L.cell[lvl][a];
  out[0] = sx + c[0]; out[1] = sy + c[1]; out[2] = sx + c[2]; out[3] = sy + c[3];
}
// Box2BoxTransform.get_deltas
__device__ __forceinline__ void get_deltas(const float* src, const float* tgt, float wx, float wy, float ww, float wh,
                                           float* d) {
  const float sw = src[2] - src[0], sh = src[3] - src[1];
  const float scx = src[0] + 0.5f * sw, scy = src[1] + 0.5f * sh;
  const float tw = tgt[2] - tgt[0], th = tgt[3] - tgt[1];
  const float tcx = tgt[0] + 0.5f * tw, tcy = tgt[1] + 0.5f * th;
  d[0] = wx * (tcx - scx) / sw;
  d[1] = wy * (tcy - scy) / sh;
  d[2] = ww * logf(tw / sw);
  d[3] = wh * logf(th / sh);
}

// one thread per (image, location): A anchors x (1 logit + 4 deltas)
template <typename T>
__global__ void __launch_bounds__(256)
rpn_loss_kernel(const float* __restrict__ rpn_out, aldi_rpn_levels L, int n_images, const signed char* __restrict__ labels,
                const int* __restrict__ matched, const float* __restrict__ gt_boxes, const int* __restrict__ gt_counts,
                int gmax, float normalizer, float w_cls, float w_loc, float gscale, T* __restrict__ drpn, int dstride,
                int accumulate, float* __restrict__ loss_out /*[2]*/) {
  __shared__ float red[32];
  const int A = L.num_anchors;
  const long long total = (long long)n_images * L.total_locs;
  float l_cls = 0.f, l_loc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(i / L.total_locs), loc = (int)(i - (long long)img * L.total_locs);
    int lvl, hw;
    level_of_loc(L, loc, lvl, hw);
    const float* row = rpn_out + (size_t)i * L.ch_stride;
    T* drow = drpn + (size_t)i * dstride;
    const int g = gt_counts[img];
    for (int a = 0; a < A; ++a) {
      const size_t r = (size_t)img * L.total_locs * A + (size_t)loc * A + a;
      const int lab = labels[r];
      float dlogit = 0.f, dd[4] = {0.f, 0.f, 0.f, 0.f};
      if (lab >= 0) {
        const float x = row[a];
        l_cls += bce_with_logits(x, (float)lab);
        dlogit = (sigmoidf_(x) - (float)lab) * (w_cls * gscale / normalizer);
      }
      if (lab == 1) {
        float anc[4], tgt[4] = {0.f, 0.f, 0.f, 0.f}, t[4];
        anchor_at(L, lvl, hw, a, anc);
        if (g > 0) {
          const float* q = gt_boxes + ((size_t)img * gmax + matched[r]) * 4;
          tgt[0] = q[0]; tgt[1] = q[1]; tgt[2] = q[2]; tgt[3] = q[3];
        }
        get_deltas(anc, tgt, 1.f, 1.f, 1.f, 1.f, t);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float diff = row[A + a * 4 + k] - t[k];
          l_loc += fabsf(diff);
          dd[k] = sgn(diff) * (w_loc * gscale / normalizer);
        }
      }
      put_grad(drow + a, dlogit, accumulate);
#pragma unroll
      for (int k = 0; k < 4; ++k) put_grad(drow + A + a * 4 + k, dd[k], accumulate);
    }
  }
  l_cls = block_sum(l_cls, red);
  l_loc = block_sum(l_loc, red);
  if (threadIdx.x == 0) {
    atomicAdd(loss_out + 0, l_cls * (w_cls * gscale / normalizer));
    atomicAdd(loss_out + 1, l_loc * (w_loc * gscale / normalizer));
  }
}

// one thread per RoI row
template <typename T>
__global__ void __launch_bounds__(256)
roi_loss_kernel(const float* __restrict__ pred, int pred_stride, int m, int num_classes, const int* __restrict__ gt_class,
                const float* __restrict__ roi_box, const float* __restrict__ gt_box, const int* __restrict__ counts,
                int n_images, float wx, float wy, float ww, float wh, float w_cls, float w_box, float gscale,
                T* __restrict__ dpred, int dstride, float* __restrict__ loss_out /*[2]*/) {
  __shared__ float red[32];
  const int K = num_classes;
  int mvalid = 0;
  for (int i = 0; i < n_images; ++i) mvalid += counts[i];
  const float inv = 1.f / fmaxf((float)mvalid, 1.f);
  float l_cls = 0.f, l_box = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
    const float* row = pred + (size_t)r * pred_stride;
    T* drow = dpred + (size_t)r * dstride;
    const int cls = gt_class[r];
    if (cls < 0) {  // padding row
      for (int k = 0; k < 5 * K + 1; ++k) drow[k] = from_f32<T>(0.f);
      continue;
    }
    float mx = row[0];
    for (int k = 1; k <= K; ++k) mx = fmaxf(mx, row[k]);
    float sum = 0.f;
    for (int k = 0; k <= K; ++k) sum += expf(row[k] - mx);
    const float lse = mx + logf(sum);
    l_cls += lse - row[cls];
    for (int k = 0; k <= K; ++k)
      drow[k] = from_f32<T>((expf(row[k] - lse) - (k == cls ? 1.f : 0.f)) * (w_cls * gscale * inv));
    for (int k = 0; k < 4 * K; ++k) drow[K + 1 + k] = from_f32<T>(0.f);
    if (cls < K) {
      float t[4];
      get_deltas(roi_box + (size_t)r * 4, gt_box + (size_t)r * 4, wx, wy, ww, wh, t);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float diff = row[K + 1 + cls * 4 + k] - t[k];
        l_box += fabsf(diff);
        drow[K + 1 + cls * 4 + k] = from_f32<T>(sgn(diff) * (w_box * gscale * inv));
      }
    }
  }
  l_cls = block_sum(l_cls, red);
  l_box = block_sum(l_box, red);
  if (threadIdx.x == 0) {
    atomicAdd(loss_out + 0, l_cls * (w_cls * gscale * inv));
    atomicAdd(loss_out + 1, l_box * (w_box * gscale * inv));
  }
}

// one thread per (image, location); see file header for the (deliberately reproduced) index mapping
template <typename T>
__global__ void __launch_bounds__(256)
distill_rpn_kernel(const float* __restrict__ s_out, const float* __restrict__ t_out, aldi_rpn_levels L, int n_images,
                   const signed char* __restrict__ labels_flat, const int* __restrict__ stats /*N*2 num_pos,num_neg*/,
                   float obj_temperature, float w_obj, float w_reg, float gscale, T* __restrict__ drpn, int dstride,
                   int accumulate, float* __restrict__ loss_out /*[2]*/) {
  __shared__ float red[32];
  const int A = L.num_anchors;
  int n_valid = 0, n_fg = 0;
  for (int i = 0; i < n_images; ++i) { n_fg += stats[2 * i]; n_valid += stats[2 * i] + stats[2 * i + 1]; }
  const float inv_valid = n_valid > 0 ? 1.f / (float)n_valid : 0.f;
  const float inv_fg4 = n_fg > 0 ? 1.f / (4.f * (float)n_fg) : 0.f;
  const long long total = (long long)n_images * L.total_locs;
  float l_obj = 0.f, l_reg = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(i / L.total_locs), loc = (int)(i - (long long)img * L.total_locs);
    int lvl, hw;
    level_of_loc(L, loc, lvl, hw);
    const long long HW = (long long)L.h[lvl] * L.w[lvl];
    const long long off_l = (long long)L.loc_off[lvl] * A;  // anchors per image before this level
    const float* srow = s_out + (size_t)i * L.ch_stride;
    const float* trow = t_out + (size_t)i * L.ch_stride;
    T* drow = drpn + (size_t)i * dstride;
    for (int a = 0; a < A; ++a) {
      // objectness element (level, n, a, hw) of the flattened head output
      const long long j = (long long)n_images * off_l + (long long)img * A * HW + (long long)a * HW + hw;
      float g = 0.f;
      if (labels_flat[j] >= 0) {
        const float p_t = sigmoidf_(trow[a] / obj_temperature);
        const float x = srow[a];
        l_obj += bce_with_logits(x, p_t);
        g = (sigmoidf_(x) - p_t) * (w_obj * gscale * inv_valid);
      }
      put_grad(drow + a, g, accumulate);
    }
    for (int c = 0; c < 4 * A; ++c) {
      // delta element (level, n, c, hw); the mask entry consulted is floor(m / 4)
      const long long m = 4ll * n_images * off_l + (long long)img * 4 * A * HW + (long long)c * HW + hw;
      float g = 0.f;
      if (labels_flat[m >> 2] == 1) {
        const float diff = srow[A + c] - trow[A + c];
        l_reg += fabsf(diff);
        g = sgn(diff) * (w_reg * gscale * inv_fg4);
      }
      put_grad(drow + A + c, g, accumulate);
    }
  }
  l_obj = block_sum(l_obj, red);
  l_reg = block_sum(l_reg, red);
  if (threadIdx.x == 0) {
    atomicAdd(loss_out + 0, l_obj * (w_obj * gscale * inv_valid));
    atomicAdd(loss_out + 1, l_reg * (w_reg * gscale * inv_fg4));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
distill_roi_kernel(const float* __restrict__ s_pred, const float* __restrict__ t_pred, int pred_stride, int m,
                   int num_classes, const int* __restrict__ row_class /*-1 = padding*/, const int* __restrict__ counts,
                   int n_images, float cls_temperature, int kl, float w_cls, float w_reg, float gscale,
                   T* __restrict__ dpred, int dstride, int accumulate, float* __restrict__ loss_out /*[2]*/) {
  __shared__ float red[32];
  const int K = num_classes;
  int mvalid = 0;
  for (int i = 0; i < n_images; ++i) mvalid += counts[i];
  const float inv = mvalid > 0 ? 1.f / (float)mvalid : 0.f;
  float l_cls = 0.f, l_reg = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
    T* drow = dpred + (size_t)r * dstride;
    if (row_class[r] < 0) {
      if (!accumulate)
        for (int k = 0; k < 5 * K + 1; ++k) drow[k] = from_f32<T>(0.f);
      continue;
    }
    const float* s = s_pred + (size_t)r * pred_stride;
    const float* t = t_pred + (size_t)r * pred_stride;
    // teacher: softmax(t / T), argmax of the raw logits (first maximum)
    float tmx = t[0] / cls_temperature;
    int arg = 0;
    float rawmx = t[0];
    for (int k = 1; k <= K; ++k) {
      tmx = fmaxf(tmx, t[k] / cls_temperature);
      if (t[k] > rawmx) { rawmx = t[k]; arg = k; }
    }
    float tsum = 0.f;
    for (int k = 0; k <= K; ++k) tsum += expf(t[k] / cls_temperature - tmx);
    const float tlse = tmx + logf(tsum);
    float smx = s[0];
    for (int k = 1; k <= K; ++k) smx = fmaxf(smx, s[k]);
    float ssum = 0.f;
    for (int k = 0; k <= K; ++k) ssum += expf(s[k] - smx);
    const float slse = smx + logf(ssum);
    for (int k = 0; k <= K; ++k) {
      const float logp_t = t[k] / cls_temperature - tlse;
      const float p_t = expf(logp_t);
      const float logp_s = s[k] - slse;
      l_cls += kl ? p_t * (logp_t - logp_s) : -p_t * logp_s;
      put_grad(drow + k, (expf(logp_s) - p_t) * (w_cls * gscale * inv), accumulate);
    }
    for (int k = 0; k < 4 * K; ++k) {
      float g = 0.f;
      if (arg != K && (k >> 2) == arg) {
        const float diff = s[K + 1 + k] - t[K + 1 + k];
        l_reg += fabsf(diff);
        g = sgn(diff) * (w_reg * gscale * inv);
      }
      put_grad(drow + K + 1 + k, g, accumulate);
    }
  }
  l_cls = block_sum(l_cls, red);
  l_reg = block_sum(l_reg, red);
  if (threadIdx.x == 0) {
    atomicAdd(loss_out + 0, l_cls * (w_cls * gscale * inv));
    atomicAdd(loss_out + 1, l_reg * (w_reg * gscale * inv));
  }
}

// aldi/align.py:81-90: loss = weight * mean(BCE-with-logits(pred, domain_label)); grad wrt pred
template <typename T>
__global__ void __launch_bounds__(256)
domain_bce_kernel(const float* __restrict__ pred, int n, int stride, float label, float weight, float gscale,
                  T* __restrict__ dpred, int dstride, float* __restrict__ loss_out) {
  __shared__ float red[32];
  float l = 0.f;
  const float inv = 1.f / (float)n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = pred[(size_t)i * stride];
    l += bce_with_logits(x, label);
    dpred[(size_t)i * dstride] = from_f32<T>((sigmoidf_(x) - label) * (weight * gscale * inv));
  }
  l = block_sum(l, red);
  if (threadIdx.x == 0) atomicAdd(loss_out, l * (weight * gscale * inv));
}

// aldi/align.py:81-90 + the discriminator's final Linear(C, 1): logits = feat . w + b for every row, BCE-with-logits
// (mean over the valid rows) against the constant domain label, and in the same pass the gradients of that layer:
// dw += sum_r dl[r] * feat[r], db += sum_r dl[r], dfeat[r] = dl[r] * w (optionally gated by feat > 0, i.e. already
// the gradient w.r.t. the PRE-activation of the ReLU that produced feat) and its negation (the operand of the
// gradient-reversed data gradient, aldi/helpers.py:51-63).
template <typename T>
__global__ void __launch_bounds__(256)
domain_head_kernel(const T* __restrict__ feat, int n, int c, long long feat_stride, const int* __restrict__ counts,
                   int rows_per_image, const float* __restrict__ w, const float* __restrict__ b, float label,
                   float weight, float gscale, int relu_mask, T* __restrict__ dfeat, T* __restrict__ ndfeat,
                   long long d_stride, float* __restrict__ dw, float* __restrict__ db, float* __restrict__ loss_out) {
  constexpr int R = 32;
  __shared__ float s_dl[R];
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float denom = (float)n;
  if (counts) {
    int tot = 0;
    const int n_img = (n + rows_per_image - 1) / rows_per_image;
    for (int i = 0; i < n_img; ++i) tot += min(counts[i], rows_per_image);
    denom = (float)max(tot, 1);
  }
  const float k = weight * gscale / denom;
  const float bias = b[0];
  float loss = 0.f, dsum = 0.f;
  for (int r0 = blockIdx.x * R; r0 < n; r0 += gridDim.x * R) {
    __syncthreads();
    for (int rr = warp; rr < R; rr += 8) {
      const int r = r0 + rr;
      float dl = 0.f;
      if (r < n) {
        const bool valid = !counts || (r % rows_per_image) < counts[r / rows_per_image];
        float dot = 0.f;
        if (valid) {
          const T* f = feat + (long long)r * feat_stride;
          for (int ch = lane; ch < c; ch += 32) dot += to_f32<T>(f[ch]) * w[ch];
        }
        dot = warp_sum(dot);
        if (valid) {
          const float x = dot + bias;
          dl = (sigmoidf_(x) - label) * k;
          if (lane == 0) { loss += bce_with_logits(x, label); dsum += dl; }
        }
      }
      if (lane == 0) s_dl[rr] = dl;
    }
    __syncthreads();
    const int rmax = min(R, n - r0);
    for (int ch = threadIdx.x; ch < c; ch += 256) {
      const float wv = w[ch];
      float acc = 0.f;
      for (int rr = 0; rr < rmax; ++rr) {
        const long long r = r0 + rr;
        const float f = to_f32<T>(feat[r * feat_stride + ch]);
        const float dl = s_dl[rr];
        acc += dl * f;
        float g = dl * wv;
        if (relu_mask && !(f > 0.f)) g = 0.f;
        dfeat[r * d_stride + ch] = from_f32<T>(g);
        if (ndfeat) ndfeat[r * d_stride + ch] = from_f32<T>(-g);
      }
      atomicAdd(dw + ch, acc);
    }
  }
  loss = block_sum(loss, red);
  dsum = block_sum(dsum, red);
  if (threadIdx.x == 0) {
    if (loss != 0.f) atomicAdd(loss_out, loss * k);
    if (dsum != 0.f) atomicAdd(db, dsum);
  }
}

int blocks_for(long long work) {
  long long b = (work + 255) / 256;
  long long cap = (long long)aldi_num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int aldi_rpn_loss(const float* rpn_out, const aldi_rpn_levels* L, int n_images, const signed char* labels,
                             const int* matched, const float* gt_boxes, const int* gt_counts, int gmax,
                             int batch_size_per_image, float w_cls, float w_loc, float gscale, void* drpn, int dtype,
                             int dstride, int accumulate, float* loss_out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(rpn_out && L && labels && matched && gt_boxes && gt_counts && drpn && loss_out,
                 "aldi_rpn_loss: null pointer");
  ALDI_CHECK_ARG(dstride >= 5 * L->num_anchors, "aldi_rpn_loss: dstride too small");
  const float normalizer = (float)batch_size_per_image * (float)n_images;
  const int grid = blocks_for((long long)n_images * L->total_locs);
  if (dtype == ALDI_DTYPE_BF16)
    rpn_loss_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(rpn_out, *L, n_images, labels, matched, gt_boxes, gt_counts,
                                                             gmax, normalizer, w_cls, w_loc, gscale,
                                                             (__nv_bfloat16*)drpn, dstride, accumulate, loss_out);
  else
    rpn_loss_kernel<float><<<grid, 256, 0, stream>>>(rpn_out, *L, n_images, labels, matched, gt_boxes, gt_counts, gmax,
                                                     normalizer, w_cls, w_loc, gscale, (float*)drpn, dstride,
                                                     accumulate, loss_out);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_rpn_loss");
  return ALDI_OK;
}

extern "C" int aldi_roi_loss(const float* pred, int pred_stride, int m, int num_classes, const int* gt_class,
                             const float* roi_box, const float* gt_box, const int* counts, int n_images,
                             const float* h_weights4, float w_cls, float w_box, float gscale, void* dpred, int dtype,
                             int dstride, float* loss_out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(pred && gt_class && roi_box && gt_box && counts && h_weights4 && dpred && loss_out,
                 "aldi_roi_loss: null pointer");
  ALDI_CHECK_ARG(pred_stride >= 5 * num_classes + 1 && dstride >= 5 * num_classes + 1, "aldi_roi_loss: strides too small");
  const int grid = blocks_for(m);
  if (dtype == ALDI_DTYPE_BF16)
    roi_loss_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(pred, pred_stride, m, num_classes, gt_class, roi_box, gt_box,
                                                             counts, n_images, h_weights4[0], h_weights4[1],
                                                             h_weights4[2], h_weights4[3], w_cls, w_box, gscale,
                                                             (__nv_bfloat16*)dpred, dstride, loss_out);
  else
    roi_loss_kernel<float><<<grid, 256, 0, stream>>>(pred, pred_stride, m, num_classes, gt_class, roi_box, gt_box, counts,
                                                     n_images, h_weights4[0], h_weights4[1], h_weights4[2],
                                                     h_weights4[3], w_cls, w_box, gscale, (float*)dpred, dstride,
                                                     loss_out);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_roi_loss");
  return ALDI_OK;
}

extern "C" int aldi_distill_rpn_loss(const float* student_rpn_out, const float* teacher_rpn_out,
                                     const aldi_rpn_levels* L, int n_images, const signed char* labels,
                                     const int* stats, float obj_temperature, float w_obj, float w_reg, float gscale,
                                     void* drpn, int dtype, int dstride, int accumulate, float* loss_out,
                                     void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(student_rpn_out && teacher_rpn_out && L && labels && stats && drpn && loss_out,
                 "aldi_distill_rpn_loss: null pointer");
  ALDI_CHECK_ARG(dstride >= 5 * L->num_anchors, "aldi_distill_rpn_loss: dstride too small");
  const int grid = blocks_for((long long)n_images * L->total_locs);
  if (dtype == ALDI_DTYPE_BF16)
    distill_rpn_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(student_rpn_out, teacher_rpn_out, *L, n_images, labels,
                                                                stats, obj_temperature, w_obj, w_reg, gscale,
                                                                (__nv_bfloat16*)drpn, dstride, accumulate, loss_out);
  else
    distill_rpn_kernel<float><<<grid, 256, 0, stream>>>(student_rpn_out, teacher_rpn_out, *L, n_images, labels, stats,
                                                        obj_temperature, w_obj, w_reg, gscale, (float*)drpn, dstride,
                                                        accumulate, loss_out);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_distill_rpn_loss");
  return ALDI_OK;
}

extern "C" int aldi_distill_roi_loss(const float* student_pred, const float* teacher_pred, int pred_stride, int m,
                                     int num_classes, const int* row_class, const int* counts, int n_images,
                                     float cls_temperature, int use_kl, float w_cls, float w_reg, float gscale,
                                     void* dpred, int dtype, int dstride, int accumulate, float* loss_out,
                                     void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(student_pred && teacher_pred && row_class && counts && dpred && loss_out,
                 "aldi_distill_roi_loss: null pointer");
  const int grid = blocks_for(m);
  if (dtype == ALDI_DTYPE_BF16)
    distill_roi_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(student_pred, teacher_pred, pred_stride, m, num_classes,
                                                                row_class, counts, n_images, cls_temperature, use_kl,
                                                                w_cls, w_reg, gscale, (__nv_bfloat16*)dpred, dstride,
                                                                accumulate, loss_out);
  else
    distill_roi_kernel<float><<<grid, 256, 0, stream>>>(student_pred, teacher_pred, pred_stride, m, num_classes, row_class,
                                                        counts, n_images, cls_temperature, use_kl, w_cls, w_reg, gscale,
                                                        (float*)dpred, dstride, accumulate, loss_out);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_distill_roi_loss");
  return ALDI_OK;
}

extern "C" int aldi_domain_bce_loss(const float* pred, int n, int stride, float domain_label, float weight,
                                    float gscale, void* dpred, int dtype, int dstride, float* loss_out,
                                    void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(pred && dpred && loss_out && n > 0, "aldi_domain_bce_loss: bad args");
  const int grid = blocks_for(n);
  if (dtype == ALDI_DTYPE_BF16)
    domain_bce_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(pred, n, stride, domain_label, weight, gscale,
                                                               (__nv_bfloat16*)dpred, dstride, loss_out);
  else
    domain_bce_kernel<float><<<grid, 256, 0, stream>>>(pred, n, stride, domain_label, weight, gscale, (float*)dpred,
                                                       dstride, loss_out);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_domain_bce_loss");
  return ALDI_OK;
}

extern "C" int aldi_domain_head_loss(const void* feat, int dtype, int n, int c, long long feat_stride, const int* counts,
                                     int rows_per_image, const float* w, const float* b, float domain_label,
                                     float weight, float gscale, int relu_mask, void* dfeat, void* ndfeat,
                                     long long d_stride, float* dw, float* db, float* loss_out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(feat && w && b && dfeat && dw && db && loss_out && n > 0 && c > 0, "aldi_domain_head_loss: bad args");
  ALDI_CHECK_ARG(!counts || rows_per_image > 0, "aldi_domain_head_loss: counts need rows_per_image");
  int grid = (n + 31) / 32;
  const int cap = aldi_num_sms() * 4;
  if (grid > cap) grid = cap;
  if (dtype == ALDI_DTYPE_BF16)
    domain_head_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        (const __nv_bfloat16*)feat, n, c, feat_stride, counts, rows_per_image, w, b, domain_label, weight, gscale,
        relu_mask, (__nv_bfloat16*)dfeat, (__nv_bfloat16*)ndfeat, d_stride, dw, db, loss_out);
  else
    domain_head_kernel<float><<<grid, 256, 0, stream>>>((const float*)feat, n, c, feat_stride, counts, rows_per_image, w,
                                                        b, domain_label, weight, gscale, relu_mask, (float*)dfeat,
                                                        (float*)ndfeat, d_stride, dw, db, loss_out);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_domain_head_loss");
  return ALDI_OK;
}
