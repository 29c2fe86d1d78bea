// Selection kernels of the detector: RPN per-level top-k + box decode, score-sorted categorical NMS,
// anchor / proposal matching and seeded subsampling.  Everything here decides INDEX SETS, which must be
// bit-exact against the reference (BASELINE north_star), so this file is compiled with -fmad=false and
// mirrors the reference's floating-point expression order (detectron2 box_regression.py / boxes.py /
// matcher.py / sampling.py, torchvision nms).  No host synchronisation anywhere: variable-size results
// are written into fixed-capacity buffers with device-side counts.
//
// Replaces: detectron2 find_top_rpn_proposals / batched_nms / Matcher / subsample_labels /
// fast_rcnn_inference (reached from aldi/pseudolabeler.py:21, aldi/distill.py:157,162,200-202,
// aldi/trainer.py:87).
#include "common.cuh"
#include "../../include/aldi_b200.h"
#include <float.h>

namespace {

// ------------------------------------------------------------------------------------------------
// order-preserving float <-> uint32
__device__ __forceinline__ uint32_t fkey(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// murmur3 finaliser based counter hash: the sampling "random permutation" is the order of these keys.
// Must stay in sync with aldi_b200/sampling.py (host/oracle emulation).
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ uint32_t sample_hash(uint32_t seed, uint32_t salt, uint32_t index) {
  uint32_t s = fmix32(seed ^ (salt * 0x27D4EB2Fu + 0x165667B1u));
  return fmix32((index * 0x9E3779B1u) ^ s);
}

// ------------------------------------------------------------------------------------------------
// Block-wide selection of the k LARGEST 32-bit keys among n candidates (MSB-first radix select).
// kf(i, key) -> bool valid.  Result: every element with key > T is selected, plus `take_eq` of the
// elements with key == T (lowest index first); count_eq = number of elements with key == T.
struct SelectResult {
  uint32_t T;
  int take_eq, count_eq;
};

template <typename KeyFn>
__device__ SelectResult block_select(int n, int k, KeyFn kf, uint32_t* s_hist /*>=260 words*/) {
  SelectResult res;
  uint32_t prefix = 0, mask = 0;
  int remaining = k;
  for (int pass = 3; pass >= 0; --pass) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int shift = 8 * pass;
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
      const int i = i0 + threadIdx.x;
      uint32_t key = 0;
      bool ok = (i < n) && kf(i, key) && ((key & mask) == prefix);
      const uint32_t bin = (key >> shift) & 255u;
      // warp-aggregated histogram update (values cluster in few bins on the high digits)
      const uint32_t active = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const uint32_t peers = __match_any_sync(active, bin);
        if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&s_hist[bin], (uint32_t)__popc(peers));
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int cum = 0, b = 255;
      for (; b > 0; --b) {
        if (cum + (int)s_hist[b] >= remaining) break;
        cum += (int)s_hist[b];
      }
      s_hist[256] = (uint32_t)b;
      s_hist[257] = (uint32_t)(remaining - cum);
      s_hist[258] = s_hist[b];
    }
    __syncthreads();
    prefix |= s_hist[256] << shift;
    mask |= 255u << shift;
    remaining = (int)s_hist[257];
    res.count_eq = (int)s_hist[258];
    __syncthreads();
  }
  res.T = prefix;
  res.take_eq = remaining;
  return res;
}

// Ordered (index-ascending) compaction of the selected set into out[0..k): used by every selector.
// Elements > T are written in index order interleaved with the first take_eq elements == T.
template <typename KeyFn, typename Emit>
__device__ void block_emit_selected(int n, int k, const SelectResult& r, KeyFn kf, Emit emit, int* s_scan /*>=40*/) {
  // running counters: s_scan[32] = written so far, s_scan[33] = eq taken so far
  if (threadIdx.x == 0) { s_scan[32] = 0; s_scan[33] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    uint32_t key = 0;
    const bool ok = (i < n) && kf(i, key);
    const bool gt = ok && key > r.T;
    const bool eq = ok && key == r.T;
    // ordered ranks of eq elements (needed only to cap them at take_eq)
    const uint32_t beq = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) s_scan[warp] = __popc(beq);
    __syncthreads();
    int eq_before = s_scan[33];
    for (int w = 0; w < warp; ++w) eq_before += s_scan[w];
    const int eq_rank = eq_before + __popc(beq & ((1u << lane) - 1u));
    const bool sel = gt || (eq && eq_rank < r.take_eq);
    int eq_total = 0;
    for (int w = 0; w < nwarps; ++w) eq_total += s_scan[w];
    __syncthreads();
    const uint32_t bsel = __ballot_sync(0xffffffffu, sel);
    if (lane == 0) s_scan[warp] = __popc(bsel);
    __syncthreads();
    int before = s_scan[32];
    for (int w = 0; w < warp; ++w) before += s_scan[w];
    const int pos = before + __popc(bsel & ((1u << lane) - 1u));
    if (sel && pos < k) emit(pos, i, key);
    int sel_total = 0;
    for (int w = 0; w < nwarps; ++w) sel_total += s_scan[w];
    __syncthreads();
    if (threadIdx.x == 0) { s_scan[32] += sel_total; s_scan[33] += eq_total; }
    __syncthreads();
    if (s_scan[32] >= k) break;
  }
  __syncthreads();
}

// Emission in arbitrary order (the caller sorts afterwards); falls back to the ordered pass only when
// equal keys straddle the cut (lowest index wins).
template <typename KeyFn, typename Emit>
__device__ void block_emit_any_order(int n, int k, const SelectResult& r, KeyFn kf, Emit emit, int* s_scan) {
  if (r.count_eq > r.take_eq) {
    block_emit_selected(n, k, r, kf, emit, s_scan);
    return;
  }
  if (threadIdx.x == 0) s_scan[32] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    uint32_t key = 0;
    if (kf(i, key) && key >= r.T) {
      const int pos = atomicAdd(&s_scan[32], 1);
      if (pos < k) emit(pos, i, key);
    }
  }
  __syncthreads();
}

// in-place bitonic sort (descending) of n_pow2 64-bit keys in shared memory
__device__ void block_bitonic_desc(unsigned long long* s, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = s[lo], b = s[hi];
        if ((a < b) == desc) { s[lo] = b; s[hi] = a; }
      }
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Box2BoxTransform.apply_deltas + Boxes.clip (detectron2 box_regression.py:88-116, boxes.py clip)
__device__ __forceinline__ void apply_deltas(const float* box, float d0, float d1, float d2, float d3, float wx,
                                             float wy, float ww, float wh, float clampv, float* out) {
  const float widths = box[2] - box[0], heights = box[3] - box[1];
  const float ctr_x = box[0] + 0.5f * widths, ctr_y = box[1] + 0.5f * heights;
  const float dx = d0 / wx, dy = d1 / wy;
  float dw = d2 / ww, dh = d3 / wh;
  dw = fminf(dw, clampv);
  dh = fminf(dh, clampv);
  const float pcx = dx * widths + ctr_x, pcy = dy * heights + ctr_y;
  const float pw = expf(dw) * widths, ph = expf(dh) * heights;
  out[0] = pcx - 0.5f * pw;
  out[1] = pcy - 0.5f * ph;
  out[2] = pcx + 0.5f * pw;
  out[3] = pcy + 0.5f * ph;
}
__device__ __forceinline__ void clip_box(float* b, float h, float w) {
  b[0] = fminf(fmaxf(b[0], 0.f), w);
  b[1] = fminf(fmaxf(b[1], 0.f), h);
  b[2] = fminf(fmaxf(b[2], 0.f), w);
  b[3] = fminf(fmaxf(b[3], 0.f), h);
}
__device__ __forceinline__ void anchor_box(const aldi_rpn_levels& L, int lvl, int e, float* out) {
  // e = (h*W + w)*A + a ; DefaultAnchorGenerator: shift (w*stride, h*stride) + cell anchor
  const int A = L.num_anchors;
  const int a = e % A;
  const int loc = e / A;
  const int w = loc % L.w[lvl], h = loc / L.w[lvl];
  const float sx = (float)(w * L.stride[lvl]), sy = (float)(h * L.stride[lvl]);
  const float* c = L.cell[lvl][a];
  out[0] = sx + c[0]; out[1] = sy + c[1]; out[2] = sx + c[2]; out[3] = sy + c[3];
}
// detectron2 pairwise_iou (boxes1 = gt `g`, boxes2 = `b`)
__device__ __forceinline__ float d2_iou(const float* g, float garea, const float* b, float barea) {
  float w = fminf(g[2], b[2]) - fmaxf(g[0], b[0]);
  float h = fminf(g[3], b[3]) - fmaxf(g[1], b[1]);
  w = fmaxf(w, 0.f);
  h = fmaxf(h, 0.f);
  const float inter = w * h;
  return inter > 0.f ? inter / (garea + barea - inter) : 0.f;
}

// ------------------------------------------------------------------------------------------------
// K1: per (image, level): top-k objectness logits (sorted descending), decode, clip, validity.
// grid = (levels, N), block = 1024.
__global__ void __launch_bounds__(1024)
rpn_topk_decode_kernel(const float* __restrict__ rpn_out, aldi_rpn_levels L, int pre_topk, const int* __restrict__ img_sizes,
                       float* __restrict__ cand_box, float* __restrict__ cand_score, int* __restrict__ cand_cat,
                       int* __restrict__ cand_idx, unsigned char* __restrict__ cand_valid, int cand_stride,
                       int* __restrict__ err_flag) {
  extern __shared__ unsigned long long s_keys[];  // 2048 entries
  __shared__ uint32_t s_hist[260];
  __shared__ int s_scan[40];
  const int lvl = blockIdx.x, img = blockIdx.y;
  const int A = L.num_anchors;
  const int n = L.h[lvl] * L.w[lvl] * A;
  const int k = n < pre_topk ? n : pre_topk;
  int cand_off = 0;
  for (int l = 0; l < lvl; ++l) {
    const int nl = L.h[l] * L.w[l] * A;
    cand_off += nl < pre_topk ? nl : pre_topk;
  }
  const float* base = rpn_out + ((size_t)img * L.total_locs + L.loc_off[lvl]) * L.ch_stride;
  auto kf = [&](int i, uint32_t& key) -> bool {
    const int loc = i / A, a = i - loc * A;
    key = fkey(__ldg(base + (size_t)loc * L.ch_stride + a));
    return true;
  };
  const SelectResult r = block_select(n, k, kf, s_hist);
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) s_keys[i] = 0ull;
  __syncthreads();
  auto emit = [&](int pos, int i, uint32_t key) {
    s_keys[pos] = ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
  };
  block_emit_any_order(n, k, r, kf, emit, s_scan);
  block_bitonic_desc(s_keys, 2048);
  const float img_h = (float)img_sizes[2 * img], img_w = (float)img_sizes[2 * img + 1];
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const int e = (int)(0xFFFFFFFFu - (uint32_t)(s_keys[j] & 0xFFFFFFFFull));
    const int loc = e / A, a = e - loc * A;
    const float* row = base + (size_t)loc * L.ch_stride;
    const float score = row[a];
    float anc[4], box[4];
    anchor_box(L, lvl, e, anc);
    const float* d = row + A + a * 4;
    apply_deltas(anc, d[0], d[1], d[2], d[3], 1.f, 1.f, 1.f, 1.f, L.scale_clamp, box);
    const bool finite = isfinite(box[0]) && isfinite(box[1]) && isfinite(box[2]) && isfinite(box[3]) && isfinite(score);
    if (!finite && err_flag) atomicOr(err_flag, 1);
    clip_box(box, img_h, img_w);
    const bool nonempty = (box[2] - box[0] > L.min_box_size) && (box[3] - box[1] > L.min_box_size);
    const size_t o = (size_t)img * cand_stride + cand_off + j;
    cand_box[o * 4 + 0] = box[0]; cand_box[o * 4 + 1] = box[1]; cand_box[o * 4 + 2] = box[2]; cand_box[o * 4 + 3] = box[3];
    cand_score[o] = score;
    cand_cat[o] = lvl;
    cand_idx[o] = e;
    cand_valid[o] = (finite && nonempty) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// K3a: per image: sort valid candidates by score (descending, ties by position) and gather.
// block = 1024, dynamic smem = cpad * 8 bytes.
__global__ void __launch_bounds__(1024)
nms_sort_kernel(const float* __restrict__ box, const float* __restrict__ score, const int* __restrict__ cat,
                const unsigned char* __restrict__ valid, const int* __restrict__ counts, int cand_stride, int cpad,
                float* __restrict__ sbox, float* __restrict__ sscore, int* __restrict__ scat, int* __restrict__ sperm,
                int* __restrict__ nvalid) {
  extern __shared__ unsigned long long s_keys[];
  __shared__ int s_cnt;
  const int img = blockIdx.x;
  const int c = counts ? min(counts[img], cand_stride) : cand_stride;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < cpad; i += blockDim.x) {
    unsigned long long key = 0ull;
    if (i < c && (!valid || valid[(size_t)img * cand_stride + i])) {
      key = ((unsigned long long)fkey(score[(size_t)img * cand_stride + i]) << 32) |
            (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
      ++local;
    }
    s_keys[i] = key;
  }
  atomicAdd(&s_cnt, local);
  block_bitonic_desc(s_keys, cpad);
  const int nv = s_cnt;
  if (threadIdx.x == 0) nvalid[img] = nv;
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    const int i = (int)(0xFFFFFFFFu - (uint32_t)(s_keys[j] & 0xFFFFFFFFull));
    const size_t src = (size_t)img * cand_stride + i, dst = (size_t)img * cpad + j;
    reinterpret_cast<float4*>(sbox)[dst] = reinterpret_cast<const float4*>(box)[src];
    sscore[dst] = score[src];
    scat[dst] = cat[src];
    sperm[dst] = i;
  }
}

// K3b: suppression bit matrix among same-category pairs (torchvision nms IoU expression, `>` threshold).
// grid = (col blocks, row blocks, N), block = 64.
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float* __restrict__ sbox, const int* __restrict__ scat, const int* __restrict__ nvalid, int cpad,
                float thresh, unsigned long long* __restrict__ mask) {
  const int img = blockIdx.z, rb = blockIdx.y, cb = blockIdx.x;
  const int nv = nvalid[img];
  if (cb < rb || rb * 64 >= nv || cb * 64 >= nv) return;
  __shared__ float4 cbox[64];
  __shared__ int ccat[64];
  const int words = cpad >> 6;
  const int cj = cb * 64 + threadIdx.x;
  if (cj < nv) {
    cbox[threadIdx.x] = reinterpret_cast<const float4*>(sbox)[(size_t)img * cpad + cj];
    ccat[threadIdx.x] = scat[(size_t)img * cpad + cj];
  }
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i >= nv) return;
  const float4 bi = reinterpret_cast<const float4*>(sbox)[(size_t)img * cpad + i];
  const int ci = scat[(size_t)img * cpad + i];
  const float iarea = (bi.z - bi.x) * (bi.w - bi.y);
  unsigned long long bits = 0ull;
  const int ncol = min(64, nv - cb * 64);
  const int start = (rb == cb) ? threadIdx.x + 1 : 0;
  for (int j = start; j < ncol; ++j) {
    if (ccat[j] != ci) continue;
    const float4 bj = cbox[j];
    const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
    const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
    const float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
    const float inter = w * h;
    const float jarea = (bj.z - bj.x) * (bj.w - bj.y);
    const float ovr = inter / (iarea + jarea - inter);
    if (ovr > thresh) bits |= 1ull << j;
  }
  mask[((size_t)img * cpad + i) * words + cb] = bits;
}

// K3c: greedy scan in score order, 64 boxes at a time.  grid = N, block = 256 (thread w owns word w of
// the `removed` bitmap).  Writes the first `post_topk` kept boxes.
__global__ void __launch_bounds__(256)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const float* __restrict__ sbox,
                const float* __restrict__ sscore, const int* __restrict__ scat, const int* __restrict__ sperm,
                const int* __restrict__ nvalid, int cpad, int post_topk, float* __restrict__ out_box,
                float* __restrict__ out_score, int* __restrict__ out_cat, int* __restrict__ out_src,
                int* __restrict__ out_count) {
  const int img = blockIdx.x;
  const int nv = nvalid[img];
  const int words = cpad >> 6;
  const int w = threadIdx.x;
  __shared__ unsigned long long s_kept, s_removed_c;
  __shared__ int s_total;
  unsigned long long removed = 0ull;  // this thread's word of the removed bitmap
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  const int nchunks = (nv + 63) >> 6;
  for (int c = 0; c < nchunks; ++c) {
    if (w == c) s_removed_c = removed;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long rem = s_removed_c, kept = 0ull;
      const int lim = min(64, nv - c * 64);
      int total = s_total;
      for (int b = 0; b < lim && total < post_topk; ++b) {
        if (!((rem >> b) & 1ull)) {
          kept |= 1ull << b;
          rem |= mask[((size_t)img * cpad + c * 64 + b) * words + c];
          const size_t src = (size_t)img * cpad + c * 64 + b, dst = (size_t)img * post_topk + total;
          reinterpret_cast<float4*>(out_box)[dst] = reinterpret_cast<const float4*>(sbox)[src];
          out_score[dst] = sscore[src];
          out_cat[dst] = scat[src];
          out_src[dst] = sperm[src];
          ++total;
        }
      }
      s_kept = kept;
      s_total = total;
    }
    __syncthreads();
    if (s_total >= post_topk) break;
    unsigned long long kept = s_kept;
    if (w > c && w < words) {
      while (kept) {
        const int b = __ffsll((long long)kept) - 1;
        kept &= kept - 1;
        removed |= mask[((size_t)img * cpad + c * 64 + b) * words + w];
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) out_count[img] = s_total;
}

// ------------------------------------------------------------------------------------------------
// K4a: RPN anchor matching (detectron2 Matcher with allow_low_quality_matches, rpn.py:label_and_sample_anchors)
//   pass 0: per anchor best IoU / argmax over GT, threshold labels; per-GT best IoU via atomicMax
//   pass 1: anchors whose IoU equals a GT's best IoU become positive
__global__ void __launch_bounds__(256)
rpn_match_kernel(aldi_rpn_levels L, const float* __restrict__ gt_boxes, const int* __restrict__ gt_counts, int gmax,
                 float lo, float hi, int pass, int* __restrict__ gt_best /*N*gmax, float bits*/,
                 signed char* __restrict__ labels, int* __restrict__ matched, int total_anchors) {
  extern __shared__ float s_gt[];  // gmax*5 : box + area, then gmax ints of per-GT best IoU bits
  int* s_best = reinterpret_cast<int*>(s_gt + (size_t)gmax * 5);
  const int img = blockIdx.y;
  const int g = min(gt_counts[img], gmax);
  for (int i = threadIdx.x; i < g; i += blockDim.x) {
    s_best[i] = 0;
    const float* b = gt_boxes + ((size_t)img * gmax + i) * 4;
    s_gt[i * 5 + 0] = b[0]; s_gt[i * 5 + 1] = b[1]; s_gt[i * 5 + 2] = b[2]; s_gt[i * 5 + 3] = b[3];
    s_gt[i * 5 + 4] = (b[2] - b[0]) * (b[3] - b[1]);
  }
  __syncthreads();
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < total_anchors; r += gridDim.x * blockDim.x) {
    // r is the anchor index in (level, h, w, a) order
    int lvl = 0, e = r;
    const int A = L.num_anchors;
    while (lvl < L.num_levels - 1 && e >= L.h[lvl] * L.w[lvl] * A) { e -= L.h[lvl] * L.w[lvl] * A; ++lvl; }
    float anc[4];
    anchor_box(L, lvl, e, anc);
    const float aarea = (anc[2] - anc[0]) * (anc[3] - anc[1]);
    const size_t o = (size_t)img * total_anchors + r;
    if (pass == 0) {
      float best = -1.f;
      int arg = 0;
      for (int i = 0; i < g; ++i) {
        const float v = d2_iou(&s_gt[i * 5], s_gt[i * 5 + 4], anc, aarea);
        if (v > best) { best = v; arg = i; }
        // per-GT maximum: warp reduce -> shared -> (at block end) one global atomic per GT
        const uint32_t active = __activemask();
        const int wmax = __reduce_max_sync(active, __float_as_int(v));  // v >= 0: int order == float order
        if (wmax > 0 && (int)(__ffs(active) - 1) == (int)(threadIdx.x & 31)) atomicMax(&s_best[i], wmax);
      }
      signed char lab;
      if (g == 0) { lab = 0; arg = 0; }          // Matcher on an empty matrix: all labels[0] == 0
      else if (best < lo) lab = 0;
      else if (best < hi) lab = -1;
      else lab = 1;
      labels[o] = lab;
      matched[o] = arg;
    } else {
      bool low_quality = false;
      for (int i = 0; i < g; ++i) {
        const float v = d2_iou(&s_gt[i * 5], s_gt[i * 5 + 4], anc, aarea);
        if (__float_as_int(v) == gt_best[(size_t)img * gmax + i]) low_quality = true;
      }
      if (low_quality) labels[o] = 1;
    }
  }
  if (pass == 0) {
    __syncthreads();
    for (int i = threadIdx.x; i < g; i += blockDim.x)
      if (s_best[i] > 0) atomicMax(gt_best + (size_t)img * gmax + i, s_best[i]);
  }
}

// K4b: subsample_labels on a dense label array: keep num_pos positives / num_neg negatives with the
// smallest hash keys, everything else -> -1.  grid = N, block = 1024.
__global__ void __launch_bounds__(1024)
subsample_kernel(signed char* __restrict__ labels, int n, int num_samples, float pos_fraction, uint32_t seed,
                 const uint32_t* __restrict__ salts, int* __restrict__ stats /*N*2: num_pos,num_neg*/) {
  __shared__ uint32_t s_hist[260];
  __shared__ int s_cnt[2];
  const int img = blockIdx.x;
  signed char* lab = labels + (size_t)img * n;
  const uint32_t salt = salts[img];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  int cp = 0, cn = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const signed char v = lab[i];
    cp += (v == 1);
    cn += (v == 0);
  }
  cp = warp_sum_i(cp);
  cn = warp_sum_i(cn);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt[0], cp); atomicAdd(&s_cnt[1], cn); }
  __syncthreads();
  int num_pos = (int)(num_samples * pos_fraction);
  num_pos = min(s_cnt[0], num_pos);
  int num_neg = min(s_cnt[1], num_samples - num_pos);
  __syncthreads();
  // positives: keep the num_pos smallest hashes == largest ~hash
  auto kpos = [&](int i, uint32_t& key) -> bool { key = ~sample_hash(seed, salt * 2u + 0u, (uint32_t)i); return lab[i] == 1; };
  auto kneg = [&](int i, uint32_t& key) -> bool { key = ~sample_hash(seed, salt * 2u + 1u, (uint32_t)i); return lab[i] == 0; };
  SelectResult rp, rn;
  rp.T = 0xFFFFFFFFu; rp.take_eq = 0; rp.count_eq = 0;
  rn = rp;
  if (num_pos > 0) rp = block_select(n, num_pos, kpos, s_hist);
  if (num_neg > 0) rn = block_select(n, num_neg, kneg, s_hist);
  // ties at the threshold are resolved lowest-index-first; with 32-bit hashes they are rare, so the
  // (sequential) ordered pass only runs when count_eq > take_eq
  __syncthreads();
  const bool tie_p = num_pos > 0 && rp.count_eq > rp.take_eq;
  const bool tie_n = num_neg > 0 && rn.count_eq > rn.take_eq;
  if ((tie_p || tie_n) && threadIdx.x == 0) {
    int tp = 0, tn = 0;
    for (int i = 0; i < n; ++i) {
      uint32_t key;
      if (tie_p && kpos(i, key) && key == rp.T) { if (tp >= rp.take_eq) lab[i] = -2; ++tp; }
      if (tie_n && kneg(i, key) && key == rn.T) { if (tn >= rn.take_eq) lab[i] = -2; ++tn; }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const signed char v = lab[i];
    signed char o = -1;
    uint32_t key;
    if (v == 1 && num_pos > 0) { kpos(i, key); if (key >= rp.T) o = 1; }
    else if (v == 0 && num_neg > 0) { kneg(i, key); if (key >= rn.T) o = 0; }
    lab[i] = o;
  }
  if (threadIdx.x == 0 && stats) { stats[2 * img] = num_pos; stats[2 * img + 1] = num_neg; }
}

// ------------------------------------------------------------------------------------------------
// K5: ROI heads label_and_sample_proposals for one image per block (detectron2 roi_heads.py:StandardROIHeads):
// append GT to the proposals, IoU-match (thr, no low-quality), sample `num_samples` with `pos_fraction`
// foreground, emit [fg..., bg...] in ascending candidate order.  block = 1024, candidates <= 4096.
__global__ void __launch_bounds__(1024)
roi_sample_kernel(const float* __restrict__ prop_box, const int* __restrict__ prop_count, int prop_stride,
                  const float* __restrict__ gt_boxes, const int* __restrict__ gt_classes,
                  const int* __restrict__ gt_counts, int gmax, float iou_thr, int num_classes, int num_samples,
                  float pos_fraction, uint32_t seed, const uint32_t* __restrict__ salts, int append_gt,
                  float* __restrict__ out_box, int* __restrict__ out_batch, int* __restrict__ out_class,
                  float* __restrict__ out_gtbox, int* __restrict__ out_src, int* __restrict__ out_count,
                  int* __restrict__ stats) {
  __shared__ signed char s_lab[4096];   // 1 fg, 0 bg, -1 ignore
  __shared__ short s_match[4096];
  __shared__ uint32_t s_hist[260];
  __shared__ int s_scan[40];
  __shared__ int s_cnt[2];
  const int img = blockIdx.x;
  const int np = min(prop_count[img], prop_stride);
  const int g = min(gt_counts[img], gmax);
  const int n = np + (append_gt ? g : 0);
  const float* gtb = gt_boxes + (size_t)img * gmax * 4;
  auto cand = [&](int i, float* b) {
    const float* p = (i < np) ? prop_box + ((size_t)img * prop_stride + i) * 4 : gtb + (size_t)(i - np) * 4;
    b[0] = p[0]; b[1] = p[1]; b[2] = p[2]; b[3] = p[3];
  };
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  int cp = 0, cn = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float b[4];
    cand(i, b);
    const float barea = (b[2] - b[0]) * (b[3] - b[1]);
    float best = -1.f;
    int arg = 0;
    for (int k = 0; k < g; ++k) {
      const float* q = gtb + (size_t)k * 4;
      const float v = d2_iou(q, (q[2] - q[0]) * (q[3] - q[1]), b, barea);
      if (v > best) { best = v; arg = k; }
    }
    // Matcher([thr],[0,1]) ; no GT -> every proposal is background
    const signed char lab = (g > 0 && best >= iou_thr) ? 1 : 0;
    s_lab[i] = lab;
    s_match[i] = (short)arg;
    cp += lab == 1;
    cn += lab == 0;
  }
  cp = warp_sum_i(cp);
  cn = warp_sum_i(cn);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt[0], cp); atomicAdd(&s_cnt[1], cn); }
  __syncthreads();
  int num_pos = min(s_cnt[0], (int)(num_samples * pos_fraction));
  int num_neg = min(s_cnt[1], num_samples - num_pos);
  const uint32_t salt = salts[img];
  auto kpos = [&](int i, uint32_t& key) -> bool { key = ~sample_hash(seed, salt * 2u + 0u, (uint32_t)i); return s_lab[i] == 1; };
  auto kneg = [&](int i, uint32_t& key) -> bool { key = ~sample_hash(seed, salt * 2u + 1u, (uint32_t)i); return s_lab[i] == 0; };
  const size_t obase = (size_t)img * num_samples;
  auto emit_at = [&](int row, int i) {
    float b[4];
    cand(i, b);
    const size_t o = obase + row;
    out_box[o * 4 + 0] = b[0]; out_box[o * 4 + 1] = b[1]; out_box[o * 4 + 2] = b[2]; out_box[o * 4 + 3] = b[3];
    out_batch[o] = img;
    out_src[o] = i;
    const int m = s_match[i];
    if (s_lab[i] == 1) {
      out_class[o] = gt_classes[(size_t)img * gmax + m];
    } else {
      out_class[o] = num_classes;
    }
    const float* q = (g > 0) ? gtb + (size_t)m * 4 : b;   // no GT: D2 falls back to the proposal box (unused by the loss)
    out_gtbox[o * 4 + 0] = q[0]; out_gtbox[o * 4 + 1] = q[1]; out_gtbox[o * 4 + 2] = q[2]; out_gtbox[o * 4 + 3] = q[3];
  };
  if (num_pos > 0) {
    const SelectResult r = block_select(n, num_pos, kpos, s_hist);
    auto emit = [&](int pos, int i, uint32_t) { emit_at(pos, i); };
    block_emit_selected(n, num_pos, r, kpos, emit, s_scan);
  }
  if (num_neg > 0) {
    const SelectResult r = block_select(n, num_neg, kneg, s_hist);
    auto emit = [&](int pos, int i, uint32_t) { emit_at(num_pos + pos, i); };
    block_emit_selected(n, num_neg, r, kneg, emit, s_scan);
  }
  __syncthreads();
  for (int row = num_pos + num_neg + threadIdx.x; row < num_samples; row += blockDim.x) {
    const size_t o = obase + row;
    out_box[o * 4 + 0] = 0.f; out_box[o * 4 + 1] = 0.f; out_box[o * 4 + 2] = 0.f; out_box[o * 4 + 3] = 0.f;
    out_gtbox[o * 4 + 0] = 0.f; out_gtbox[o * 4 + 1] = 0.f; out_gtbox[o * 4 + 2] = 0.f; out_gtbox[o * 4 + 3] = 0.f;
    out_batch[o] = img;
    out_class[o] = -1;  // padding row: ignored by every consumer
    out_src[o] = -1;
  }
  if (threadIdx.x == 0) {
    out_count[img] = num_pos + num_neg;
    if (stats) { stats[2 * img] = num_pos; stats[2 * img + 1] = num_neg; }
  }
}

// ------------------------------------------------------------------------------------------------
// K6: FastRCNNOutputLayers.inference front half: softmax, per-class box decode (weights 10,10,5,5), clip,
// score filter -> candidate list in (proposal, class) row-major order.  One block per image.
__global__ void __launch_bounds__(1024)
roi_candidates_kernel(const float* __restrict__ pred /*(N*P, pred_stride): K+1 logits then 4K deltas*/, int pred_stride,
                      const float* __restrict__ prop_box, const int* __restrict__ prop_count, int prop_stride,
                      int num_classes, const int* __restrict__ img_sizes, float score_thresh, float wx, float wy,
                      float ww, float wh, float clampv, float* __restrict__ cand_box, float* __restrict__ cand_score,
                      int* __restrict__ cand_cat, int* __restrict__ cand_src, int* __restrict__ cand_count,
                      int cand_stride) {
  __shared__ int s_scan[40];
  const int img = blockIdx.x;
  const int np = min(prop_count[img], prop_stride);
  const int K = num_classes;
  const float img_h = (float)img_sizes[2 * img], img_w = (float)img_sizes[2 * img + 1];
  const int total = np * K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) s_scan[32] = 0;
  __syncthreads();
  for (int i0 = 0; i0 < total; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    bool sel = false;
    float sc = 0.f, box[4] = {0.f, 0.f, 0.f, 0.f};
    int p = 0, c = 0;
    if (i < total) {
      p = i / K;
      c = i - p * K;
      const float* row = pred + ((size_t)img * prop_stride + p) * pred_stride;
      float mx = row[0];
      for (int k = 1; k <= K; ++k) mx = fmaxf(mx, row[k]);
      float sum = 0.f;
      for (int k = 0; k <= K; ++k) sum += expf(row[k] - mx);
      sc = expf(row[c] - mx) / sum;
      const float* d = row + (K + 1) + c * 4;
      const float* pb = prop_box + ((size_t)img * prop_stride + p) * 4;
      apply_deltas(pb, d[0], d[1], d[2], d[3], wx, wy, ww, wh, clampv, box);
      bool finite = isfinite(box[0]) && isfinite(box[1]) && isfinite(box[2]) && isfinite(box[3]);
      for (int k = 0; k <= K; ++k) finite = finite && isfinite(row[k]);
      clip_box(box, img_h, img_w);
      sel = finite && (sc > score_thresh);
    }
    const uint32_t b = __ballot_sync(0xffffffffu, sel);
    if (lane == 0) s_scan[warp] = __popc(b);
    __syncthreads();
    int before = s_scan[32];
    for (int w = 0; w < warp; ++w) before += s_scan[w];
    const int pos = before + __popc(b & ((1u << lane) - 1u));
    if (sel && pos < cand_stride) {
      const size_t o = (size_t)img * cand_stride + pos;
      cand_box[o * 4 + 0] = box[0]; cand_box[o * 4 + 1] = box[1]; cand_box[o * 4 + 2] = box[2]; cand_box[o * 4 + 3] = box[3];
      cand_score[o] = sc;
      cand_cat[o] = c;
      cand_src[o] = p;
    }
    int tot = 0;
    for (int w = 0; w < nwarps; ++w) tot += s_scan[w];
    __syncthreads();
    if (threadIdx.x == 0) s_scan[32] += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) cand_count[img] = min(s_scan[32], cand_stride);
}

// K7: aldi/pseudolabeler.py:51-67 process_bbox: keep detections with score > threshold, order preserved.
__global__ void pseudo_threshold_kernel(const float* __restrict__ box, const float* __restrict__ score,
                                        const int* __restrict__ cls, const int* __restrict__ count, int stride,
                                        float thr, float* __restrict__ gt_box, int* __restrict__ gt_cls,
                                        float* __restrict__ gt_score, int* __restrict__ gt_count, int gmax) {
  const int img = blockIdx.x;
  if (threadIdx.x != 0) return;
  const int c = min(count[img], stride);
  int o = 0;
  for (int i = 0; i < c && o < gmax; ++i) {
    const size_t s = (size_t)img * stride + i;
    if (score[s] > thr) {
      const size_t d = (size_t)img * gmax + o;
      gt_box[d * 4 + 0] = box[s * 4 + 0]; gt_box[d * 4 + 1] = box[s * 4 + 1];
      gt_box[d * 4 + 2] = box[s * 4 + 2]; gt_box[d * 4 + 3] = box[s * 4 + 3];
      gt_cls[d] = cls[s];
      gt_score[d] = score[s];
      ++o;
    }
  }
  gt_count[img] = o;
}

}  // namespace

// =================================================================================================
extern "C" int aldi_rpn_topk_decode(const float* rpn_out, const aldi_rpn_levels* L, int n_images, int pre_topk,
                                    const int* img_sizes, float* cand_box, float* cand_score, int* cand_cat,
                                    int* cand_idx, unsigned char* cand_valid, int cand_stride, int* err_flag,
                                    void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(rpn_out && L && img_sizes && cand_box && cand_score && cand_cat && cand_idx && cand_valid,
                 "aldi_rpn_topk_decode: null pointer");
  ALDI_CHECK_ARG(pre_topk > 0 && pre_topk <= 2048, "aldi_rpn_topk_decode: pre_topk must be in (0, 2048]");
  ALDI_CHECK_ARG(L->num_levels >= 1 && L->num_levels <= 5 && L->num_anchors >= 1 && L->num_anchors <= 3,
                 "aldi_rpn_topk_decode: bad level table");
  int need = 0;
  for (int l = 0; l < L->num_levels; ++l) {
    int nl = L->h[l] * L->w[l] * L->num_anchors;
    need += nl < pre_topk ? nl : pre_topk;
  }
  ALDI_CHECK_ARG(cand_stride >= need, "aldi_rpn_topk_decode: cand_stride %d < %d", cand_stride, need);
  rpn_topk_decode_kernel<<<dim3(L->num_levels, n_images), 1024, 2048 * sizeof(unsigned long long), stream>>>(
      rpn_out, *L, pre_topk, img_sizes, cand_box, cand_score, cand_cat, cand_idx, cand_valid, cand_stride, err_flag);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_rpn_topk_decode");
  return ALDI_OK;
}

extern "C" size_t aldi_nms_workspace_bytes(int n_images, int cand_stride) {
  size_t cpad = 64;
  while ((int)cpad < cand_stride) cpad <<= 1;
  size_t per = cpad * (16 + 4 + 4 + 4) + 256 + cpad * (cpad / 64) * 8;
  return per * (size_t)n_images + 1024;
}

extern "C" int aldi_nms_sorted(const float* cand_box, const float* cand_score, const int* cand_cat,
                               const unsigned char* cand_valid, const int* cand_count, int n_images, int cand_stride,
                               float iou_thresh, int post_topk, void* workspace, size_t workspace_bytes,
                               float* out_box, float* out_score, int* out_cat, int* out_src, int* out_count,
                               void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(cand_box && cand_score && cand_cat && workspace && out_box && out_score && out_cat && out_src &&
                     out_count, "aldi_nms_sorted: null pointer");
  ALDI_CHECK_ARG(n_images > 0 && cand_stride > 0 && cand_stride <= 16384, "aldi_nms_sorted: cand_stride must be <= 16384");
  ALDI_CHECK_ARG(workspace_bytes >= aldi_nms_workspace_bytes(n_images, cand_stride), "aldi_nms_sorted: workspace too small");
  int cpad = 64;
  while (cpad < cand_stride) cpad <<= 1;
  const int words = cpad / 64;
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  float* sbox = reinterpret_cast<float*>(ws);                 ws += (size_t)n_images * cpad * 16;
  float* sscore = reinterpret_cast<float*>(ws);               ws += (size_t)n_images * cpad * 4;
  int* scat = reinterpret_cast<int*>(ws);                     ws += (size_t)n_images * cpad * 4;
  int* sperm = reinterpret_cast<int*>(ws);                    ws += (size_t)n_images * cpad * 4;
  int* nvalid = reinterpret_cast<int*>(ws);                   ws += 256;
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws);

  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
    attr_set = true;
  }
  nms_sort_kernel<<<n_images, 1024, (size_t)cpad * 8, stream>>>(cand_box, cand_score, cand_cat, cand_valid, cand_count,
                                                                cand_stride, cpad, sbox, sscore, scat, sperm, nvalid);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_nms_sorted(sort)");
  nms_mask_kernel<<<dim3(words, words, n_images), 64, 0, stream>>>(sbox, scat, nvalid, cpad, iou_thresh, mask);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_nms_sorted(mask)");
  nms_scan_kernel<<<n_images, 256, 0, stream>>>(mask, sbox, sscore, scat, sperm, nvalid, cpad, post_topk, out_box,
                                                out_score, out_cat, out_src, out_count);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_nms_sorted(scan)");
  return ALDI_OK;
}

extern "C" int aldi_rpn_label_anchors(const aldi_rpn_levels* L, int n_images, const float* gt_boxes,
                                      const int* gt_counts, int gmax, float iou_lo, float iou_hi, int num_samples,
                                      float pos_fraction, unsigned int seed, const unsigned int* salts,
                                      int* gt_best_ws, signed char* labels, int* matched, int* stats, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(L && gt_boxes && gt_counts && salts && gt_best_ws && labels && matched,
                 "aldi_rpn_label_anchors: null pointer");
  ALDI_CHECK_ARG(gmax > 0 && gmax <= 2048, "aldi_rpn_label_anchors: gmax must be in (0, 2048]");
  int total = 0;
  for (int l = 0; l < L->num_levels; ++l) total += L->h[l] * L->w[l] * L->num_anchors;
  cudaError_t e = cudaMemsetAsync(gt_best_ws, 0, (size_t)n_images * gmax * sizeof(int), stream);
  if (e != cudaSuccess) { aldi_set_error("aldi_rpn_label_anchors: memset failed"); return ALDI_ERR_CUDA; }
  int gx = (total + 255) / 256;
  int cap = aldi_num_sms() * 4;
  if (gx > cap) gx = cap;
  for (int pass = 0; pass < 2; ++pass) {
    rpn_match_kernel<<<dim3(gx, n_images), 256, (size_t)gmax * 6 * sizeof(float), stream>>>(
        *L, gt_boxes, gt_counts, gmax, iou_lo, iou_hi, pass, gt_best_ws, labels, matched, total);
    ALDI_COUNT_LAUNCH();
    ALDI_CUDA_LAUNCH_CHECK("aldi_rpn_label_anchors(match)");
  }
  if (num_samples > 0) {
    subsample_kernel<<<n_images, 1024, 0, stream>>>(labels, total, num_samples, pos_fraction, seed, salts, stats);
    ALDI_COUNT_LAUNCH();
    ALDI_CUDA_LAUNCH_CHECK("aldi_rpn_label_anchors(subsample)");
  }
  return ALDI_OK;
}

extern "C" int aldi_roi_label_sample(const float* prop_box, const int* prop_count, int prop_stride, int n_images,
                                     const float* gt_boxes, const int* gt_classes, const int* gt_counts, int gmax,
                                     float iou_thresh, int num_classes, int num_samples, float pos_fraction,
                                     unsigned int seed, const unsigned int* salts, int append_gt, float* out_box,
                                     int* out_batch, int* out_class, float* out_gtbox, int* out_src, int* out_count,
                                     int* stats, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(prop_box && prop_count && gt_boxes && gt_classes && gt_counts && salts && out_box && out_batch &&
                     out_class && out_gtbox && out_src && out_count, "aldi_roi_label_sample: null pointer");
  ALDI_CHECK_ARG(prop_stride + gmax <= 4096, "aldi_roi_label_sample: proposals + gt must be <= 4096 per image");
  roi_sample_kernel<<<n_images, 1024, 0, stream>>>(prop_box, prop_count, prop_stride, gt_boxes, gt_classes, gt_counts,
                                                   gmax, iou_thresh, num_classes, num_samples, pos_fraction, seed,
                                                   salts, append_gt, out_box, out_batch, out_class, out_gtbox,
                                                   out_src, out_count, stats);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_roi_label_sample");
  return ALDI_OK;
}

extern "C" int aldi_roi_inference_candidates(const float* pred, int pred_stride, const float* prop_box,
                                             const int* prop_count, int prop_stride, int n_images, int num_classes,
                                             const int* img_sizes, float score_thresh, const float* h_weights4,
                                             float scale_clamp, float* cand_box, float* cand_score, int* cand_cat,
                                             int* cand_src, int* cand_count, int cand_stride, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(pred && prop_box && prop_count && img_sizes && h_weights4 && cand_box && cand_score && cand_cat &&
                     cand_src && cand_count, "aldi_roi_inference_candidates: null pointer");
  roi_candidates_kernel<<<n_images, 1024, 0, stream>>>(pred, pred_stride, prop_box, prop_count, prop_stride,
                                                       num_classes, img_sizes, score_thresh, h_weights4[0],
                                                       h_weights4[1], h_weights4[2], h_weights4[3], scale_clamp,
                                                       cand_box, cand_score, cand_cat, cand_src, cand_count,
                                                       cand_stride);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_roi_inference_candidates");
  return ALDI_OK;
}

extern "C" int aldi_pseudo_label_threshold(const float* det_box, const float* det_score, const int* det_class,
                                           const int* det_count, int det_stride, int n_images, float threshold,
                                           float* gt_box, int* gt_class, float* gt_score, int* gt_count, int gmax,
                                           void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(det_box && det_score && det_class && det_count && gt_box && gt_class && gt_score && gt_count,
                 "aldi_pseudo_label_threshold: null pointer");
  pseudo_threshold_kernel<<<n_images, 32, 0, stream>>>(det_box, det_score, det_class, det_count, det_stride, threshold,
                                                       gt_box, gt_class, gt_score, gt_count, gmax);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_pseudo_label_threshold");
  return ALDI_OK;
}
