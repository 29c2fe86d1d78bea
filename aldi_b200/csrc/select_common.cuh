// Device helpers shared by the selection kernels (select.cu, select_mb.cu).  Both files are compiled with
// -fmad=false: these expressions decide index sets and must round exactly like the reference's.
#pragma once
#include "common.cuh"
#include "../../include/aldi_b200.h"
#include <float.h>

namespace aldi_sel {

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
__device__ inline void block_bitonic_desc(unsigned long long* s, int n_pow2) {
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


}  // namespace aldi_sel
