// Multi-block exact selection: the two "choose k of half a million" steps of the detector, spread over the
// whole GPU instead of one thread block per image.
//
//   * RPN per-level top-k objectness (detectron2 find_top_rpn_proposals: logits_i.topk(pre_nms_topk)) followed by
//     decode + clip of the selected anchors;
//   * subsample_labels over the 523 776 anchor labels of an image (detectron2 sampling.py, reached from
//     RPN.label_and_sample_anchors; aldi/distill.py:200-202 calls it a second time on the teacher with pseudo GT).
//
// Both are an exact radix select over 32-bit keys in three grid-wide passes (12 + 12 + 8 bits): every block
// histograms its slice into shared memory, flushes to a per-segment global histogram, and the LAST block to arrive
// (atomic ticket) resolves the digit and publishes {prefix, remaining} for the next pass.  Selection = keys > T plus
// the `take_eq` lowest-index keys == T; ties straddling the cut (count_eq > take_eq) take an ordered slow path in a
// single block, which only degenerate inputs (constant logits, 32-bit hash collisions) ever reach.
// Compiled with -fmad=false like select.cu: decode / clip round exactly like the reference expressions.
#include "select_common.cuh"

using namespace aldi_sel;

namespace {

constexpr int kThreads = 1024;
constexpr int kHistWords = 4096 + 4096 + 256;  // the three digit histograms of one segment

struct SelState {          // one per segment, zero-initialised by the caller's memset
  uint32_t prefix;         // resolved high bits; after pass 2 the threshold key T
  int remaining;           // still to take among keys matching the prefix; after pass 2 = take_eq
  int count_eq;            // keys == T
  int k;                   // requested size of the selection
  unsigned int done[3];    // block tickets of the three passes
  unsigned int cursor;     // append cursor (collect pass)
  int pad[8];
};
static_assert(sizeof(SelState) == 64, "SelState is 64 bytes");

__device__ __forceinline__ int digit_of(uint32_t key, int pass) {
  return pass == 0 ? (int)(key >> 20) : pass == 1 ? (int)((key >> 8) & 0xFFFu) : (int)(key & 0xFFu);
}
__device__ __forceinline__ bool prefix_match(uint32_t key, uint32_t prefix, int pass) {
  return pass == 0 ? true : pass == 1 ? ((key >> 20) == (prefix >> 20)) : ((key >> 8) == (prefix >> 8));
}
__device__ __forceinline__ int nbins_of(int pass) { return pass == 2 ? 256 : 4096; }
__device__ __forceinline__ int shift_of(int pass) { return pass == 0 ? 20 : pass == 1 ? 8 : 0; }
__device__ __forceinline__ int hist_off(int pass) { return pass == 0 ? 0 : pass == 1 ? 4096 : 8192; }

// exclusive prefix sum over the block (blockDim.x == 1024); s_warp has >= 33 ints
__device__ int block_excl_scan(int v, int* s_warp) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  if (w == 0) {
    int x = s_warp[lane];
    int xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) xi += t;
    }
    s_warp[lane] = xi - x;
    if (lane == 31) s_warp[32] = xi;
  }
  __syncthreads();
  return s_warp[w] + inc - v;
}

// Largest bin d with (sum of bins >= d) >= remaining, for 1 <= remaining <= total.
// s_res[0] = d, s_res[1] = remaining - (sum of bins > d), s_res[2] = hist[d].  All threads call; result after return.
__device__ void resolve_digit(const uint32_t* g_hist, int nbins, int remaining, int* s_res, int* s_warp) {
  if (threadIdx.x == 0) { s_res[0] = nbins; s_res[1] = 0; s_res[2] = 0; }
  const int per = (nbins + kThreads - 1) / kThreads;  // 4 or 1
  const int j0 = threadIdx.x * per;
  uint32_t h[4] = {0u, 0u, 0u, 0u};
  int local = 0;
  for (int u = 0; u < per; ++u) {
    const int j = j0 + u;  // j-th bin from the TOP
    if (j < nbins) { h[u] = __ldcg(g_hist + (nbins - 1 - j)); local += (int)h[u]; }
  }
  int cum = block_excl_scan(local, s_warp);
  for (int u = 0; u < per; ++u) {
    if (h[u] && cum < remaining && cum + (int)h[u] >= remaining) {
      s_res[0] = nbins - 1 - (j0 + u);
      s_res[1] = remaining - cum;
      s_res[2] = (int)h[u];
    }
    cum += (int)h[u];
  }
  __syncthreads();
}

// one histogram pass over the block's slice of a segment; flushes into g_hist (this pass's bins of the segment)
template <typename KeyFn>
__device__ void hist_slice(int pass, int n, KeyFn kf, uint32_t prefix, uint32_t* s_hist, uint32_t* g_hist) {
  const int nb = nbins_of(pass);
  for (int i = threadIdx.x; i < nb; i += kThreads) s_hist[i] = 0;
  __syncthreads();
  for (int i0 = blockIdx.x * kThreads; i0 < n; i0 += gridDim.x * kThreads) {
    const int i = i0 + threadIdx.x;
    uint32_t key = 0;
    const bool ok = (i < n) && kf(i, key) && prefix_match(key, prefix, pass);
    const uint32_t bin = (uint32_t)digit_of(key, pass);
    // warp-aggregated update: the top digit of float keys falls into a handful of bins
    const uint32_t active = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const uint32_t peers = __match_any_sync(active, bin);
      if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&s_hist[bin], (uint32_t)__popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += kThreads)
    if (s_hist[i]) atomicAdd(&g_hist[i], s_hist[i]);
}

// true in every thread of exactly one block per segment and pass: the last one to get here
__device__ bool last_block(unsigned int* ticket, int* s_flag) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *s_flag = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (*s_flag) __threadfence();
  return *s_flag != 0;
}

// after the last pass-`pass` block resolved the digit: fold it into the state
__device__ void publish(SelState* st, int pass, const int* s_res) {
  if (threadIdx.x == 0) {
    st->prefix |= (uint32_t)s_res[0] << shift_of(pass);
    st->remaining = s_res[1];
    if (pass == 2) st->count_eq = s_res[2];
  }
}

// =================================================================================================
// RPN top-k: segment = (image, level); key = order-preserving bits of the objectness logit
// =================================================================================================
struct TopkSeg {
  const float* base;
  int n, k, A, ch_stride;
};
__device__ __forceinline__ TopkSeg topk_seg(const float* rpn_out, const aldi_rpn_levels& L, int pre_topk, int img, int lvl) {
  TopkSeg s;
  s.A = L.num_anchors;
  s.ch_stride = L.ch_stride;
  s.n = L.h[lvl] * L.w[lvl] * s.A;
  s.k = s.n < pre_topk ? s.n : pre_topk;
  s.base = rpn_out + ((size_t)img * L.total_locs + L.loc_off[lvl]) * L.ch_stride;
  return s;
}

__global__ void __launch_bounds__(kThreads)
topk_hist_kernel(const float* __restrict__ rpn_out, aldi_rpn_levels L, int pre_topk, int pass, uint32_t* __restrict__ hist,
                 SelState* __restrict__ states) {
  __shared__ uint32_t s_hist[4096];
  __shared__ int s_warp[33], s_res[3], s_flag;
  const int lvl = blockIdx.y, img = blockIdx.z;
  const int seg = img * L.num_levels + lvl;
  const TopkSeg sg = topk_seg(rpn_out, L, pre_topk, img, lvl);
  SelState* st = states + seg;
  uint32_t* gh = hist + (size_t)seg * kHistWords + hist_off(pass);
  const uint32_t prefix = pass ? st->prefix : 0u;
  const int remaining = pass ? st->remaining : sg.k;
  auto kf = [&](int i, uint32_t& key) -> bool {
    const int loc = i / sg.A, a = i - loc * sg.A;
    key = fkey(__ldg(sg.base + (size_t)loc * sg.ch_stride + a));
    return true;
  };
  hist_slice(pass, sg.n, kf, prefix, s_hist, gh);
  if (last_block(&st->done[pass], &s_flag)) {
    resolve_digit(gh, nbins_of(pass), remaining, s_res, s_warp);
    if (threadIdx.x == 0 && pass == 0) st->k = sg.k;
    publish(st, pass, s_res);
  }
}

// keys > T (and keys == T when they all fit) -> unordered candidate list of the segment
__global__ void __launch_bounds__(kThreads)
topk_collect_kernel(const float* __restrict__ rpn_out, aldi_rpn_levels L, int pre_topk, SelState* __restrict__ states,
                    unsigned long long* __restrict__ list, int list_stride) {
  const int lvl = blockIdx.y, img = blockIdx.z;
  const int seg = img * L.num_levels + lvl;
  const TopkSeg sg = topk_seg(rpn_out, L, pre_topk, img, lvl);
  SelState* st = states + seg;
  const uint32_t T = st->prefix;
  const bool eq_all = st->count_eq == st->remaining;
  unsigned long long* out = list + (size_t)seg * list_stride;
  for (int i0 = blockIdx.x * kThreads; i0 < sg.n; i0 += gridDim.x * kThreads) {
    const int i = i0 + threadIdx.x;
    uint32_t key = 0;
    bool sel = false;
    if (i < sg.n) {
      const int loc = i / sg.A, a = i - loc * sg.A;
      key = fkey(__ldg(sg.base + (size_t)loc * sg.ch_stride + a));
      sel = key > T || (key == T && eq_all);
    }
    const uint32_t b = __ballot_sync(0xffffffffu, sel);
    if (b) {
      const int lane = threadIdx.x & 31;
      unsigned int basepos = 0;
      if (lane == (int)(__ffs(b) - 1)) basepos = atomicAdd(&st->cursor, (unsigned int)__popc(b));
      basepos = __shfl_sync(0xffffffffu, basepos, __ffs(b) - 1);
      if (sel) {
        const unsigned int pos = basepos + __popc(b & ((1u << lane) - 1u));
        if ((int)pos < list_stride)
          out[pos] = ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
      }
    }
  }
}

// per segment: (tie slow path,) sort the <= 2048 selected keys descending, decode + clip the anchors
__global__ void __launch_bounds__(kThreads)
topk_finish_kernel(const float* __restrict__ rpn_out, aldi_rpn_levels L, int pre_topk, const SelState* __restrict__ states,
                   const unsigned long long* __restrict__ list, int list_stride, const int* __restrict__ img_sizes,
                   float* __restrict__ cand_box, float* __restrict__ cand_score, int* __restrict__ cand_cat,
                   int* __restrict__ cand_idx, unsigned char* __restrict__ cand_valid, int cand_stride,
                   int* __restrict__ err_flag) {
  __shared__ unsigned long long s_keys[2048];
  __shared__ int s_scan[40];
  const int lvl = blockIdx.x, img = blockIdx.y;
  const int seg = img * L.num_levels + lvl;
  const TopkSeg sg = topk_seg(rpn_out, L, pre_topk, img, lvl);
  const SelState st = states[seg];
  const int k = sg.k, A = sg.A;
  int cand_off = 0;
  for (int l = 0; l < lvl; ++l) {
    const int nl = L.h[l] * L.w[l] * A;
    cand_off += nl < pre_topk ? nl : pre_topk;
  }
  const bool straddle = st.count_eq != st.remaining;
  const int n_listed = straddle ? k - st.remaining : k;  // keys > T only when ties straddle the cut
  const unsigned long long* in = list + (size_t)seg * list_stride;
  for (int i = threadIdx.x; i < 2048; i += kThreads) s_keys[i] = (i < n_listed) ? in[i] : 0ull;
  __syncthreads();
  if (straddle) {
    // the first `remaining` logits equal to T, in index order (lowest index wins, like a stable sort)
    auto kf = [&](int i, uint32_t& key) -> bool {
      const int loc = i / A, a = i - loc * A;
      key = fkey(__ldg(sg.base + (size_t)loc * sg.ch_stride + a));
      return key == st.prefix;
    };
    SelectResult r;
    r.T = st.prefix; r.take_eq = st.remaining; r.count_eq = st.count_eq;
    auto emit = [&](int pos, int i, uint32_t key) {
      s_keys[n_listed + pos] = ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
    };
    block_emit_selected(sg.n, st.remaining, r, kf, emit, s_scan);
  }
  block_bitonic_desc(s_keys, 2048);
  const float img_h = (float)img_sizes[2 * img], img_w = (float)img_sizes[2 * img + 1];
  for (int j = threadIdx.x; j < k; j += kThreads) {
    const int e = (int)(0xFFFFFFFFu - (uint32_t)(s_keys[j] & 0xFFFFFFFFull));
    const int loc = e / A, a = e - loc * A;
    const float* row = sg.base + (size_t)loc * sg.ch_stride;
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

// =================================================================================================
// subsample_labels: segment = (image, class); class 0 = positives (label 1), class 1 = negatives (label 0);
// key = ~hash so that the k SMALLEST hashes are the k largest keys.  Must match aldi_b200/sampling.py.
// =================================================================================================
__device__ __forceinline__ uint32_t sub_key(uint32_t seed, uint32_t salt, int cls, int i) {
  return ~sample_hash(seed, salt * 2u + (uint32_t)cls, (uint32_t)i);
}

__global__ void __launch_bounds__(kThreads)
subsample_hist_kernel(const signed char* __restrict__ labels, int n, int pass, const uint32_t* __restrict__ seed_ptr,
                      const uint32_t* __restrict__ salts, int num_samples, float pos_fraction,
                      uint32_t* __restrict__ hist, SelState* __restrict__ states) {
  __shared__ uint32_t s_hist[2][4096];
  __shared__ int s_warp[33], s_res[3], s_flag;
  const int img = blockIdx.y;
  const uint32_t seed = *seed_ptr;
  const signed char* lab = labels + (size_t)img * n;
  const uint32_t salt = salts[img];
  SelState* st = states + 2 * img;  // [pos, neg]
  uint32_t* gh0 = hist + (size_t)(2 * img) * kHistWords + hist_off(pass);
  uint32_t* gh1 = gh0 + kHistWords;
  const int nb = nbins_of(pass);
  const uint32_t pre0 = pass ? st[0].prefix : 0u, pre1 = pass ? st[1].prefix : 0u;
  const bool on0 = pass == 0 || st[0].k > 0, on1 = pass == 0 || st[1].k > 0;
  for (int i = threadIdx.x; i < nb; i += kThreads) { s_hist[0][i] = 0; s_hist[1][i] = 0; }
  __syncthreads();
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const signed char v = lab[i];
    if (v == 1 && on0) {
      const uint32_t key = sub_key(seed, salt, 0, i);
      if (prefix_match(key, pre0, pass)) atomicAdd(&s_hist[0][digit_of(key, pass)], 1u);
    } else if (v == 0 && on1) {
      const uint32_t key = sub_key(seed, salt, 1, i);
      if (prefix_match(key, pre1, pass)) atomicAdd(&s_hist[1][digit_of(key, pass)], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += kThreads) {
    if (s_hist[0][i]) atomicAdd(&gh0[i], s_hist[0][i]);
    if (s_hist[1][i]) atomicAdd(&gh1[i], s_hist[1][i]);
  }
  if (last_block(&st[0].done[pass], &s_flag)) {
    int k0, k1;
    if (pass == 0) {
      // class sizes = histogram totals; detectron2 subsample_labels: num_pos = min(#pos, int(S*f)), num_neg = min(#neg, S-num_pos)
      int c0 = 0, c1 = 0;
      for (int i = threadIdx.x; i < nb; i += kThreads) { c0 += (int)__ldcg(gh0 + i); c1 += (int)__ldcg(gh1 + i); }
      const int e0 = block_excl_scan(c0, s_warp);
      (void)e0;
      const int t0 = s_warp[32];
      __syncthreads();
      const int e1 = block_excl_scan(c1, s_warp);
      (void)e1;
      const int t1 = s_warp[32];
      __syncthreads();
      k0 = min(t0, (int)(num_samples * pos_fraction));
      k1 = min(t1, num_samples - k0);
      if (threadIdx.x == 0) { st[0].k = k0; st[1].k = k1; }
    } else {
      k0 = st[0].k; k1 = st[1].k;
    }
    if (k0 > 0) {
      resolve_digit(gh0, nb, pass ? st[0].remaining : k0, s_res, s_warp);
      publish(&st[0], pass, s_res);
      __syncthreads();
    }
    if (k1 > 0) {
      resolve_digit(gh1, nb, pass ? st[1].remaining : k1, s_res, s_warp);
      publish(&st[1], pass, s_res);
    }
  }
}

// final labels: selected positives stay 1, selected negatives stay 0, everything else -1; ties that straddle the cut
// are parked as -2 (pos) / -3 (neg) for subsample_tie_kernel
__global__ void __launch_bounds__(kThreads)
subsample_apply_kernel(signed char* __restrict__ labels, int n, const uint32_t* __restrict__ seed_ptr,
                       const uint32_t* __restrict__ salts, const SelState* __restrict__ states, int* __restrict__ stats) {
  const int img = blockIdx.y;
  const uint32_t seed = *seed_ptr;
  signed char* lab = labels + (size_t)img * n;
  const uint32_t salt = salts[img];
  const SelState s0 = states[2 * img], s1 = states[2 * img + 1];
  const bool str0 = s0.count_eq != s0.remaining, str1 = s1.count_eq != s1.remaining;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const signed char v = lab[i];
    signed char o = -1;
    if (v == 1 && s0.k > 0) {
      const uint32_t key = sub_key(seed, salt, 0, i);
      if (key > s0.prefix) o = 1;
      else if (key == s0.prefix) o = str0 ? -2 : 1;
    } else if (v == 0 && s1.k > 0) {
      const uint32_t key = sub_key(seed, salt, 1, i);
      if (key > s1.prefix) o = 0;
      else if (key == s1.prefix) o = str1 ? -3 : 0;
    }
    lab[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && stats) { stats[2 * img] = s0.k; stats[2 * img + 1] = s1.k; }
}

// rare: equal hashes at the threshold -> the lowest `take_eq` indices win (serial scan by one thread per image)
__global__ void subsample_tie_kernel(signed char* __restrict__ labels, int n, const SelState* __restrict__ states) {
  const int img = blockIdx.x;
  const SelState s0 = states[2 * img], s1 = states[2 * img + 1];
  const bool str0 = s0.k > 0 && s0.count_eq != s0.remaining, str1 = s1.k > 0 && s1.count_eq != s1.remaining;
  if ((!str0 && !str1) || threadIdx.x != 0) return;
  signed char* lab = labels + (size_t)img * n;
  int t0 = 0, t1 = 0;
  for (int i = 0; i < n; ++i) {
    const signed char v = lab[i];
    if (v == -2) { lab[i] = (t0 < s0.remaining) ? 1 : -1; ++t0; }
    else if (v == -3) { lab[i] = (t1 < s1.remaining) ? 0 : -1; ++t1; }
  }
}

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace

// =================================================================================================
extern "C" size_t aldi_rpn_topk_workspace_bytes(int n_images, int num_levels) {
  const size_t segs = (size_t)n_images * num_levels;
  return align256(segs * kHistWords * 4) + align256(segs * sizeof(SelState)) + align256(segs * 2048 * 8) + 256;
}

extern "C" int aldi_rpn_topk_decode(const float* rpn_out, const aldi_rpn_levels* L, int n_images, int pre_topk,
                                    const int* img_sizes, float* cand_box, float* cand_score, int* cand_cat,
                                    int* cand_idx, unsigned char* cand_valid, int cand_stride, int* err_flag,
                                    void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(rpn_out && L && img_sizes && cand_box && cand_score && cand_cat && cand_idx && cand_valid && workspace,
                 "aldi_rpn_topk_decode: null pointer");
  ALDI_CHECK_ARG(pre_topk > 0 && pre_topk <= 2048, "aldi_rpn_topk_decode: pre_topk must be in (0, 2048]");
  ALDI_CHECK_ARG(L->num_levels >= 1 && L->num_levels <= 5 && L->num_anchors >= 1 && L->num_anchors <= 3,
                 "aldi_rpn_topk_decode: bad level table");
  ALDI_CHECK_ARG(workspace_bytes >= aldi_rpn_topk_workspace_bytes(n_images, L->num_levels),
                 "aldi_rpn_topk_decode: workspace too small");
  int need = 0, nmax = 0;
  for (int l = 0; l < L->num_levels; ++l) {
    int nl = L->h[l] * L->w[l] * L->num_anchors;
    need += nl < pre_topk ? nl : pre_topk;
    nmax = nl > nmax ? nl : nmax;
  }
  ALDI_CHECK_ARG(cand_stride >= need, "aldi_rpn_topk_decode: cand_stride %d < %d", cand_stride, need);
  const size_t segs = (size_t)n_images * L->num_levels;
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  uint32_t* hist = reinterpret_cast<uint32_t*>(ws);
  SelState* states = reinterpret_cast<SelState*>(ws + align256(segs * kHistWords * 4));
  unsigned long long* list =
      reinterpret_cast<unsigned long long*>(ws + align256(segs * kHistWords * 4) + align256(segs * sizeof(SelState)));
  cudaError_t e = cudaMemsetAsync(ws, 0, align256(segs * kHistWords * 4) + align256(segs * sizeof(SelState)), stream);
  if (e != cudaSuccess) { aldi_set_error("aldi_rpn_topk_decode: memset failed"); return ALDI_ERR_CUDA; }
  int bx = (nmax + kThreads * 8 - 1) / (kThreads * 8);  // ~8 elements per thread on the largest level
  bx = bx < 1 ? 1 : bx > 64 ? 64 : bx;
  const dim3 grid(bx, L->num_levels, n_images);
  for (int pass = 0; pass < 3; ++pass) {
    topk_hist_kernel<<<grid, kThreads, 0, stream>>>(rpn_out, *L, pre_topk, pass, hist, states);
    ALDI_COUNT_LAUNCH();
    ALDI_CUDA_LAUNCH_CHECK("aldi_rpn_topk_decode(hist)");
  }
  topk_collect_kernel<<<grid, kThreads, 0, stream>>>(rpn_out, *L, pre_topk, states, list, 2048);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_rpn_topk_decode(collect)");
  topk_finish_kernel<<<dim3(L->num_levels, n_images), kThreads, 0, stream>>>(
      rpn_out, *L, pre_topk, states, list, 2048, img_sizes, cand_box, cand_score, cand_cat, cand_idx, cand_valid,
      cand_stride, err_flag);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_rpn_topk_decode(finish)");
  return ALDI_OK;
}

// subsample_labels over dense label arrays (N, n): called by aldi_rpn_label_anchors (select.cu)
size_t aldi_subsample_workspace_bytes(int n_images) {
  const size_t segs = (size_t)n_images * 2;
  return align256(segs * kHistWords * 4) + align256(segs * sizeof(SelState)) + 256;
}

int aldi_subsample_labels(signed char* labels, int n_images, int n, int num_samples, float pos_fraction,
                          const unsigned int* d_seed, const unsigned int* salts, int* stats, void* workspace,
                          cudaStream_t stream) {
  const size_t segs = (size_t)n_images * 2;
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  uint32_t* hist = reinterpret_cast<uint32_t*>(ws);
  SelState* states = reinterpret_cast<SelState*>(ws + align256(segs * kHistWords * 4));
  cudaError_t e = cudaMemsetAsync(ws, 0, align256(segs * kHistWords * 4) + align256(segs * sizeof(SelState)), stream);
  if (e != cudaSuccess) { aldi_set_error("aldi_rpn_label_anchors: memset failed"); return ALDI_ERR_CUDA; }
  int bx = (n + kThreads * 8 - 1) / (kThreads * 8);
  bx = bx < 1 ? 1 : bx > 64 ? 64 : bx;
  const dim3 grid(bx, n_images);
  for (int pass = 0; pass < 3; ++pass) {
    subsample_hist_kernel<<<grid, kThreads, 0, stream>>>(labels, n, pass, d_seed, salts, num_samples, pos_fraction, hist,
                                                         states);
    ALDI_COUNT_LAUNCH();
    ALDI_CUDA_LAUNCH_CHECK("aldi_rpn_label_anchors(subsample hist)");
  }
  subsample_apply_kernel<<<grid, kThreads, 0, stream>>>(labels, n, d_seed, salts, states, stats);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_rpn_label_anchors(subsample apply)");
  subsample_tie_kernel<<<n_images, 32, 0, stream>>>(labels, n, states);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_rpn_label_anchors(subsample tie)");
  return ALDI_OK;
}
