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
#include "select_common.cuh"

using namespace aldi_sel;

namespace {

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
    ccat[threadIdx.x] = scat ? scat[(size_t)img * cpad + cj] : 0;   // scat == nullptr: one category per segment
  }
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i >= nv) return;
  const float4 bi = reinterpret_cast<const float4*>(sbox)[(size_t)img * cpad + i];
  const int ci = scat ? scat[(size_t)img * cpad + i] : 0;
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

// K3c: greedy scan in score order, 64 boxes at a time.  grid = N, block = 256, dynamic smem = post_topk ints.
// "Pull" formulation: the kept boxes so far live in a shared list; for chunk c every thread ORs column word c of
// its share of the kept rows (independent loads, one round trip), a block OR-reduction yields the chunk's
// `removed` word, one thread resolves the 64 boxes serially against the diagonal words staged in shared memory
// (prefetched a chunk ahead), and the newly kept boxes are written out in parallel.
__global__ void __launch_bounds__(256)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const float* __restrict__ sbox,
                const float* __restrict__ sscore, const int* __restrict__ scat, const int* __restrict__ sperm,
                const int* __restrict__ nvalid, int cpad, int post_topk, float* __restrict__ out_box,
                float* __restrict__ out_score, int* __restrict__ out_cat, int* __restrict__ out_src,
                int* __restrict__ out_count, int* __restrict__ kept_out) {
  extern __shared__ int s_keptlist[];  // sorted-order indices of the kept boxes
  const int img = blockIdx.x;
  const int nv = nvalid[img];
  const int words = cpad >> 6;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const unsigned long long* mrow = mask + (size_t)img * cpad * words;
  __shared__ unsigned long long s_diag[64];
  __shared__ unsigned long long s_part[8];
  __shared__ unsigned long long s_kept;
  __shared__ int s_total, s_base;
  if (t == 0) s_total = 0;
  __syncthreads();
  const int nchunks = (nv + 63) >> 6;
  unsigned long long diag_next = 0ull;
  if (t < 64 && t < nv) diag_next = mrow[(size_t)t * words];
  for (int c = 0; c < nchunks; ++c) {
    const int total0 = s_total;
    unsigned long long part = 0ull;
    for (int k = t; k < total0; k += 256) part |= mrow[(size_t)s_keptlist[k] * words + c];
    uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)part);
    uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(part >> 32));
    if (lane == 0) s_part[warp] = ((unsigned long long)hi << 32) | lo;
    if (t < 64) s_diag[t] = diag_next;
    __syncthreads();
    if (t < 64 && c + 1 < nchunks) {
      const int r = (c + 1) * 64 + t;
      diag_next = (r < nv) ? mrow[(size_t)r * words + c + 1] : 0ull;
    }
    if (t == 0) {
      unsigned long long rem = 0ull, kept = 0ull;
#pragma unroll
      for (int i = 0; i < 8; ++i) rem |= s_part[i];
      const int lim = min(64, nv - c * 64);
      int total = total0;
      s_base = total;
      for (int b = 0; b < lim && total < post_topk; ++b) {
        if (!((rem >> b) & 1ull)) {
          kept |= 1ull << b;
          rem |= s_diag[b];
          s_keptlist[total] = c * 64 + b;
          ++total;
        }
      }
      s_kept = kept;
      s_total = total;
    }
    __syncthreads();
    const unsigned long long kept = s_kept;
    if (!kept_out && t < 64 && ((kept >> t) & 1ull)) {
      const int rank = __popcll(kept & ((1ull << t) - 1ull));
      const size_t src = (size_t)img * cpad + c * 64 + t, dst = (size_t)img * post_topk + s_base + rank;
      reinterpret_cast<float4*>(out_box)[dst] = reinterpret_cast<const float4*>(sbox)[src];
      out_score[dst] = sscore[src];
      out_cat[dst] = scat[src];
      out_src[dst] = sperm[src];
    }
    if (s_total >= post_topk) break;
  }
  __syncthreads();
  if (t == 0) out_count[img] = s_total;
  if (kept_out)   // segmented NMS: the merge kernel gathers; hand it the kept positions (ascending = score order)
    for (int k = t; k < s_total; k += 256) kept_out[(size_t)img * post_topk + k] = s_keptlist[k];
}

// ------------------------------------------------------------------------------------------------
// Segmented NMS (RPN: category = FPN level, and every level's candidates arrive already sorted by score from
// aldi_rpn_topk_decode).  Boxes of different segments never suppress each other, so the greedy scans of the
// segments are independent: instead of ONE 16K-key sort, a (16K)^2/2 bit matrix of which 80 % is cross-level and
// a 150-chunk serial scan per image, each (image, level) gets a stable compaction, a 2K x 2K matrix and a 32-chunk
// scan, all in parallel, and a rank-by-binary-search merge restores the global score order and the keep[:post_topk]
// cut.  Result identical to aldi_nms_sorted (same keys: score descending, candidate position ascending).
struct SegTable {
  int num_seg;
  int off[8], len[8];
};

// stable compaction of the valid candidates of one (image, segment): block = 1024
__global__ void __launch_bounds__(1024)
nms_seg_compact_kernel(const float* __restrict__ box, const float* __restrict__ score,
                       const unsigned char* __restrict__ valid, int cand_stride, SegTable segs, int cap,
                       float* __restrict__ sbox, float* __restrict__ sscore, int* __restrict__ sperm,
                       int* __restrict__ nvalid) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int seg = blockIdx.x, img = blockIdx.y;
  const int off = segs.off[seg], len = segs.len[seg];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t dst0 = ((size_t)img * segs.num_seg + seg) * cap;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < len; c0 += 1024) {
    const int i = c0 + threadIdx.x;
    const size_t src = (size_t)img * cand_stride + off + i;
    const bool ok = i < len && (!valid || valid[src]);
    const unsigned b = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = __popc(b);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (ok) {
      const size_t dst = dst0 + before + __popc(b & ((1u << lane) - 1u));
      reinterpret_cast<float4*>(sbox)[dst] = reinterpret_cast<const float4*>(box)[src];
      sscore[dst] = score[src];
      sperm[dst] = off + i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) nvalid[img * segs.num_seg + seg] = s_base;
}

// global rank of every kept box = its index in its own segment's kept list + the number of kept boxes of the other
// segments that sort before it; block = 1024, grid = N
__global__ void __launch_bounds__(1024)
nms_seg_merge_kernel(const float* __restrict__ sbox, const float* __restrict__ sscore, const int* __restrict__ sperm,
                     const int* __restrict__ kept, const int* __restrict__ kept_count, int num_seg, int cap,
                     int post_topk, float* __restrict__ out_box, float* __restrict__ out_score,
                     int* __restrict__ out_cat, int* __restrict__ out_src, int* __restrict__ out_count) {
  const int img = blockIdx.x;
  int total = 0;
  for (int s = 0; s < num_seg; ++s) total += kept_count[img * num_seg + s];
  auto key_of = [&](int s, int k) -> unsigned long long {
    const size_t base = ((size_t)img * num_seg + s);
    const int j = kept[base * post_topk + k];
    return ((unsigned long long)fkey(sscore[base * cap + j]) << 32) |
           (unsigned long long)(0xFFFFFFFFu - (uint32_t)sperm[base * cap + j]);
  };
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int s = 0, k = e;
    while (k >= kept_count[img * num_seg + s]) { k -= kept_count[img * num_seg + s]; ++s; }
    const unsigned long long mine = key_of(s, k);
    int rank = k;
    for (int o = 0; o < num_seg; ++o) {
      if (o == s) continue;
      int lo = 0, hi = kept_count[img * num_seg + o];   // first index whose key is NOT greater than mine
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (key_of(o, mid) > mine) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < post_topk) {
      const size_t base = ((size_t)img * num_seg + s);
      const int j = kept[base * post_topk + k];
      const size_t dst = (size_t)img * post_topk + rank;
      reinterpret_cast<float4*>(out_box)[dst] = reinterpret_cast<const float4*>(sbox)[base * cap + j];
      out_score[dst] = sscore[base * cap + j];
      out_cat[dst] = s;
      out_src[dst] = sperm[base * cap + j];
    }
  }
  if (threadIdx.x == 0) out_count[img] = min(total, post_topk);
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

// ------------------------------------------------------------------------------------------------
// K5: ROI heads label_and_sample_proposals for one image per block (detectron2 roi_heads.py:StandardROIHeads):
// append GT to the proposals, IoU-match (thr, no low-quality), sample `num_samples` with `pos_fraction`
// foreground, emit [fg..., bg...] in ascending candidate order.  block = 1024, candidates <= 4096.
__global__ void __launch_bounds__(1024)
roi_sample_kernel(const float* __restrict__ prop_box, const int* __restrict__ prop_count, int prop_stride,
                  const float* __restrict__ gt_boxes, const int* __restrict__ gt_classes,
                  const int* __restrict__ gt_counts, int gmax, float iou_thr, int num_classes, int num_samples,
                  float pos_fraction, const uint32_t* __restrict__ seed_ptr, const uint32_t* __restrict__ salts, int append_gt,
                  float* __restrict__ out_box, int* __restrict__ out_batch, int* __restrict__ out_class,
                  float* __restrict__ out_gtbox, int* __restrict__ out_src, int* __restrict__ out_count,
                  int* __restrict__ stats) {
  __shared__ signed char s_lab[4096];   // 1 fg, 0 bg, -1 ignore
  __shared__ short s_match[4096];
  __shared__ uint32_t s_hist[260];
  __shared__ int s_scan[40];
  __shared__ int s_cnt[2];
  const int img = blockIdx.x;
  const uint32_t seed = *seed_ptr;
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
  const int wused = (cand_stride + 63) / 64;  // nvalid <= cand_stride: blocks beyond it would exit immediately
  nms_mask_kernel<<<dim3(wused, wused, n_images), 64, 0, stream>>>(sbox, scat, nvalid, cpad, iou_thresh, mask);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_nms_sorted(mask)");
  ALDI_CHECK_ARG(post_topk > 0 && post_topk <= 8192, "aldi_nms_sorted: post_topk must be in (0, 8192]");
  nms_scan_kernel<<<n_images, 256, (size_t)post_topk * sizeof(int), stream>>>(mask, sbox, sscore, scat, sperm, nvalid, cpad, post_topk, out_box,
                                                out_score, out_cat, out_src, out_count, nullptr);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_nms_sorted(scan)");
  return ALDI_OK;
}

static int seg_cap_for(const int* seg_len, int num_seg) {
  int mx = 1;
  for (int i = 0; i < num_seg; ++i) mx = seg_len[i] > mx ? seg_len[i] : mx;
  int cap = 64;
  while (cap < mx) cap <<= 1;
  return cap;
}

extern "C" size_t aldi_nms_segmented_workspace_bytes(int n_images, int num_seg, const int* seg_len, int post_topk) {
  const size_t cap = (size_t)seg_cap_for(seg_len, num_seg);
  const size_t per = cap * (16 + 4 + 4) + cap * (cap / 64) * 8 + (size_t)post_topk * 4 + 8;
  return per * (size_t)n_images * num_seg + 4096;
}

extern "C" int aldi_nms_segmented(const float* cand_box, const float* cand_score, const unsigned char* cand_valid,
                                  int n_images, int cand_stride, int num_seg, const int* seg_off, const int* seg_len,
                                  float iou_thresh, int post_topk, void* workspace, size_t workspace_bytes,
                                  float* out_box, float* out_score, int* out_cat, int* out_src, int* out_count,
                                  void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(cand_box && cand_score && seg_off && seg_len && workspace && out_box && out_score && out_cat && out_src &&
                     out_count, "aldi_nms_segmented: null pointer");
  ALDI_CHECK_ARG(n_images > 0 && num_seg >= 1 && num_seg <= 8, "aldi_nms_segmented: 1..8 segments per image");
  ALDI_CHECK_ARG(post_topk > 0 && post_topk <= 8192, "aldi_nms_segmented: post_topk must be in (0, 8192]");
  SegTable T;
  T.num_seg = num_seg;
  for (int i = 0; i < 8; ++i) { T.off[i] = 0; T.len[i] = 0; }
  for (int i = 0; i < num_seg; ++i) {
    ALDI_CHECK_ARG(seg_len[i] >= 0 && seg_len[i] <= 4096 && seg_off[i] >= 0 && seg_off[i] + seg_len[i] <= cand_stride,
                   "aldi_nms_segmented: segment %d (%d at %d) outside the candidate stride %d or longer than 4096", i,
                   seg_len[i], seg_off[i], cand_stride);
    T.off[i] = seg_off[i]; T.len[i] = seg_len[i];
  }
  ALDI_CHECK_ARG(workspace_bytes >= aldi_nms_segmented_workspace_bytes(n_images, num_seg, seg_len, post_topk),
                 "aldi_nms_segmented: workspace too small");
  const int cap = seg_cap_for(seg_len, num_seg);
  const size_t ns = (size_t)n_images * num_seg;
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  float* sbox = reinterpret_cast<float*>(ws);      ws += ns * cap * 16;
  float* sscore = reinterpret_cast<float*>(ws);    ws += ns * cap * 4;
  int* sperm = reinterpret_cast<int*>(ws);         ws += ns * cap * 4;
  int* kept = reinterpret_cast<int*>(ws);          ws += ((ns * post_topk * 4 + 255) / 256) * 256;
  int* nvalid = reinterpret_cast<int*>(ws);        ws += ((ns * 4 + 255) / 256) * 256;
  int* kept_count = reinterpret_cast<int*>(ws);    ws += ((ns * 4 + 255) / 256) * 256;
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws);
  nms_seg_compact_kernel<<<dim3(num_seg, n_images), 1024, 0, stream>>>(cand_box, cand_score, cand_valid, cand_stride, T, cap,
                                                                       sbox, sscore, sperm, nvalid);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_nms_segmented(compact)");
  const int wused = cap / 64;
  nms_mask_kernel<<<dim3(wused, wused, (unsigned)ns), 64, 0, stream>>>(sbox, nullptr, nvalid, cap, iou_thresh, mask);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_nms_segmented(mask)");
  nms_scan_kernel<<<(unsigned)ns, 256, (size_t)post_topk * sizeof(int), stream>>>(mask, sbox, sscore, nullptr, sperm, nvalid, cap,
                                                                                post_topk, nullptr, nullptr, nullptr, nullptr,
                                                                                kept_count, kept);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_nms_segmented(scan)");
  nms_seg_merge_kernel<<<n_images, 1024, 0, stream>>>(sbox, sscore, sperm, kept, kept_count, num_seg, cap, post_topk, out_box,
                                                      out_score, out_cat, out_src, out_count);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_nms_segmented(merge)");
  return ALDI_OK;
}

// multi-block subsample_labels (select_mb.cu)
size_t aldi_subsample_workspace_bytes(int n_images);
int aldi_subsample_labels(signed char* labels, int n_images, int n, int num_samples, float pos_fraction,
                          const unsigned int* d_seed, const unsigned int* salts, int* stats, void* workspace,
                          cudaStream_t stream);

extern "C" size_t aldi_rpn_label_workspace_bytes(int n_images, int gmax) {
  return (((size_t)n_images * gmax * sizeof(int) + 255) & ~size_t(255)) + aldi_subsample_workspace_bytes(n_images) + 256;
}

extern "C" int aldi_rpn_label_anchors(const aldi_rpn_levels* L, int n_images, const float* gt_boxes,
                                      const int* gt_counts, int gmax, float iou_lo, float iou_hi, int num_samples,
                                      float pos_fraction, const unsigned int* d_seed, const unsigned int* salts,
                                      void* workspace, size_t workspace_bytes, signed char* labels, int* matched,
                                      int* stats, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(L && gt_boxes && gt_counts && d_seed && salts && workspace && labels && matched,
                 "aldi_rpn_label_anchors: null pointer");
  ALDI_CHECK_ARG(gmax > 0 && gmax <= 2048, "aldi_rpn_label_anchors: gmax must be in (0, 2048]");
  ALDI_CHECK_ARG(workspace_bytes >= aldi_rpn_label_workspace_bytes(n_images, gmax),
                 "aldi_rpn_label_anchors: workspace too small");
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  int* gt_best_ws = reinterpret_cast<int*>(ws);
  void* sub_ws = ws + (((size_t)n_images * gmax * sizeof(int) + 255) & ~size_t(255));
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
  if (num_samples > 0)
    return aldi_subsample_labels(labels, n_images, total, num_samples, pos_fraction, d_seed, salts, stats, sub_ws, stream);
  return ALDI_OK;
}

extern "C" int aldi_roi_label_sample(const float* prop_box, const int* prop_count, int prop_stride, int n_images,
                                     const float* gt_boxes, const int* gt_classes, const int* gt_counts, int gmax,
                                     float iou_thresh, int num_classes, int num_samples, float pos_fraction,
                                     const unsigned int* d_seed, const unsigned int* salts, int append_gt, float* out_box,
                                     int* out_batch, int* out_class, float* out_gtbox, int* out_src, int* out_count,
                                     int* stats, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(prop_box && prop_count && gt_boxes && gt_classes && gt_counts && d_seed && salts && out_box && out_batch &&
                     out_class && out_gtbox && out_src && out_count, "aldi_roi_label_sample: null pointer");
  ALDI_CHECK_ARG(prop_stride + gmax <= 4096, "aldi_roi_label_sample: proposals + gt must be <= 4096 per image");
  roi_sample_kernel<<<n_images, 1024, 0, stream>>>(prop_box, prop_count, prop_stride, gt_boxes, gt_classes, gt_counts,
                                                   gmax, iou_thresh, num_classes, num_samples, pos_fraction, d_seed,
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
