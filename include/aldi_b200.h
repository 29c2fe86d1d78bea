/*
 * aldi_b200 C ABI — the drop-in boundary for the ALDI++ teacher–student train step on B200 (sm_100a).
 *
 * The reference (justinkay/aldi) has no C FFI: its hot path sits behind a Python class/registry surface
 * (aldi/trainer.py:28-136, aldi/distill.py:87-278, aldi/ema.py:8-60, aldi/pseudolabeler.py:7-73,
 * aldi/align.py:17-136) and delegates every arithmetic op to Detectron2 -> ATen/cuDNN/torchvision.
 * This header declares the kernels that replace those delegated ops.  Each entry cites the reference
 * call site (paths relative to the reference root) whose arithmetic it takes over.
 *
 * Conventions:
 *   - plain pointers + sizes; no torch types.  All pointers are DEVICE pointers unless named `h_*`.
 *   - caller owns every buffer, including workspaces;
 *   - every call enqueues on the `stream` argument (a cudaStream_t passed as void*) and returns
 *     immediately: 0 on success, <0 on error (aldi_last_error() gives the message);
 *   - no exceptions, no internal streams, no host synchronisation unless stated;
 *   - activations are channels-last (N,H,W,C), dtype code ALDI_F32 (0) or ALDI_BF16 (1);
 *   - weights are (Cout, kh, kw, Cin) ("OHWI").
 */
#ifndef ALDI_B200_H_
#define ALDI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALDI_DTYPE_F32 0
#define ALDI_DTYPE_BF16 1
#define ALDI_DTYPE_F64 2

/* ---- library bookkeeping -------------------------------------------------------------------- */
const char* aldi_last_error(void);
int aldi_abi_version(void);
/* number of kernels of THIS library launched since load / since the last reset */
unsigned long long aldi_launch_count(void);
void aldi_reset_launch_count(void);
/* programmatic dependent launch of the library's kernels: off by default (ALDI_PDL=1 in the environment turns it on --
 * measured neutral on the graph-replayed step); the data-parallel step keeps it off explicitly: with NCCL kernels
 * sharing the SMs, early-launched dependents only park on them. */
void aldi_set_pdl(int on);

/* ---- EMA teacher update: aldi/ema.py:32-57 (EMA._update_ema / _init_ema_weights) -------------
 * teacher[i] = student[i]*(1-alpha) + teacher[i]*alpha over one flat fp32 buffer holding EVERY
 * state_dict entry (parameters and FrozenBN buffers).  alpha==0 degenerates to the init copy
 * (iter <= start_iter).  Rounding matches the reference expression exactly (two fp32 products of the
 * fp32-rounded scalars (1-alpha) and alpha, then one fp32 add; no FMA contraction).              */
int aldi_ema_update(float* teacher, const float* student, size_t n, double alpha, void* stream);

/* ---- SGD-momentum step over a contiguous range of the flat buffers (D2 build_optimizer ->
 * torch.optim.SGD, reached from aldi/trainer.py:199-208 and stepped at aldi/dropin.py:121,177)
 *   g = grad*grad_scale + wd*p ; m = momentum*m + g ; p -= lr*m      (dampening 0, no nesterov)
 * `teacher` (nullable): fuse the EMA of the freshly updated weights into the same pass
 * (what aldi/ema.py would compute at the next before_step).                          */
int aldi_sgd_momentum_step(float* params, float* momentum_buf, const float* grads, size_t n, float lr,
                           float weight_decay, float momentum, float grad_scale, float* teacher,
                           double ema_alpha, void* stream);

/* ---- weight packing: fp32 master (cout, taps, cin) -> GEMM operand, zero padded to (cout_p, cin_p)
 *   dgrad == 0: out[co][t][ci]            = w[co][t][ci]                     (forward operand)
 *   dgrad == 1: out[ci][taps-1-t][co]     = w[co][t][ci] * scale[co]         (data-gradient operand:
 *               spatially flipped, in/out swapped, FrozenBN scale folded)                         */
int aldi_pack_weight(const float* w, const float* scale, void* out, int out_dtype, int dgrad, int cout, int taps,
                     int cin, int cout_p, int cin_p, void* stream);

/* ---- batched operand refresh: all packs / FrozenBN folds / bias copies of one model in ONE launch.
 * `d_descs` and `d_block_start` are DEVICE arrays built once by the host (block_start[i] = sum of
 * aldi_refresh_blocks(desc[j]) for j < i; total_blocks = the full sum).
 *   kind 0: forward operand pack (as aldi_pack_weight dgrad=0)
 *   kind 1: data-gradient operand pack, FrozenBN scale derived in place from bn_w / bn_var (NULL: no scale)
 *   kind 2: FrozenBN fold: out[i] = bn_w*rsqrt(bn_var+eps) (scale), out2[i] = bn_b - bn_mean*scale (shift), i < cout
 *   kind 3: out2[i] = w[i], i < cout (conv bias -> epilogue shift)
 *   kind 5: transposed forward operand out[(t*cin + ci)][co] = w[co][t][ci], zero padded to cout_p columns (the
 *           "scatter" operand of the sparse RPN backward: a 1x1 conv from cout channels to taps*cin outputs)
 *   kind 4: stem weights (cout,7,7,3) -> bf16 [cout_p][4][4][2][2][4] taps over the space-to-depth map
 *           of aldi_stem_s2d (row tap a, column tap b, r+1 = 2a+dy, s+1 = 2b+dx, zero where r/s = -1, c = 3) */
typedef struct {
  int kind, out_dtype;
  const float* w;
  const float* bn_w;
  const float* bn_b;
  const float* bn_mean;
  const float* bn_var;
  void* out;
  float* out2;
  int cout, taps, cin, cout_p, cin_p;
  float eps;
} aldi_refresh_desc;
int aldi_refresh_blocks(const aldi_refresh_desc* desc);
int aldi_refresh_operands(const aldi_refresh_desc* d_descs, const int* d_block_start, int n_desc, int total_blocks,
                          void* stream);

/* ---- implicit-GEMM convolution / linear on tcgen05 tensor cores (bf16 in, fp32 accumulate) ----
 * Replaces cuDNN/cuBLAS calls under detectron2 ResNet/FPN/RPN/box-head (reached from aldi/trainer.py:87,
 * aldi/distill.py:157,162, aldi/pseudolabeler.py:21).  One call = forward OR data-gradient of one layer:
 *   acc[n,h,w,co] = sum_{r,s,ci} x[n, h*stride+r-pad_h, w*stride+s-pad_w, ci] * w[co, r, s, ci]
 *   v = acc*scale[co] + bias[co] (+ residual) ; relu ; (* (mask>0)) ; (+= out)
 * `x` is an arbitrary strided channels-last VIEW (stride-2 1x1 convs pass x[:, ::2, ::2, :]).     */
typedef struct {
  const void* x;            /* bf16 view, channel stride 1 */
  int x_c, x_w, x_h, x_n;   /* view extents (x_c multiple of 64) */
  long long x_sw, x_sh, x_sn; /* element strides */
  const void* w;            /* bf16 [cout_p][taps_h*taps_w*x_c], cout_p multiple of 64 */
  int cout_p;
  int taps_h, taps_w, pad_h, pad_w;
  int stride;               /* input coord = out*stride + tap - pad (tensor-core path requires 1: pass a strided view) */
  int n, ho, wo;            /* output extents */
  const float* scale;       /* [cout_p] or NULL (=1) */
  const float* bias;        /* [cout_p] or NULL (=0) */
  const void* residual;     /* bf16 or NULL */
  int res_mode;             /* 0 none, 1 same geometry, 2 nearest-2x upsample of a (ho/2, wo/2) map */
  long long res_sw, res_sh, res_sn;
  const void* mask;         /* bf16 or NULL: v *= (mask > 0)  (ReLU backward of the producer) */
  long long mask_sw, mask_sh, mask_sn;
  void* out;
  int out_dtype;            /* ALDI_DTYPE_* */
  int cout_store;           /* channels actually written (<= cout_p) */
  long long out_sw, out_sh, out_sn;
  int relu;
  int accumulate;           /* out += v */
} aldi_conv_params;
int aldi_conv_tc(const aldi_conv_params* p, void* stream);
/* same contract, fp32 activations/weights on CUDA cores (parity mode; also checks the tcgen05 path) */
int aldi_conv_f32(const aldi_conv_params* p, void* stream);

/* ---- split-bf16 parity mode of the tensor-core path ("bf16x3" / "bf16x6", StepConfig.dtype) -----------------------
 * Holds the tcgen05 kernels themselves to the reference's fp32 arithmetic (the 1e-3 bar of the parity tests): every
 * fp32 operand is written as a sum of 2 or 3 bf16 tensors, aldi_conv_tc / aldi_wgrad_tc run once per kept product
 * term into ONE fp32 accumulation, and the fused epilogue is applied once on that sum.
 *   aldi_split_bf16: part[0] = bf16(x), part[k] = bf16(x - part[0] - .. - part[k-1]); x: strided channels-last fp32
 *       view (element strides sn/sh/sw, unit channel stride); parts: contiguous (n,h,w,c) bf16, `part_stride`
 *       elements apart in `out`.
 *   aldi_conv_epilogue_f32: the epilogue of aldi_conv_params (scale, bias, residual, relu, mask, accumulate) on the
 *       contiguous fp32 accumulation `raw` (n, ho, wo, cout_p); residual / mask / out are fp32 views here; x, w and
 *       the tap fields of `p` are ignored.                                                                          */
int aldi_split_bf16(const float* x, int n, int h, int w, int c, long long sn, long long sh, long long sw, void* out,
                    long long part_stride, int parts, void* stream);
int aldi_conv_epilogue_f32(const float* raw, const aldi_conv_params* p, void* stream);

/* ---- weight gradient: dw[co, r, s, ci] (+)= scale[co] * sum_{n,h,w} dy[n,h,w,co] * x[n,h+r-pad_h,w+s-pad_w,ci]
 * fp32 result accumulated atomically into the flat gradient buffer (split-K over pixels).        */
typedef struct {
  const void* x;
  int x_c, x_w, x_h, x_n;
  long long x_sw, x_sh, x_sn;
  const void* dy;           /* (n, ho, wo, dy_c) channels-last, dy_c multiple of 64 */
  int dy_c;
  long long dy_sw, dy_sh, dy_sn;
  int n, ho, wo;
  int taps_h, taps_w, pad_h, pad_w;
  int stride;               /* as in aldi_conv_params */
  const float* scale;       /* [dy_c] or NULL */
  float* dw;                /* fp32 [cout_store][taps][cin_store] */
  int cout_store, cin_store;
  float* dbias;             /* NULL, or fp32 [cout_store]: dbias[co] += sum_{n,h,w} dy[n,h,w,co] in the same pass
                             * (the bias gradient of the layer; aldi_wgrad_tc only, aldi_wgrad_f32 ignores it) */
} aldi_wgrad_params;
int aldi_wgrad_tc(const aldi_wgrad_params* p, void* stream);
int aldi_wgrad_f32(const aldi_wgrad_params* p, void* stream);

/* ---- glue kernels of the trunk (detectron2 GeneralizedRCNN.preprocess_image, BasicStem, FPN backward) ---- */
int aldi_preprocess(const uint8_t* images /*(N,3,hin,win) uint8*/, const int* sizes /*(N,2) valid h,w*/, float* out
                    /*(N,hp,wp,4) fp32*/, int n, int hin, int win, int hp, int wp, const float* h_mean,
                    const float* h_std, void* stream);
/* fused normalise + im2col of the stem's 7x7/2 conv: out bf16 (N,ho,wo,192), k=(r*7+s)*3+c, zero for k>=147 */
int aldi_stem_im2col(const uint8_t* images, const int* sizes, void* out_bf16, int n, int hin, int win, int ho, int wo,
                     const float* h_mean, const float* h_std, void* stream);
/* fused normalise + 2x2 space-to-depth of the uint8 canvas: out bf16 (N, hin/2, win/2 + 4, 16), channel
 * (dy*2+dx)*4 + c, two zero columns each side; the stem conv reads it through an overlapping-row view
 * (N, hin/2, win/2, 64) with element strides (.., (win/2+4)*16, 16, 1) as a 4x1-tap conv, pad_h 2 */
int aldi_stem_s2d(const uint8_t* images, const int* sizes, void* out_bf16, int n, int hin, int win,
                  const float* h_mean, const float* h_std, void* stream);
/* the same map in fp32 (split-bf16 parity mode: split afterwards with aldi_split_bf16) */
int aldi_stem_s2d_f32(const uint8_t* images, const int* sizes, float* out, int n, int hin, int win,
                      const float* h_mean, const float* h_std, void* stream);
int aldi_maxpool3x3s2(const void* in, void* out, int dtype, int n, int h, int w, int c, void* stream);
/* coarse[n,h,w,c] += sum of the 2x2 block of fine (backward of nearest-2x upsample + add in FPN top-down) */
int aldi_sum2x2_accum(const void* fine, void* coarse, int dtype, int n, int h, int w, int c, void* stream);
int aldi_add_f32(void* dst, int dtype, const float* src, size_t n, void* stream);
/* dst = (dtype) src: the fp32 RoIAlign gradient map becomes the activation-dtype gradient of p2..p5 (first writer) */
int aldi_cast_f32(void* dst, int dtype, const float* src, size_t n, void* stream);
/* out[ch] += scale * sum over images and rows of x[img*img_stride + row*row_stride + ch]  (bias gradients;
 * strides in elements, rows 16-byte aligned) */
int aldi_colsum(const void* x, int dtype, int n_img, long long rows, long long img_stride, long long row_stride, int c,
                float scale, float* out, void* stream);
/* FrozenBatchNorm2d -> per-channel (scale, shift): scale = w*rsqrt(var+eps), shift = b - mean*scale */
int aldi_frozenbn_fold(const float* weight, const float* bias, const float* mean, const float* var, float eps,
                       float* scale, float* shift, int n, void* stream);

/* ---- RoIAlign over FPN levels (detectron2 ROIPooler + torchvision roi_align, aligned=True, sampling_ratio=0) ---- */
typedef struct {
  const void* feat[4];      /* per level (N, feat_h, feat_w, channels) channels-last */
  float* dfeat[4];          /* backward: fp32 gradient accumulators, same geometry (zeroed by caller) */
  int feat_h[4], feat_w[4];
  float scale[4];           /* 1/stride */
  int num_levels, min_level;
  float canonical_box_size; /* 224 */
  int canonical_level;      /* 4 */
  int channels, pooled, dtype;
  const float* rois;        /* (M,4) xyxy, image pixels */
  const int* roi_batch;     /* (M) image index */
  const int* num_valid;     /* device scalar: rows >= *num_valid are padding (NULL: all valid) */
  int num_rois;             /* M */
  void* out;                /* forward: (M, pooled, pooled, channels) */
  const void* dout;         /* backward */
} aldi_roialign_params;
int aldi_roi_align_forward(const aldi_roialign_params* p, void* stream);
int aldi_roi_align_backward(const aldi_roialign_params* p, void* stream);

/* ---- selection ops (bit-exact index sets; compiled without FMA contraction) ------------------ */
typedef struct {
  int num_levels, num_anchors;
  int h[5], w[5], stride[5], loc_off[5]; /* loc_off: first location of the level in the concatenated map */
  int total_locs, ch_stride;             /* rpn_out is (N, total_locs, ch_stride) fp32: A logits then A*4 deltas */
  float cell[5][3][4];                   /* DefaultAnchorGenerator cell anchors */
  float scale_clamp, min_box_size;
} aldi_rpn_levels;
/* detectron2 find_top_rpn_proposals front half: per (image, level) top-k logits (descending), decode
 * (Box2BoxTransform weights 1), clip, finite/non-empty validity.  Candidates are level-major.        */
size_t aldi_rpn_topk_workspace_bytes(int n_images, int num_levels);
int aldi_rpn_topk_decode(const float* rpn_out, const aldi_rpn_levels* levels, int n_images, int pre_topk,
                         const int* img_sizes, float* cand_box, float* cand_score, int* cand_cat, int* cand_idx,
                         unsigned char* cand_valid, int cand_stride, int* err_flag, void* workspace,
                         size_t workspace_bytes, void* stream);
/* batched_nms (per-category greedy NMS, IoU > thresh suppresses) + keep[:post_topk]; output in score order */
size_t aldi_nms_workspace_bytes(int n_images, int cand_stride);
int aldi_nms_sorted(const float* cand_box, const float* cand_score, const int* cand_cat,
                    const unsigned char* cand_valid, const int* cand_count, int n_images, int cand_stride,
                    float iou_thresh, int post_topk, void* workspace, size_t workspace_bytes, float* out_box,
                    float* out_score, int* out_cat, int* out_src, int* out_count, void* stream);
/* Same result as aldi_nms_sorted for candidates that arrive as `num_seg` score-sorted segments per image whose
 * category is the segment index (the RPN: one segment per FPN level, straight from aldi_rpn_topk_decode): the
 * segments are compacted, masked and scanned in parallel and merged by rank.  seg_off / seg_len: HOST arrays. */
size_t aldi_nms_segmented_workspace_bytes(int n_images, int num_seg, const int* seg_len, int post_topk);
int aldi_nms_segmented(const float* cand_box, const float* cand_score, const unsigned char* cand_valid, int n_images,
                       int cand_stride, int num_seg, const int* seg_off, const int* seg_len, float iou_thresh,
                       int post_topk, void* workspace, size_t workspace_bytes, float* out_box, float* out_score,
                       int* out_cat, int* out_src, int* out_count, void* stream);
/* Sampling seed: `d_seed` is a DEVICE pointer to one uint32 (the value aldi/helpers.py:17-26 ManualSeed would
 * feed torch.manual_seed) so that a captured CUDA graph of the step can be replayed with a new seed.       */
/* RPN.label_and_sample_anchors: Matcher([lo,hi],[0,-1,1], low-quality) + subsample_labels; labels (N,R) int8 */
size_t aldi_rpn_label_workspace_bytes(int n_images, int gmax);
int aldi_rpn_label_anchors(const aldi_rpn_levels* levels, int n_images, const float* gt_boxes, const int* gt_counts,
                           int gmax, float iou_lo, float iou_hi, int num_samples, float pos_fraction,
                           const unsigned int* d_seed, const unsigned int* salts, void* workspace, size_t workspace_bytes,
                           signed char* labels, int* matched, int* stats, void* stream);
/* StandardROIHeads.label_and_sample_proposals: append GT, Matcher([thr],[0,1]), subsample; (N*num_samples) rows */
int aldi_roi_label_sample(const float* prop_box, const int* prop_count, int prop_stride, int n_images,
                          const float* gt_boxes, const int* gt_classes, const int* gt_counts, int gmax,
                          float iou_thresh, int num_classes, int num_samples, float pos_fraction,
                          const unsigned int* d_seed, const unsigned int* salts, int append_gt, float* out_box, int* out_batch, int* out_class,
                          float* out_gtbox, int* out_src, int* out_count, int* stats, void* stream);
/* FastRCNNOutputLayers.inference front half: softmax, per-class decode + clip, score filter */
int aldi_roi_inference_candidates(const float* pred, int pred_stride, const float* prop_box, const int* prop_count,
                                  int prop_stride, int n_images, int num_classes, const int* img_sizes,
                                  float score_thresh, const float* h_weights4, float scale_clamp, float* cand_box,
                                  float* cand_score, int* cand_cat, int* cand_src, int* cand_count, int cand_stride,
                                  void* stream);
/* aldi/pseudolabeler.py:51-67 process_bbox: keep detections with score > threshold (order preserved) */
int aldi_pseudo_label_threshold(const float* det_box, const float* det_score, const int* det_class,
                                const int* det_count, int det_stride, int n_images, float threshold, float* gt_box,
                                int* gt_class, float* gt_score, int* gt_count, int gmax, void* stream);

/* ---- fused losses (scalar values accumulated into loss_out[0..1], gradients written densely) ---- */
/* detectron2 RPN.losses: loss_out[0] += loss_rpn_cls, loss_out[1] += loss_rpn_loc (both * gscale) */
int aldi_rpn_loss(const float* rpn_out, const aldi_rpn_levels* levels, int n_images, const signed char* labels,
                  const int* matched, const float* gt_boxes, const int* gt_counts, int gmax,
                  int batch_size_per_image, float w_cls, float w_loc, float gscale, void* drpn, int dtype,
                  int dstride, int accumulate, float* loss_out, void* stream);
/* detectron2 FastRCNNOutputLayers.losses: loss_out[0] += loss_cls, loss_out[1] += loss_box_reg */
int aldi_roi_loss(const float* pred, int pred_stride, int m, int num_classes, const int* gt_class,
                  const float* roi_box, const float* gt_box, const int* counts, int n_images,
                  const float* h_weights4, float w_cls, float w_box, float gscale, void* dpred, int dtype,
                  int dstride, float* loss_out, void* stream);
/* aldi/distill.py:193-229: loss_out[0] += loss_obj_bce, loss_out[1] += loss_rpn_l1 */
int aldi_distill_rpn_loss(const float* student_rpn_out, const float* teacher_rpn_out, const aldi_rpn_levels* levels,
                          int n_images, const signed char* labels, const int* stats, float obj_temperature,
                          float w_obj, float w_reg, float gscale, void* drpn, int dtype, int dstride, int accumulate,
                          float* loss_out, void* stream);
/* ---- sparse backward of the RPN head (csrc/rpn_sparse.cu) ---------------------------------------------------------
 * The gradient RPN.losses (detectron2) and aldi/distill.py:193-229 send into the head outputs is non-zero only at the
 * anchors `subsample_labels` picked (256 per image): the non-zero rows of `drpn` (n_images * total_locs rows) are
 * compacted, the operands of the head's two layers gathered for those locations, the GEMMs run on the gathered rows
 * (aldi_conv_tc / aldi_wgrad_tc on a 1 x 1 x cap x C image) and the feature-gradient rows scattered back.
 *   compact: idx[cap] <- row indices with a non-zero among the first `channels` columns (-1 beyond *count)
 *   gather : dy_g[cap][64] = drpn rows, t_g[cap][C] = hidden rows, x_g[cap][9][C] = 3x3 neighbourhoods of the FPN
 *            feature (zeros outside the map = the conv's padding; all-zero rows beyond *count); sets bit 1 of
 *            *err_flag if *count > cap (rows would be lost)
 *   scatter: dfeat[img, y+r-1, x+s-1, :] += dx_g[k][r*3+s][:]   (fp32 atomics; the map RoIAlign backward adds into)
 * Per level: `feat` the layer input p_l (strided channels-last view: p6 is p5[:, ::2, ::2]), `hidden` the ReLU'd 3x3
 * output (contiguous), `dfeat` the fp32 gradient map of p_l (p6: the strided view of p5's).                          */
typedef struct {
  const aldi_rpn_levels* levels;
  int dtype, channels;                   /* activation dtype of feat / hidden / drpn / *_g; C (multiple of 8) */
  const void* feat[5];
  long long feat_sn[5], feat_sh[5], feat_sw[5];
  const void* hidden[5];
  float* dfeat[5];
  long long dfeat_sn[5], dfeat_sh[5], dfeat_sw[5];
  const void* drpn;
  int dstride;                           /* 64 */
  const int* idx;
  const int* count;
  int cap;
  void* dy_g;
  void* t_g;
  void* x_g;
  int* err_flag;                         /* nullable */
} aldi_rpn_sparse_params;
int aldi_rpn_sparse_compact(const void* drpn, int dtype, int n_images, int total_locs, int dstride, int channels, int cap,
                            int* idx, int* count, void* stream);
int aldi_rpn_sparse_gather(const aldi_rpn_sparse_params* p, void* stream);
int aldi_rpn_sparse_scatter(const aldi_rpn_sparse_params* p, const void* dx_g, void* stream);

/* aldi/distill.py:231-278: loss_out[0] += loss_cls_ce (CE or KL), loss_out[1] += loss_roih_l1 */
int aldi_distill_roi_loss(const float* student_pred, const float* teacher_pred, int pred_stride, int m,
                          int num_classes, const int* row_class, const int* counts, int n_images,
                          float cls_temperature, int use_kl, float w_cls, float w_reg, float gscale, void* dpred,
                          int dtype, int dstride, int accumulate, float* loss_out, void* stream);
/* aldi/align.py:81-90: loss_out[0] += weight * mean BCE-with-logits(pred, domain_label) */
int aldi_domain_bce_loss(const float* pred, int n, int stride, float domain_label, float weight, float gscale,
                         void* dpred, int dtype, int dstride, float* loss_out, void* stream);
/* aldi/align.py:81-90,103-136: the discriminator's last Linear(C,1) + BCE-with-logits (mean over valid rows) against
 * the constant domain label, fused with that layer's gradients: loss_out[0] += weight*gscale*mean BCE;
 * dw[C] += sum_r dl[r]*feat[r]; db[0] += sum_r dl[r]; dfeat[r] = dl[r]*w (gated by feat>0 when relu_mask) and
 * ndfeat = -dfeat (nullable; operand of the gradient-reversed data gradient, aldi/helpers.py:51-63).
 * counts (nullable): row r is valid iff (r % rows_per_image) < counts[r / rows_per_image]. */
int aldi_domain_head_loss(const void* feat, int dtype, int n, int c, long long feat_stride, const int* counts,
                          int rows_per_image, const float* w, const float* b, float domain_label, float weight,
                          float gscale, int relu_mask, void* dfeat, void* ndfeat, long long d_stride, float* dw,
                          float* db, float* loss_out, void* stream);
/* backward of ReLU -> AdaptiveAvgPool2d(1) (aldi/align.py:113): dh[n,p,ch] = h>0 ? dgap[n,ch]*scale : 0, ndh = -dh */
int aldi_gap_backward(const void* h, const float* dgap, int dtype, int n, long long pix, int c_p, int c, float scale,
                      void* dh, void* ndh, void* stream);

/* ---- multi-scale deformable attention (Deformable-DETR, BASELINE configs[3]) --------------------------------------
 * Replaces MSDA.ms_deform_attn_forward / ms_deform_attn_backward of the reference's CUDA extension
 * (aldi/detr/libs/DeformableDETRDetectron2/deformable_detr/models/ops/src/vision.cpp:13-16,
 * src/ms_deform_attn.h:21-62, called from functions/ms_deform_attn_func.py:24-36).  Same tensors, same layouts:
 * value (N,S,M,D); sampling_loc (N,Lq,M,L,P,2) as (x,y) in [0,1]; attn_weight (N,Lq,M,L,P); out / grad_out (N,Lq,M*D).
 * spatial_h/w and level_start are HOST arrays of L ints (the reference passes device int64 tensors and reads them in
 * every thread).  dtype F32 or F64 (what the reference dispatches).  grad_value must be zero-filled by the caller
 * (it is accumulated with atomics); grad_loc and grad_attn are fully overwritten.  The reference's im2col_step only
 * batches its launches and has no numerical effect; the host mirror keeps its `N % im2col_step == 0` check. */
typedef struct {
  const void* value;
  const void* sampling_loc;
  const void* attn_weight;
  const int* spatial_h;     /* host, L */
  const int* spatial_w;     /* host, L */
  const int* level_start;   /* host, L */
  int n, s, m, d, lq, l, p;
  int dtype;
  void* out;                /* forward */
  const void* grad_out;     /* backward */
  void* grad_value;
  void* grad_loc;
  void* grad_attn;
} aldi_msda_params;
int aldi_msda_forward(const aldi_msda_params* p, void* stream);
int aldi_msda_backward(const aldi_msda_params* p, void* stream);

/* ---- strong augmentation on the device ---------------------------------------------------------------
 * Replaces, per planar uint8 image (3, h, w), aldi/aug.py:39-60 build_strong_augmentation (+ MICTransform,
 * aldi/aug.py:154-176) that the reference runs on dataloader workers: colour jitter / grayscale / gaussian blur over
 * H, W and channels / up to three random-erase rectangles / MIC block mask, with the random PARAMETERS drawn by the
 * host in the reference's RNG order.  src and dst may alias only when no blur is requested. */
#define ALDI_AUG_MAX_RADIUS 8            /* int(4 * sigma + 0.5) for sigma <= 2.0 (aldi/aug.py:50) */
typedef struct {
  int h, w;
  long long src_plane, src_row, dst_plane, dst_row; /* strides in bytes */
  int do_color;                                      /* RandomApply(p=0.8) fired */
  double contrast_w, brightness_w, saturation_w;     /* the three np.random.uniform(0.6, 1.4) draws */
  int do_gray;                                       /* RandomApply(RandomSaturation(0,0), p=0.2) fired */
  int blur_radius;                                   /* -1: no blur; else int(4*sigma+0.5) */
  double blur_taps[2 * ALDI_AUG_MAX_RADIUS + 1];     /* scipy _gaussian_kernel1d(sigma, 0, radius), centre at [radius] */
  int num_erase;
  int erase_rect[3][4];                              /* h0, w0, h, w (aldi/aug.py:124-137) */
  unsigned int erase_seed[3];                        /* fill noise = hash(seed, pixel, channel) * 255 */
  const unsigned char* mic_mask;                     /* device (mic_h, mic_w) bytes, 1 = keep; NULL = no MIC */
  int mic_h, mic_w;
} aldi_aug_params;
size_t aldi_strong_augment_workspace_bytes(int h, int w);
int aldi_strong_augment(const unsigned char* src, unsigned char* dst, const aldi_aug_params* p, void* workspace,
                        size_t workspace_bytes, void* stream);

/* ---- ConvNeXt building blocks (BASELINE configs[4], aldi/backbone.py:189-346), channels-last, dtype F32 | BF16 -------
 * rows = pixels, `stride` = padded channel count of a row (multiple of 8), c = real channels.                      */
/* LayerNorm over channels per pixel (both data formats of aldi/backbone.py:321-346); stats[row] = (mean, rstd) */
int aldi_layernorm_forward(const void* x, const float* gamma, const float* beta, float eps, long long rows, int c,
                           int stride, int dtype, void* y, float* stats, void* stream);
int aldi_layernorm_backward(const void* x, const float* gamma, const float* stats, const void* dy, long long rows, int c,
                            int stride, int dtype, void* dx, int accumulate, float* dgamma, float* dbeta, void* stream);
/* depthwise 7x7, padding 3 (ConvNextBlock.dwconv); w fp32 [c][49]; flip=1 -> data gradient (reversed taps, bias NULL) */
int aldi_dwconv7(const void* x, const float* w, const float* bias, int n, int h, int wd, int c, int stride, int dtype,
                 int flip, void* y, int accumulate, void* stream);
int aldi_dwconv7_wgrad(const void* x, const void* dy, int n, int h, int wd, int c, int stride, int dtype, float* dw,
                       void* stream);
/* exact GELU: da == NULL -> out = gelu(h); else out = da * gelu'(h) */
int aldi_gelu(const void* h, const void* da, void* out, size_t n, int dtype, void* stream);
/* out = input + gamma * u * keep[image]  (layer scale + DropPath + residual, aldi/backbone.py:222-227); and its gradients */
int aldi_layerscale_forward(const void* u, const void* input, const float* gamma, const float* keep, long long rows,
                            long long rows_per_image, int c, int stride, int dtype, void* out, void* stream);
int aldi_layerscale_backward(const void* u, const void* dy, const float* gamma, const float* keep, long long rows,
                             long long rows_per_image, int c, int stride, int dtype, void* du, float* dgamma, void* stream);
/* k x k stride-k "patchify" convs (aldi/backbone.py:249-258) as 1x1 GEMMs: out[n,y,x,(dy*b+dx)*c+ch] = in[n,b*y+dy,b*x+dx,ch];
 * inverse=1 scatters a (coarse, b*b*c) gradient back onto the fine map */
int aldi_space_to_depth(const void* in, void* out, int n, int ho, int wo, int block, int c, int in_stride, int out_stride,
                        int dtype, int inverse, void* stream);
int aldi_patchify_image(const unsigned char* images, const int* sizes, void* out, int n, int hp, int wp, int block,
                        int out_stride, int dtype, const float* h_mean, const float* h_std, void* stream);
/* torch.optim.AdamW over flat fp32 buffers (aldi/trainer.py:205-206); step counts from 1 */
int aldi_adamw_step(float* params, float* exp_avg, float* exp_avg_sq, const float* grads, size_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

/* ---- ViTDet backbone (BASELINE configs[2]; aldi/backbone.py:21-64 -> detectron2 modeling/backbone/vit.py + utils.py) ------
 * Everything the plain ViT + SimpleFeaturePyramid needs beyond the GEMM / LayerNorm / GELU kernels above.               */
/* window_partition / window_unpartition (vit utils): x (n, h, w, C) <-> win (n * ceil(h/ws) * ceil(w/ws), ws, ws, C), `stride` =
 * elements per token row in both.  inverse = 0: x -> win, bottom / right padding tokens written as ZEROS (they are keys of
 * their window, detectron2 does not mask them); inverse = 1: win -> x (padding tokens dropped).  The gradient of one
 * direction is the other.                                                                                              */
int aldi_window_partition(const void* in, void* out, int n, int h, int w, int ws, int stride, int dtype, int inverse,
                          void* stream);
/* x[img, r, :] += pos[r, :] (absolute position embedding, broadcast over the batch); pos fp32 with the same row stride */
int aldi_add_rows_bcast(void* x, const float* pos, int n, long long rows_per_image, int stride, int dtype, void* stream);
/* out[r, :] += sum over images of dx[img, r, :]  (gradient of the broadcast add); out fp32 */
int aldi_sum_over_batch(const void* dx, int n, long long rows_per_image, int stride, int dtype, float* out, void* stream);
/* get_abs_pos: F.interpolate(mode="bicubic", align_corners=False) of a channels-last fp32 (sh, sw, c) table to (dh, dw, c)
 * rows of `dst_stride` floats (A = -0.75, border indices clamped, as ATen upsample_bicubic2d).  backward = 1: the transpose,
 * src[...] += taps * dst[...] (src is the pos_embed gradient).                                                          */
int aldi_bicubic_resize(float* src, int sh, int sw, float* dst, int dh, int dw, int c, int dst_stride, int backward,
                        void* stream);
/* get_rel_pos: F.interpolate(mode="linear") of a (rows_in, c) fp32 table to (rows_out, c); backward = 1: transpose into src */
int aldi_linear_resize_rows(float* src, int rows_in, float* dst, int rows_out, int c, int backward, void* stream);
/* nn.MaxPool2d(2, 2) of SimpleFeaturePyramid's scale-0.5 branch; backward routes dy to the first maximum of each 2x2 block */
int aldi_maxpool2x2(const void* x, void* out, int dtype, int n, int h, int w, int stride, void* stream);
int aldi_maxpool2x2_backward(const void* x, const void* dy, void* dx, int dtype, int n, int h, int w, int stride,
                             void* stream);
/* ReLU after a LayerNorm (FastRCNNConvFCHead with NORM "LN": conv -> LN -> ReLU, configs/Base-RCNN-VitDetB.yaml:7-12):
 * da == NULL -> out = max(x, 0); else out = da * (x > 0) */
int aldi_relu(const void* x, const void* da, void* out, size_t n, int dtype, void* stream);

/* Multi-head attention with MViTv2's decomposed relative position term (detectron2 vit.Attention + add_decomposed_rel_pos),
 * head dim 64, flash-style (the (tokens x tokens) matrix is never materialised):
 *   S[q, k] = scale * q.k + rel_h[kh, q] + rel_w[kw, q],   out = softmax_k(S) v        (k = (kh, kw) on the gh x gw token grid)
 * qkv: (batch, gh*gw tokens, ...) rows of `row_stride` elements = [q | k | v], each dim = heads*64 wide (the output of the
 * qkv Linear as it stands).  rel_h / rel_w (nullable, both or neither): fp32 (batch, heads, gh, tokens) / (batch, heads, gw,
 * tokens), KEY-major with the query index fastest -- a warp's 32 query rows read 128 contiguous bytes per key row / column:
 *   rel_h[kh, q] = q_vec . Rh[gh-1 + qh-kh],   rel_w[kw, q] = q_vec . Rw[gw-1 + qw-kw]        (q_vec UNSCALED)
 * They come from ONE plain GEMM of the q view with the concatenated tables [Rh (2gh-1 rows); Rw (2gw-1 rows)]
 * (aldi_conv_tc / aldi_conv_f32: fp32 (batch, tokens, heads, rp_stride) rows) followed by aldi_relpos_transpose, so the
 * tables' gradients and the term's part of dq are again plain GEMMs on the transposed-back gradient.
 * dtype BF16: tcgen05 kernels (QK^T, PV, and in the backward dO V^T, dS K, P^T dO, dS^T Q as 128 x N x 16 UMMA tiles with
 * TMEM accumulators, 16 x 8-token TMA patches as tiles); dtype F32 or impl = 1: CUDA-core fp32 kernels (parity mode / the
 * cross-check of the tensor-core path).
 * forward writes out (batch, tokens, heads*64) rows of out_stride and lse (batch, heads, tokens) fp32;
 * backward reads out / dout / lse and writes dqkv (same layout as qkv, all three thirds), drel_h / drel_w (every element)
 * and the delta workspace (batch, heads, tokens).                                                                        */
typedef struct {
  const void* qkv;
  int batch, gh, gw, heads;
  long long row_stride, batch_stride;     /* elements */
  const float* rel_h;
  const float* rel_w;
  float scale;
  int dtype;
  void* out;
  long long out_stride, out_batch_stride;
  float* lse;
  const void* dout;                       /* backward only from here */
  void* dqkv;
  float* drel_h;
  float* drel_w;
  float* delta;
  int impl;
} aldi_attn_params;
/* GEMM layout <-> key-major layout of the relative-position products.  backward = 0: rel (batch, tokens, heads, rp_stride)
 * -> rel_h, rel_w; backward = 1: the gradients drel_h, drel_w -> drel in the GEMM layout, EVERY column written (zeros where
 * a (query, table row) pair is not used and in the padding columns).                                                     */
int aldi_relpos_transpose(float* rel, int rp_stride, float* rel_h, float* rel_w, int batch, int gh, int gw, int heads, int backward,
                          void* stream);
int aldi_attention_forward(const aldi_attn_params* p, void* stream);
int aldi_attention_backward(const aldi_attn_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ALDI_B200_H_ */
