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

/* ---- library bookkeeping -------------------------------------------------------------------- */
const char* aldi_last_error(void);
int aldi_abi_version(void);
/* number of kernels of THIS library launched since load / since the last reset */
unsigned long long aldi_launch_count(void);
void aldi_reset_launch_count(void);

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
} aldi_wgrad_params;
int aldi_wgrad_tc(const aldi_wgrad_params* p, void* stream);
int aldi_wgrad_f32(const aldi_wgrad_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ALDI_B200_H_ */
