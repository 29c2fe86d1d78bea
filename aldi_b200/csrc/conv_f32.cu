// fp32 CUDA-core implicit-GEMM convolution, data-gradient (via transformed weights) and weight-gradient.
// This is the PARITY-MODE arithmetic path: fp32 in, fp32 FMA accumulate, fp32 out, so that whole-step
// results can be compared with the CPU oracle at ~1e-5 and selection ops (top-k / NMS / matcher) see the
// same values.  It honours exactly the same aldi_conv_params / aldi_wgrad_params contract as the
// tcgen05 kernels (and additionally a conv stride and arbitrary channel counts), so it doubles as the
// on-device checker of the tensor-core path at sizes the CPU oracle cannot reach.
#include "common.cuh"
#include "../../include/aldi_b200.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

struct F32ConvArgs {
  const float* x;
  int x_c, x_w, x_h, x_n;
  long long x_sw, x_sh, x_sn;
  const float* w;
  int cout_p, taps_w, ktot, pad_h, pad_w, stride;
  int n, ho, wo;
  long long npix;
  const float* scale;
  const float* bias;
  const float* residual;
  int res_mode;
  long long res_sw, res_sh, res_sn;
  const float* mask;
  long long mask_sw, mask_sh, mask_sn;
  float* out;
  int cout_store;
  long long out_sw, out_sh, out_sn;
  int relu, accumulate;
};

__global__ void __launch_bounds__(256) conv_f32_kernel(const F32ConvArgs a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long p0 = (long long)blockIdx.x * BM;
  const int c0 = blockIdx.y * BN;

  // A-load assignment: thread loads pixel (tid>>2), 4 consecutive k at (tid&3)*4
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;
  const long long ap = p0 + a_row;
  const bool a_valid = ap < a.npix;
  int an = 0, ah = 0, aw = 0;
  if (a_valid) {
    long long t = ap;
    aw = (int)(t % a.wo); t /= a.wo;
    ah = (int)(t % a.ho);
    an = (int)(t / a.ho);
  }
  // B-load assignment: thread loads cout (tid>>2), 4 consecutive k
  const int b_row = tid >> 2, b_k = (tid & 3) * 4;
  const int bc = c0 + b_row;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.ktot; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + a_k + e;
      float v = 0.f;
      if (a_valid && k < a.ktot) {
        const int tap = k / a.x_c, ci = k - tap * a.x_c;
        const int r = tap / a.taps_w, s = tap - r * a.taps_w;
        const int hi = ah * a.stride + r - a.pad_h, wi = aw * a.stride + s - a.pad_w;
        if (hi >= 0 && hi < a.x_h && wi >= 0 && wi < a.x_w)
          v = __ldg(a.x + (long long)an * a.x_sn + (long long)hi * a.x_sh + (long long)wi * a.x_sw + ci);
      }
      As[a_k + e][a_row] = v;
      const int kb = k0 + b_k + e;
      float wv = 0.f;
      if (bc < a.cout_p && kb < a.ktot) wv = __ldg(a.w + (long long)bc * a.ktot + kb);
      Bs[b_k + e][b_row] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long p = p0 + ty * 4 + i;
    if (p >= a.npix) continue;
    long long t = p;
    const int w = (int)(t % a.wo); t /= a.wo;
    const int h = (int)(t % a.ho);
    const int n = (int)(t / a.ho);
    const long long oo = (long long)n * a.out_sn + (long long)h * a.out_sh + (long long)w * a.out_sw;
    long long ro = 0, mo = 0;
    if (a.res_mode == 1) ro = (long long)n * a.res_sn + (long long)h * a.res_sh + (long long)w * a.res_sw;
    else if (a.res_mode == 2) ro = (long long)n * a.res_sn + (long long)(h >> 1) * a.res_sh + (long long)(w >> 1) * a.res_sw;
    if (a.mask) mo = (long long)n * a.mask_sn + (long long)h * a.mask_sh + (long long)w * a.mask_sw;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c >= a.cout_store) continue;
      float v = acc[i][j];
      if (a.scale) v *= a.scale[c];
      if (a.bias) v += a.bias[c];
      if (a.res_mode) v += a.residual[ro + c];
      if (a.relu) v = fmaxf(v, 0.f);
      if (a.mask) v = (a.mask[mo + c] > 0.f) ? v : 0.f;
      if (a.accumulate) v += a.out[oo + c];
      a.out[oo + c] = v;
    }
  }
}

struct F32WgradArgs {
  const float* x;
  int x_c, x_w, x_h;
  long long x_sw, x_sh, x_sn;
  const float* dy;
  long long dy_sw, dy_sh, dy_sn;
  int n, ho, wo;
  long long npix;
  int taps_w, pad_h, pad_w, stride, taps;
  const float* scale;
  float* dw;
  int cout_store, cin_store, ncols;  // ncols = taps*cin_store
  long long pix_per_split;
};

// GEMM: M = cout, N = (tap, cin), K = pixels (split over blockIdx.z)
__global__ void __launch_bounds__(256) wgrad_f32_kernel(const F32WgradArgs a) {
  __shared__ float As[BK][BM + 4];  // [pixel][cout]
  __shared__ float Bs[BK][BN + 4];  // [pixel][col]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const long long pk0 = (long long)blockIdx.z * a.pix_per_split;
  long long pk1 = pk0 + a.pix_per_split;
  if (pk1 > a.npix) pk1 = a.npix;

  // load assignment: thread loads pixel (tid>>4) , 4 consecutive m (or n) at (tid&15)*4
  const int l_k = tid >> 4, l_m = (tid & 15) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long k0 = pk0; k0 < pk1; k0 += BK) {
    const long long p = k0 + l_k;
    const bool pv = p < pk1;
    int pn = 0, ph = 0, pw = 0;
    if (pv) {
      long long t = p;
      pw = (int)(t % a.wo); t /= a.wo;
      ph = (int)(t % a.ho);
      pn = (int)(t / a.ho);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int co = m0 + l_m + e;
      float v = 0.f;
      if (pv && co < a.cout_store)
        v = __ldg(a.dy + (long long)pn * a.dy_sn + (long long)ph * a.dy_sh + (long long)pw * a.dy_sw + co);
      As[l_k][l_m + e] = v;
      const int col = n0 + l_m + e;
      float xv = 0.f;
      if (pv && col < a.ncols) {
        const int tap = col / a.cin_store, ci = col - tap * a.cin_store;
        const int r = tap / a.taps_w, s = tap - r * a.taps_w;
        const int hi = ph * a.stride + r - a.pad_h, wi = pw * a.stride + s - a.pad_w;
        if (hi >= 0 && hi < a.x_h && wi >= 0 && wi < a.x_w)
          xv = __ldg(a.x + (long long)pn * a.x_sn + (long long)hi * a.x_sh + (long long)wi * a.x_sw + ci);
      }
      Bs[l_k][l_m + e] = xv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = m0 + ty * 4 + i;
    if (co >= a.cout_store) continue;
    const float sc = a.scale ? a.scale[co] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= a.ncols) continue;
      atomicAdd(a.dw + (long long)co * a.ncols + col, acc[i][j] * sc);
    }
  }
}

}  // namespace

extern "C" int aldi_conv_f32(const aldi_conv_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(p && p->x && p->w && p->out, "aldi_conv_f32: null pointer");
  ALDI_CHECK_ARG(p->out_dtype == ALDI_DTYPE_F32, "aldi_conv_f32: output must be fp32");
  ALDI_CHECK_ARG(p->n > 0 && p->ho > 0 && p->wo > 0 && p->x_c > 0 && p->cout_p > 0, "aldi_conv_f32: empty");
  ALDI_CHECK_ARG(p->cout_store > 0 && p->cout_store <= p->cout_p, "aldi_conv_f32: bad cout_store");
  F32ConvArgs a;
  a.x = (const float*)p->x;
  a.x_c = p->x_c; a.x_w = p->x_w; a.x_h = p->x_h; a.x_n = p->x_n;
  a.x_sw = p->x_sw; a.x_sh = p->x_sh; a.x_sn = p->x_sn;
  a.w = (const float*)p->w;
  a.cout_p = p->cout_p;
  a.taps_w = p->taps_w;
  a.ktot = p->taps_h * p->taps_w * p->x_c;
  a.pad_h = p->pad_h; a.pad_w = p->pad_w;
  a.stride = p->stride > 0 ? p->stride : 1;
  a.n = p->n; a.ho = p->ho; a.wo = p->wo;
  a.npix = (long long)p->n * p->ho * p->wo;
  a.scale = p->scale; a.bias = p->bias;
  a.residual = (const float*)p->residual;
  a.res_mode = p->res_mode;
  a.res_sw = p->res_sw; a.res_sh = p->res_sh; a.res_sn = p->res_sn;
  a.mask = (const float*)p->mask;
  a.mask_sw = p->mask_sw; a.mask_sh = p->mask_sh; a.mask_sn = p->mask_sn;
  a.out = (float*)p->out;
  a.cout_store = p->cout_store;
  a.out_sw = p->out_sw; a.out_sh = p->out_sh; a.out_sn = p->out_sn;
  a.relu = p->relu; a.accumulate = p->accumulate;
  dim3 grid(aldi_div_up(a.npix, BM), aldi_div_up(p->cout_store, BN));
  conv_f32_kernel<<<grid, 256, 0, stream>>>(a);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_conv_f32");
  return ALDI_OK;
}

extern "C" int aldi_wgrad_f32(const aldi_wgrad_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(p && p->x && p->dy && p->dw, "aldi_wgrad_f32: null pointer");
  ALDI_CHECK_ARG(p->n > 0 && p->ho > 0 && p->wo > 0, "aldi_wgrad_f32: empty");
  F32WgradArgs a;
  a.x = (const float*)p->x;
  a.x_c = p->x_c; a.x_w = p->x_w; a.x_h = p->x_h;
  a.x_sw = p->x_sw; a.x_sh = p->x_sh; a.x_sn = p->x_sn;
  a.dy = (const float*)p->dy;
  a.dy_sw = p->dy_sw; a.dy_sh = p->dy_sh; a.dy_sn = p->dy_sn;
  a.n = p->n; a.ho = p->ho; a.wo = p->wo;
  a.npix = (long long)p->n * p->ho * p->wo;
  a.taps_w = p->taps_w;
  a.taps = p->taps_h * p->taps_w;
  a.pad_h = p->pad_h; a.pad_w = p->pad_w;
  a.stride = p->stride > 0 ? p->stride : 1;
  a.scale = p->scale;
  a.dw = p->dw;
  a.cout_store = p->cout_store;
  a.cin_store = p->cin_store;
  a.ncols = a.taps * p->cin_store;
  int gx = aldi_div_up(p->cout_store, BM), gy = aldi_div_up(a.ncols, BN);
  long long want = (4LL * aldi_num_sms() + gx * gy - 1) / ((long long)gx * gy);
  long long max_split = (a.npix + 255) / 256;
  long long split = want < 1 ? 1 : (want > max_split ? max_split : want);
  a.pix_per_split = ((a.npix + split - 1) / split + BK - 1) / BK * BK;
  int gz = aldi_div_up(a.npix, a.pix_per_split);
  wgrad_f32_kernel<<<dim3(gx, gy, gz), 256, 0, stream>>>(a);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_wgrad_f32");
  return ALDI_OK;
}
