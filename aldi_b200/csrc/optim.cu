// Flat-buffer parameter kernels: EMA teacher update, SGD-momentum step, weight packing for the conv paths.
// All are pure HBM-bandwidth kernels: 128-bit accesses, grid sized to a multiple of the SM count.
#include "common.cuh"
#include "../../include/aldi_b200.h"

namespace {

// ---- EMA: aldi/ema.py:43-46  new_t = student * (1 - alpha) + teacher * alpha ------------------------
// torch evaluates this as three separate fp32 kernels (mul, mul, add), so no FMA contraction: use the
// explicitly rounded intrinsics to stay bit-identical with the reference expression.
__global__ void __launch_bounds__(256) ema_kernel(float* __restrict__ t, const float* __restrict__ s, size_t n4,
                                                  size_t n, float a, float oma) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float4* t4 = reinterpret_cast<float4*>(t);
  const float4* s4 = reinterpret_cast<const float4*>(s);
  for (size_t k = i; k < n4; k += stride) {
    float4 tv = t4[k];
    const float4 sv = __ldg(s4 + k);
    tv.x = __fadd_rn(__fmul_rn(sv.x, oma), __fmul_rn(tv.x, a));
    tv.y = __fadd_rn(__fmul_rn(sv.y, oma), __fmul_rn(tv.y, a));
    tv.z = __fadd_rn(__fmul_rn(sv.z, oma), __fmul_rn(tv.z, a));
    tv.w = __fadd_rn(__fmul_rn(sv.w, oma), __fmul_rn(tv.w, a));
    t4[k] = tv;
  }
  for (size_t k = n4 * 4 + i; k < n; k += stride) t[k] = __fadd_rn(__fmul_rn(s[k], oma), __fmul_rn(t[k], a));
}

// ---- SGD momentum (torch.optim.SGD, dampening 0, nesterov off) ------------------------------------
//   g = grad*grad_scale + wd*p ; m = mom*m + g ; p = p - lr*m ; optional fused EMA of the new p
__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, float* __restrict__ m, const float* __restrict__ g, size_t n4, size_t n, float lr,
           float wd, float mom, float gscale, float* __restrict__ teacher, float a, float oma) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* t4 = reinterpret_cast<float4*>(teacher);
  auto upd = [&](float& pv, float& mv, float gv) {
    gv = __fadd_rn(__fmul_rn(gv, gscale), __fmul_rn(wd, pv));
    mv = __fadd_rn(__fmul_rn(mv, mom), gv);
    pv = __fadd_rn(pv, __fmul_rn(-lr, mv));
  };
  for (size_t k = i; k < n4; k += stride) {
    float4 pv = p4[k], mv = m4[k];
    const float4 gv = __ldg(g4 + k);
    upd(pv.x, mv.x, gv.x); upd(pv.y, mv.y, gv.y); upd(pv.z, mv.z, gv.z); upd(pv.w, mv.w, gv.w);
    p4[k] = pv;
    m4[k] = mv;
    if (teacher) {
      float4 tv = t4[k];
      tv.x = __fadd_rn(__fmul_rn(pv.x, oma), __fmul_rn(tv.x, a));
      tv.y = __fadd_rn(__fmul_rn(pv.y, oma), __fmul_rn(tv.y, a));
      tv.z = __fadd_rn(__fmul_rn(pv.z, oma), __fmul_rn(tv.z, a));
      tv.w = __fadd_rn(__fmul_rn(pv.w, oma), __fmul_rn(tv.w, a));
      t4[k] = tv;
    }
  }
  for (size_t k = n4 * 4 + i; k < n; k += stride) {
    float pv = p[k], mv = m[k];
    upd(pv, mv, g[k]);
    p[k] = pv;
    m[k] = mv;
    if (teacher) teacher[k] = __fadd_rn(__fmul_rn(pv, oma), __fmul_rn(teacher[k], a));
  }
}

// ---- weight packing -------------------------------------------------------------------------------
// forward operand:  out[co][t][ci] (cout_p x taps x cin_p, zero padded)  = w[co][t][ci]
template <typename T>
__global__ void __launch_bounds__(256) pack_fwd_kernel(const float* __restrict__ w, T* __restrict__ out, int cout,
                                                       int taps, int cin, int cout_p, int cin_p) {
  const size_t total = (size_t)cout_p * taps * cin_p;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_p);
    size_t r = i / cin_p;
    const int t = (int)(r % taps);
    const int co = (int)(r / taps);
    float v = 0.f;
    if (co < cout && ci < cin) v = __ldg(w + ((size_t)co * taps + t) * cin + ci);
    out[i] = from_f32<T>(v);
  }
}
// data-gradient operand: out[ci][taps-1-t][co] (cin_p x taps x cout_p) = w[co][t][ci] * scale[co]
template <typename T>
__global__ void __launch_bounds__(256)
pack_dgrad_kernel(const float* __restrict__ w, const float* __restrict__ scale, T* __restrict__ out, int cout,
                  int taps, int cin, int cout_p, int cin_p) {
  const size_t total = (size_t)cin_p * taps * cout_p;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout_p);
    size_t r = i / cout_p;
    const int tf = (int)(r % taps);
    const int ci = (int)(r / taps);
    float v = 0.f;
    if (co < cout && ci < cin) {
      v = __ldg(w + ((size_t)co * taps + (taps - 1 - tf)) * cin + ci);
      if (scale) v *= __ldg(scale + co);
    }
    out[i] = from_f32<T>(v);
  }
}

// ---- batched operand refresh: every pack / FrozenBN fold / bias copy of a model in ONE launch ---------------
// A block handles kRefreshChunk consecutive output elements of one descriptor; block_start[] (prefix sums of the
// per-descriptor block counts) maps blockIdx -> descriptor by binary search.
constexpr int kRefreshChunk = 2048;

__device__ __forceinline__ float bn_scale(const aldi_refresh_desc& d, int co) {
  return __fmul_rn(__ldg(d.bn_w + co), __frsqrt_rn(__fadd_rn(__ldg(d.bn_var + co), d.eps)));
}

// kind 0, scalar fallback (cin_p != cin) and kind 1 in fp32: one output element per thread step
template <typename T>
__device__ void refresh_pack(const aldi_refresh_desc& d, size_t begin, size_t end) {
  T* out = reinterpret_cast<T*>(d.out);
  for (size_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    float v = 0.f;
    if (d.kind == 0) {  // out[co][t][ci]
      const int ci = (int)(i % d.cin_p);
      size_t r = i / d.cin_p;
      const int t = (int)(r % d.taps);
      const int co = (int)(r / d.taps);
      if (co < d.cout && ci < d.cin) v = __ldg(d.w + ((size_t)co * d.taps + t) * d.cin + ci);
    } else {            // out[ci][taps-1-t][co] = w[co][t][ci] * scale[co]
      const int co = (int)(i % d.cout_p);
      size_t r = i / d.cout_p;
      const int tf = (int)(r % d.taps);
      const int ci = (int)(r / d.taps);
      if (co < d.cout && ci < d.cin) {
        v = __ldg(d.w + ((size_t)co * d.taps + (d.taps - 1 - tf)) * d.cin + ci);
        if (d.bn_w) v *= bn_scale(d, co);
      }
    }
    out[i] = from_f32<T>(v);
  }
}

// kind 0 with cin_p == cin: the operand is the master weight itself (rows >= cout are zero) -> straight convert
__device__ void refresh_pack_fwd_bf16(const aldi_refresh_desc& d, size_t begin, size_t end) {
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out);
  const size_t valid = (size_t)d.cout * d.taps * d.cin;
  for (size_t i = begin + (size_t)threadIdx.x * 8; i < end; i += (size_t)blockDim.x * 8) {
    float f[8];
    if (i + 8 <= valid) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(d.w + i));
      const float4 b = __ldg(reinterpret_cast<const float4*>(d.w + i + 4));
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = (i + k < valid) ? __ldg(d.w + i + k) : 0.f;
    }
    uint4 q;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int k = 0; k < 4; ++k) h2[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
    *reinterpret_cast<uint4*>(out + i) = q;
  }
}

// kind 1 in bf16: a block transposes one 64 (cout) x 32 (cin) tile of one tap through shared memory, so both the
// fp32 reads (32 consecutive cin of a cout row) and the bf16 writes (64 consecutive cout of a cin row) are coalesced
__device__ void refresh_pack_dgrad_bf16(const aldi_refresh_desc& d, int blk) {
  __shared__ float tile[64][33];
  const int nco = d.cout_p / 64, nci = d.cin_p / 32;
  const int cob = blk % nco;
  int r = blk / nco;
  const int cib = r % nci;
  const int t = r / nci;
  const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int co = cob * 64 + p * 8 + ry, ci = cib * 32 + lane;
    float v = 0.f;
    if (co < d.cout && ci < d.cin) {
      v = __ldg(d.w + ((size_t)co * d.taps + t) * d.cin + ci);
      if (d.bn_w) v *= bn_scale(d, co);
    }
    tile[p * 8 + ry][lane] = v;
  }
  __syncthreads();
  const int row = threadIdx.x >> 3, cg = threadIdx.x & 7;  // 32 cin rows x 8 groups of 8 cout
  uint4 q;
  __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int k = 0; k < 4; ++k) h2[k] = __floats2bfloat162_rn(tile[cg * 8 + 2 * k][row], tile[cg * 8 + 2 * k + 1][row]);
  const size_t ci = (size_t)cib * 32 + row;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out);
  *reinterpret_cast<uint4*>(out + (ci * d.taps + (d.taps - 1 - t)) * d.cout_p + cob * 64 + cg * 8) = q;
}

__device__ __forceinline__ bool refresh_fast_fwd(const aldi_refresh_desc& d) {
  return d.kind == 0 && d.out_dtype == ALDI_DTYPE_BF16 && d.cin_p == d.cin && (d.cin % 8) == 0;
}
__device__ __forceinline__ bool refresh_fast_dgrad(const aldi_refresh_desc& d) {
  return d.kind == 1 && d.out_dtype == ALDI_DTYPE_BF16;
}

__global__ void __launch_bounds__(256)
refresh_kernel(const aldi_refresh_desc* __restrict__ descs, const int* __restrict__ block_start, int n_desc) {
  // binary search: last descriptor whose first block <= blockIdx.x
  int lo = 0, hi = n_desc - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (block_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const aldi_refresh_desc d = descs[lo];
  const int blk = blockIdx.x - block_start[lo];
  if (refresh_fast_dgrad(d)) {
    refresh_pack_dgrad_bf16(d, blk);
    return;
  }
  const size_t begin = (size_t)blk * kRefreshChunk;
  if (d.kind <= 1) {
    const size_t total = (d.kind == 0 ? (size_t)d.cout_p * d.taps * d.cin_p : (size_t)d.cin_p * d.taps * d.cout_p);
    const size_t end = begin + kRefreshChunk < total ? begin + kRefreshChunk : total;
    if (refresh_fast_fwd(d)) refresh_pack_fwd_bf16(d, begin, end);
    else if (d.out_dtype == ALDI_DTYPE_BF16) refresh_pack<__nv_bfloat16>(d, begin, end);
    else refresh_pack<float>(d, begin, end);
  } else if (d.kind == 5) {  // transposed forward operand: out[(t * cin + ci)][co], cout_p columns
    const size_t total = (size_t)d.taps * d.cin * d.cout_p;
    for (size_t i = begin + threadIdx.x; i < begin + kRefreshChunk && i < total; i += blockDim.x) {
      const int co = (int)(i % d.cout_p);
      const size_t k = i / d.cout_p;             // t * cin + ci
      const float v = co < d.cout ? __ldg(d.w + (size_t)co * d.taps * d.cin + k) : 0.f;
      if (d.out_dtype == ALDI_DTYPE_BF16) reinterpret_cast<__nv_bfloat16*>(d.out)[i] = __float2bfloat16_rn(v);
      else reinterpret_cast<float*>(d.out)[i] = v;
    }
  } else if (d.kind == 4) {  // stem 7x7x3 -> 4x4 taps over the 2x2 space-to-depth map: out[co][a][b][dy][dx][c4], bf16
    const size_t total = (size_t)d.cout_p * 256;
    for (size_t i = begin + threadIdx.x; i < begin + kRefreshChunk && i < total; i += blockDim.x) {
      const int co = (int)(i >> 8), rem = (int)(i & 255);
      const int a = rem >> 6, b = (rem >> 4) & 3, dy = (rem >> 3) & 1, dx = (rem >> 2) & 1, c = rem & 3;
      const int r = 2 * a + dy - 1, s = 2 * b + dx - 1;
      float v = 0.f;
      if (co < d.cout && r >= 0 && s >= 0 && c < 3) v = __ldg(d.w + (((size_t)co * 7 + r) * 7 + s) * 3 + c);
      if (d.out_dtype == ALDI_DTYPE_BF16) reinterpret_cast<__nv_bfloat16*>(d.out)[i] = __float2bfloat16_rn(v);
      else reinterpret_cast<float*>(d.out)[i] = v;   // split-bf16 parity mode: the operand is split afterwards
    }
  } else if (d.kind == 2) {  // FrozenBN fold: scale -> out, shift -> out2
    for (size_t i = begin + threadIdx.x; i < begin + kRefreshChunk && i < (size_t)d.cout; i += blockDim.x) {
      const float sc = bn_scale(d, (int)i);
      reinterpret_cast<float*>(d.out)[i] = sc;
      d.out2[i] = __fsub_rn(__ldg(d.bn_b + i), __fmul_rn(__ldg(d.bn_mean + i), sc));
    }
  } else {                   // plain bias -> shift
    for (size_t i = begin + threadIdx.x; i < begin + kRefreshChunk && i < (size_t)d.cout; i += blockDim.x)
      d.out2[i] = __ldg(d.w + i);
  }
}

int grid_for(size_t work_items, int threads) {
  size_t blocks = (work_items + threads - 1) / threads;
  size_t cap = (size_t)aldi_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" int aldi_ema_update(float* teacher, const float* student, size_t n, double alpha, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(teacher && student, "aldi_ema_update: null pointer");
  if (n == 0) return ALDI_OK;
  const bool aligned = ((reinterpret_cast<uintptr_t>(teacher) | reinterpret_cast<uintptr_t>(student)) & 15) == 0;
  const size_t n4 = aligned ? n / 4 : 0;
  ema_kernel<<<grid_for(n4 ? n4 : n, 256), 256, 0, stream>>>(teacher, student, n4, n, (float)alpha,
                                                              (float)(1.0 - alpha));
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_ema_update");
  return ALDI_OK;
}

extern "C" int aldi_sgd_momentum_step(float* params, float* momentum_buf, const float* grads, size_t n, float lr,
                                      float weight_decay, float momentum, float grad_scale, float* teacher,
                                      double ema_alpha, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(params && momentum_buf && grads, "aldi_sgd_momentum_step: null pointer");
  if (n == 0) return ALDI_OK;
  uintptr_t al = reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(momentum_buf) |
                 reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(teacher);
  const size_t n4 = (al & 15) == 0 ? n / 4 : 0;
  sgd_kernel<<<grid_for(n4 ? n4 : n, 256), 256, 0, stream>>>(params, momentum_buf, grads, n4, n, lr, weight_decay,
                                                              momentum, grad_scale, teacher, (float)ema_alpha,
                                                              (float)(1.0 - ema_alpha));
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_sgd_momentum_step");
  return ALDI_OK;
}

extern "C" int aldi_refresh_blocks(const aldi_refresh_desc* d) {
  size_t total;
  if (d->kind == 1 && d->out_dtype == ALDI_DTYPE_BF16)  // one block per (tap, 32-cin, 64-cout) transpose tile
    return d->taps * (d->cin_p / 32) * (d->cout_p / 64);
  if (d->kind == 0) total = (size_t)d->cout_p * d->taps * d->cin_p;
  else if (d->kind == 1) total = (size_t)d->cin_p * d->taps * d->cout_p;
  else if (d->kind == 4) total = (size_t)d->cout_p * 256;
  else if (d->kind == 5) total = (size_t)d->taps * d->cin * d->cout_p;
  else total = (size_t)d->cout;
  return (int)((total + kRefreshChunk - 1) / kRefreshChunk);
}

extern "C" int aldi_refresh_operands(const aldi_refresh_desc* d_descs, const int* d_block_start, int n_desc,
                                     int total_blocks, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(d_descs && d_block_start && n_desc > 0 && total_blocks > 0, "aldi_refresh_operands: bad args");
  refresh_kernel<<<total_blocks, 256, 0, stream>>>(d_descs, d_block_start, n_desc);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_refresh_operands");
  return ALDI_OK;
}

extern "C" int aldi_pack_weight(const float* w, const float* scale, void* out, int out_dtype, int dgrad, int cout,
                                int taps, int cin, int cout_p, int cin_p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(w && out, "aldi_pack_weight: null pointer");
  ALDI_CHECK_ARG(cout > 0 && taps > 0 && cin > 0 && cout_p >= cout && cin_p >= cin, "aldi_pack_weight: bad dims");
  const size_t total = (size_t)cout_p * taps * cin_p;
  const int grid = grid_for(total, 256);
  if (out_dtype == ALDI_DTYPE_BF16) {
    if (dgrad)
      pack_dgrad_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(w, scale, (__nv_bfloat16*)out, cout, taps, cin,
                                                                  cout_p, cin_p);
    else
      pack_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(w, (__nv_bfloat16*)out, cout, taps, cin, cout_p, cin_p);
  } else {
    if (dgrad)
      pack_dgrad_kernel<float><<<grid, 256, 0, stream>>>(w, scale, (float*)out, cout, taps, cin, cout_p, cin_p);
    else
      pack_fwd_kernel<float><<<grid, 256, 0, stream>>>(w, (float*)out, cout, taps, cin, cout_p, cin_p);
  }
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_pack_weight");
  return ALDI_OK;
}
