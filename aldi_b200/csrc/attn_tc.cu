// Multi-head attention with decomposed relative-position bias on tcgen05 (ViTDet blocks, BASELINE configs[2];
// replaces detectron2 vit.Attention.forward + add_decomposed_rel_pos and their autograd, aldi/backbone.py:21-35).
//
// Head dim 64, bf16 operands, fp32 accumulation in TMEM.  A tile is a 16 x 8 patch of the token grid = 128 tokens: ONE
// 4-D TMA box {64 ch, 16, 8, 1} of the qkv tensor as the qkv Linear wrote it (channel coordinate selects q | k | v and the
// head) lands as 128 rows x 128 B in the 128-byte-swizzled K-major UMMA layout; patches that hang over the grid are
// zero-filled by TMA and their keys masked.  With 2-D key patches the relative-position term of a (query, key tile) is
// 8 + 16 numbers per query (one per key row / key column of the patch) instead of one per key.
//
//   forward  (CTA = query tile, loop over key tiles):   S = Q K^T (128x128x64) -> online softmax in registers, one
//            thread per query row (TMEM lane), P -> bf16 in swizzled shared memory -> O_tile = P V (128x64x128, V as the
//            MN-major B operand) -> rescale-and-add in registers.
//   backward dq  (CTA = query tile, loop over key tiles): S, dP = dO V^T, dS = P (dP - delta) -> dQ += dS K in TMEM;
//            row / column sums of dS are the gradient of the relative-position products (drelpos).
//   backward dkv (CTA = key tile, loop over query tiles): S, dP as above (thread = query row), P and dS staged ONCE and
//            read through MN-major descriptors as P^T / dS^T: dV += P^T dO, dK += dS^T Q in TMEM.
#include "common.cuh"
#include "sm100.cuh"
#include "tmap.h"
#include "../../include/aldi_b200.h"

using namespace sm100;

namespace {

constexpr int kPW = 16, kPH = 8;          // token patch of a tile
constexpr int kTile = kPW * kPH;          // 128 tokens = UMMA M
constexpr int kTileBytes = kTile * 128;   // 128 rows x 64 bf16
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct TcArgs {
  int gh, gw, heads, dim, tiles_h, tiles_w;
  float scale, scale_log2;
  const float* rel_h;      // key-major relative-position products: [(b, h, kh), q] / [(b, h, kw), q], query index fastest
  const float* rel_w;
  float* drel_h;
  float* drel_w;
  float* lse;
  float* delta;
  __nv_bfloat16* out;
  const __nv_bfloat16* dout;
  __nv_bfloat16* dqkv;
  long long row_stride, batch_stride, out_stride, out_batch_stride;
};

__device__ __forceinline__ uint32_t swz128(int row, int chunk16) {
  return (uint32_t)(row * 128 + ((chunk16 ^ (row & 7)) << 4));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2(float x) {      // 2^x, one MUFU (x <= 0 here; -inf -> 0)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// CTA-wide wait on an mbarrier: ONE warp polls (with a short sleep between polls), the other seven park on the hardware
// barrier.  Every polling warp costs issue slots whether one lane or 32 spin -- ncu on the first version of these kernels: 70 %
// of the issue slots busy at 8 % tensor-pipe activity, most of it 256 threads spinning on `try_wait` next to the co-resident
// CTA's exponentials.
__device__ __forceinline__ void cta_wait(uint64_t* bar, uint32_t parity, int warp) {
  if (warp == 0) {
    const uint32_t addr = smem_u32(bar);
    if (!mbar_try_wait(addr, parity)) {
      long long t0 = clock64();
      while (!mbar_try_wait(addr, parity)) {
        __nanosleep(40);
        if (clock64() - t0 > 40000000000LL) {
          printf("aldi_b200: attention mbarrier wait timeout (block %d,%d,%d)\n", (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z);
          __trap();
        }
      }
    }
  }
  __syncthreads();
}
// Scheduling fence for a register value: a prefetched global load must not be consumed (and so waited for) before this point
__device__ __forceinline__ void pin(float& x) { asm volatile("" : "+f"(x)); }
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

// K-major operand tile (rows = M or N index, 64 bf16 of K per 128-byte row): 16 K elements per MMA = +32 B
__device__ __forceinline__ uint64_t desc_kmajor(const uint8_t* tile) { return make_smem_desc_sw128(smem_u32(tile), 16, 1024); }
// MN-major operand (rows = K index, 64 bf16 of M/N per row; successive 64-wide M/N groups `lbo` bytes apart): 16 K rows = +2048 B
__device__ __forceinline__ uint64_t desc_mnmajor(const uint8_t* tile, uint32_t lbo) {
  return make_smem_desc_sw128(smem_u32(tile), lbo, 1024);
}

// ---------------------------------------------------------------------------------------------------------------------
// Thread layout of all three kernels: 256 threads, TWO threads per tile row (TMEM lane): warps 0-3 take key columns 0-63 of
// every tile (patch rows 0-3), warps 4-7 columns 64-127 (patch rows 4-7).  Thread 0 also issues TMA and MMA between the
// phases -- the products of the NEXT tile are issued right behind the current tile's second product, so the tensor pipe works
// while the rows are read back.  (A ninth issuer warp would put three warps on one scheduler and cap every thread at 168
// registers.)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kCompute = 256;

struct RowCtx {
  int r, half, qh, qw;
  bool valid;
  uint32_t lane_off;
};

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kFwdSmem = 6 * kTileBytes + 2048 + 1024;   // Q, K x2 (P chunk 0 reuses the K buffer S has consumed), V x2, P chunk 1,
                                                          // row-max / row-sum exchange between the two threads of a row

__global__ void __launch_bounds__(kCompute, 2)      // two CTAs per SM: one's softmax overlaps the other's MMAs / TMEM latency
attn_fwd_tc(const __grid_constant__ CUtensorMap tmQKV, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTileBytes;        // [2]
  uint8_t* sV = sK + 2 * kTileBytes;    // [2]
  uint8_t* sP1 = sV + 2 * kTileBytes;   // keys 64-127 of P; keys 0-63 go to sK[buf]
  float* sX = reinterpret_cast<float*>(sP1 + kTileBytes);   // [2][128]
  __shared__ __align__(8) uint64_t q_bar, kv_bar[2], s_bar, o_bar;
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int qh0 = (qt / a.tiles_w) * kPH, qw0 = (qt % a.tiles_w) * kPW;
  const int nkt = a.tiles_h * a.tiles_w;
  constexpr uint32_t idesc_s = make_idesc_bf16(kTile, 128, 0, 0);
  constexpr uint32_t idesc_pv = make_idesc_bf16(kTile, 64, 0, 1);

  if (tid == 0) {
    prefetch_tmap(&tmQKV);
    mbar_init(&q_bar, 1);
    mbar_init(&kv_bar[0], 1);
    mbar_init(&kv_bar[1], 1);
    mbar_init(&s_bar, 1);
    mbar_init(&o_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

  auto load_kv = [&](int kt) {       // thread 0: K and V patches of key tile kt into buffer kt & 1
    const int buf = kt & 1, kh0 = (kt / a.tiles_w) * kPH, kw0 = (kt % a.tiles_w) * kPW;
    mbar_expect_tx(&kv_bar[buf], 2 * kTileBytes);
    tma_load_4d(sK + buf * kTileBytes, &tmQKV, &kv_bar[buf], a.dim + h * 64, kw0, kh0, b);
    tma_load_4d(sV + buf * kTileBytes, &tmQKV, &kv_bar[buf], 2 * a.dim + h * 64, kw0, kh0, b);
  };
  auto issue_s = [&](int kt) {       // thread 0: S = Q K^T of key tile kt
    const int buf = kt & 1;
    mbar_wait(&kv_bar[buf], (kt >> 1) & 1);
    tc_fence_after();
    const uint64_t ad = desc_kmajor(sQ), bd = desc_kmajor(sK + buf * kTileBytes);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, ad + 2 * k, bd + 2 * k, idesc_s, k != 0);
    umma_commit(&s_bar);
  };
  if (tid == 0) {
    mbar_expect_tx(&q_bar, kTileBytes);
    tma_load_4d(sQ, &tmQKV, &q_bar, h * 64, qw0, qh0, b);
    load_kv(0);
    if (nkt > 1) load_kv(1);
    mbar_wait(&q_bar, 0);
    issue_s(0);
  }
  __syncwarp();

  const int r = tid & 127, half = tid >> 7;
  const int qh = qh0 + r / kPW, qw = qw0 + r % kPW;
  const bool valid_q = qh < a.gh && qw < a.gw;
  const int tn = a.gh * a.gw;
  const long long bh = (long long)b * a.heads + h;
  const float* rh = (a.rel_h && valid_q) ? a.rel_h + bh * a.gh * tn + qh * a.gw + qw : nullptr;
  const float* rw = (a.rel_h && valid_q) ? a.rel_w + bh * a.gw * tn + qh * a.gw + qw : nullptr;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  float o[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  float th[4], tw[kPW];
  // relative-position terms of this row for key tile kt (x log2 e): 4 key rows + 16 key columns, coalesced over the warp's rows.
  // Loaded one tile AHEAD (right after the previous tile's exponentials) so the L2 latency hides behind the MMA waits.
  auto load_bias = [&](int kt) {
    const int kh0 = (kt / a.tiles_w) * kPH + 4 * half, kw0 = (kt % a.tiles_w) * kPW;
#pragma unroll
    for (int i = 0; i < 4; ++i) th[i] = (rh && kh0 + i < a.gh) ? __ldg(rh + (long long)(kh0 + i) * tn) : 0.f;
#pragma unroll
    for (int i = 0; i < kPW; ++i) tw[i] = (rh && kw0 + i < a.gw) ? __ldg(rw + (long long)(kw0 + i) * tn) : 0.f;
  };
  load_bias(0);
  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    const int kh0 = (kt / a.tiles_w) * kPH + 4 * half, kw0 = (kt % a.tiles_w) * kPW;     // this thread's 4 x 16 keys
    const bool interior = (kh0 + 4 <= a.gh) && (kw0 + kPW <= a.gw);
    cta_wait(&s_bar, kt & 1, warp);
    tc_fence_after();
#pragma unroll
    for (int i = 0; i < 4; ++i) { pin(th[i]); th[i] *= kLog2e; }        // first use of the prefetched terms: after the wait
#pragma unroll
    for (int i = 0; i < kPW; ++i) { pin(tw[i]); tw[i] *= kLog2e; }
    float mt4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};      // independent chains: two warps per scheduler hide little
#pragma unroll
    for (int cl = 0; cl < 2; ++cl) {
      uint32_t raw[32];
      tmem_ld_32x32(tmem_s + lane_off + (2 * half + cl) * 32, raw);
      tmem_ld_wait();
      if (interior) {        // uniform branch: interior tiles (all but the last row / column of patches) skip the masks
#pragma unroll
        for (int i = 0; i < 32; ++i)
          mt4[i & 3] = fmaxf(mt4[i & 3], fmaf(__uint_as_float(raw[i]), a.scale_log2, th[2 * cl + (i >> 4)]) + tw[i & 15]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int ph = 2 * cl + (i >> 4), pw = i & 15;
          const float s2 = fmaf(__uint_as_float(raw[i]), a.scale_log2, th[ph]) + tw[pw];
          mt4[i & 3] = fmaxf(mt4[i & 3], ((kh0 + ph < a.gh) && (kw0 + pw < a.gw)) ? s2 : -INFINITY);
        }
      }
    }
    const float mt = fmaxf(fmaxf(mt4[0], mt4[1]), fmaxf(mt4[2], mt4[3]));
    sX[half * kTile + r] = mt;
    __syncthreads();
    const float mn = fmaxf(m, fmaxf(mt, sX[(half ^ 1) * kTile + r]));     // >= one valid key per tile: finite
    const float alpha = ex2(m - mn);
#pragma unroll
    for (int i = 0; i < 4; ++i) th[i] -= mn;        // the exponent's offset rides in the row term
    float lp4[4] = {0.f, 0.f, 0.f, 0.f};
    uint8_t* dst = half ? sP1 : sK + buf * kTileBytes;
#pragma unroll
    for (int cl = 0; cl < 2; ++cl) {
      uint32_t raw[32];
      tmem_ld_32x32(tmem_s + lane_off + (2 * half + cl) * 32, raw);
      tmem_ld_wait();
      float p[32];
      if (interior) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          p[i] = ex2(fmaf(__uint_as_float(raw[i]), a.scale_log2, th[2 * cl + (i >> 4)]) + tw[i & 15]);
          lp4[i & 3] += p[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int ph = 2 * cl + (i >> 4), pw = i & 15;
          const float e = ex2(fmaf(__uint_as_float(raw[i]), a.scale_log2, th[ph]) + tw[pw]);
          p[i] = ((kh0 + ph < a.gh) && (kw0 + pw < a.gw)) ? e : 0.f;
          lp4[i & 3] += p[i];
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 v;
        v.x = pack2(p[8 * g], p[8 * g + 1]);
        v.y = pack2(p[8 * g + 2], p[8 * g + 3]);
        v.z = pack2(p[8 * g + 4], p[8 * g + 5]);
        v.w = pack2(p[8 * g + 6], p[8 * g + 7]);
        *reinterpret_cast<uint4*>(dst + swz128(r, cl * 4 + g)) = v;
      }
    }
    const float lp = (lp4[0] + lp4[1]) + (lp4[2] + lp4[3]);
    if (kt + 1 < nkt) load_bias(kt + 1);
    m = mn;
    l = l * alpha + lp;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();     // P of this tile is in shared memory; every thread has finished with S
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const uint64_t ad = desc_kmajor(cc == 0 ? sK + buf * kTileBytes : sP1);
        const uint64_t bd = desc_mnmajor(sV + buf * kTileBytes + cc * 8192, 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_o, ad + 2 * k, bd + 128 * k, idesc_pv, (cc | k) != 0);
      }
      umma_commit(&o_bar);
      if (kt + 1 < nkt) issue_s(kt + 1);      // runs behind P V while the rows below read O back
    }
    __syncwarp();
    cta_wait(&o_bar, kt & 1, warp);
    tc_fence_after();
    if (tid == 0 && kt + 2 < nkt) load_kv(kt + 2);     // P V has finished with this buffer (V and the P chunk in the K slot)
    __syncwarp();
    {
      uint32_t raw[32];
      tmem_ld_32x32(tmem_o + lane_off + half * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], alpha, __uint_as_float(raw[i]));
    }
    tc_fence_before();
  }
  // row sum = both halves
  __syncthreads();
  sX[half * kTile + r] = l;
  __syncthreads();
  l += sX[(half ^ 1) * kTile + r];
  if (valid_q) {
    const float inv = 1.f / l;
    __nv_bfloat16* orow = a.out + (long long)b * a.out_batch_stride + (long long)(qh * a.gw + qw) * a.out_stride + h * 64 + half * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 v;
      v.x = pack2(o[8 * g] * inv, o[8 * g + 1] * inv);
      v.y = pack2(o[8 * g + 2] * inv, o[8 * g + 3] * inv);
      v.z = pack2(o[8 * g + 4] * inv, o[8 * g + 5] * inv);
      v.w = pack2(o[8 * g + 6] * inv, o[8 * g + 7] * inv);
      *reinterpret_cast<uint4*>(orow + 8 * g) = v;
    }
    if (half == 0) a.lse[((long long)b * a.heads + h) * tn + qh * a.gw + qw] = (m + log2f(l)) * kLn2;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward, query side: dq, drelpos, delta.  One CTA per SM (TMEM: S 128 + dP 128 + dQ 64 columns).
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int dq_smem_bytes(int tiles_w) {
  return 8 * kTileBytes + tiles_w * kPW * kTile * 4 + kPW * kTile * 4 + 1024;      // tiles, column-sum table, exchange buffer
}

__global__ void __launch_bounds__(kCompute, 1)
attn_bwd_dq_tc(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + kTileBytes;
  uint8_t* sK = sDO + kTileBytes;       // [2]
  uint8_t* sV = sK + 2 * kTileBytes;    // [2]
  uint8_t* sDS = sV + 2 * kTileBytes;   // [2 chunks of 64 keys]
  float* sDtw = reinterpret_cast<float*>(sDS + 2 * kTileBytes);   // [key column][row]: column sums of dS over the key rows
  __shared__ __align__(8) uint64_t q_bar, kv_bar[2], s_bar, o_bar;
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int qh0 = (qt / a.tiles_w) * kPH, qw0 = (qt % a.tiles_w) * kPW;
  const int nkt = a.tiles_h * a.tiles_w;
  // the upper half's per-tile column sums cross to the row's lower-half thread here (shared-memory float atomics are CAS loops)
  float* sXw = sDtw + a.tiles_w * kPW * kTile;                   // [16][row]
  constexpr uint32_t idesc_s = make_idesc_bf16(kTile, 128, 0, 0);
  constexpr uint32_t idesc_dq = make_idesc_bf16(kTile, 64, 0, 1);

  if (tid == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmDO);
    mbar_init(&q_bar, 1);
    mbar_init(&kv_bar[0], 1);
    mbar_init(&kv_bar[1], 1);
    mbar_init(&s_bar, 1);
    mbar_init(&o_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_smem);
  if (a.rel_h)
    for (int i = tid; i < a.tiles_w * kPW * kTile; i += kCompute) sDtw[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128, tmem_dq = tmem_base + 256;

  auto load_kv = [&](int kt) {
    const int buf = kt & 1, kh0 = (kt / a.tiles_w) * kPH, kw0 = (kt % a.tiles_w) * kPW;
    mbar_expect_tx(&kv_bar[buf], 2 * kTileBytes);
    tma_load_4d(sK + buf * kTileBytes, &tmQKV, &kv_bar[buf], a.dim + h * 64, kw0, kh0, b);
    tma_load_4d(sV + buf * kTileBytes, &tmQKV, &kv_bar[buf], 2 * a.dim + h * 64, kw0, kh0, b);
  };
  auto issue_s = [&](int kt) {       // S = Q K^T and dP = dO V^T of key tile kt
    const int buf = kt & 1;
    mbar_wait(&kv_bar[buf], (kt >> 1) & 1);
    tc_fence_after();
    const uint64_t qd = desc_kmajor(sQ), dod = desc_kmajor(sDO);
    const uint64_t kd = desc_kmajor(sK + buf * kTileBytes), vd = desc_kmajor(sV + buf * kTileBytes);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_dp, dod + 2 * k, vd + 2 * k, idesc_s, k != 0);
    umma_commit(&s_bar);
  };
  if (tid == 0) {
    mbar_expect_tx(&q_bar, 2 * kTileBytes);
    tma_load_4d(sQ, &tmQKV, &q_bar, h * 64, qw0, qh0, b);
    tma_load_4d(sDO, &tmDO, &q_bar, h * 64, qw0, qh0, b);
    load_kv(0);
    if (nkt > 1) load_kv(1);
    mbar_wait(&q_bar, 0);
    issue_s(0);
  }
  __syncwarp();

  const int r = tid & 127, half = tid >> 7;      // row of the query tile; which 64 key columns of every tile
  const int qh = qh0 + r / kPW, qw = qw0 + r % kPW;
  const bool valid_q = qh < a.gh && qw < a.gw;
  const int tn = a.gh * a.gw;
  const int qtok = valid_q ? qh * a.gw + qw : 0;
  const long long bh = (long long)b * a.heads + h;
  const bool has_rel = a.rel_h && valid_q;
  const float* rh = has_rel ? a.rel_h + bh * a.gh * tn + qtok : nullptr;
  const float* rw = has_rel ? a.rel_w + bh * a.gw * tn + qtok : nullptr;
  float* dh = has_rel ? a.drel_h + bh * a.gh * tn + qtok : nullptr;      // every (key row / column, query) element is written
  float* dw = has_rel ? a.drel_w + bh * a.gw * tn + qtok : nullptr;      // exactly once: no zero fill
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const long long stat = bh * tn + qtok;
  // delta = rowsum(dO * O); lse in the exp2 domain (both threads of a row compute them)
  float delta = 0.f;
  {
    const long long orow = (long long)b * a.out_batch_stride + (long long)qtok * a.out_stride + h * 64;
    const uint4* po = reinterpret_cast<const uint4*>(a.out + orow);
    const uint4* pd = reinterpret_cast<const uint4*>(a.dout + orow);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const uint4 vo = __ldg(po + g), vd = __ldg(pd + g);
      const __nv_bfloat162* ho = reinterpret_cast<const __nv_bfloat162*>(&vo);
      const __nv_bfloat162* hd = reinterpret_cast<const __nv_bfloat162*>(&vd);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fo = __bfloat1622float2(ho[k]), fd = __bfloat1622float2(hd[k]);
        delta += fo.x * fd.x + fo.y * fd.y;
      }
    }
  }
  const float lse2 = a.lse[stat] * kLog2e;
  if (valid_q && half == 0) a.delta[stat] = delta;
  float dth[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) dth[i] = 0.f;
  float th[4], tw[kPW];
  // this row's relative-position terms for key tile kt (x log2 e, minus the row's log-sum-exp), loaded one tile ahead
  auto load_bias = [&](int kt) {
    const int kh0 = (kt / a.tiles_w) * kPH + 4 * half, kw0 = (kt % a.tiles_w) * kPW;
#pragma unroll
    for (int i = 0; i < 4; ++i) th[i] = (rh && kh0 + i < a.gh) ? __ldg(rh + (long long)(kh0 + i) * tn) : 0.f;
#pragma unroll
    for (int i = 0; i < kPW; ++i) tw[i] = (rh && kw0 + i < a.gw) ? __ldg(rw + (long long)(kw0 + i) * tn) : 0.f;
  };
  load_bias(0);
  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    const int ktw = kt % a.tiles_w;
    const int kh0 = (kt / a.tiles_w) * kPH + 4 * half, kw0 = ktw * kPW;     // this thread's 4 x 16 keys
    float dtw[kPW];
#pragma unroll
    for (int i = 0; i < kPW; ++i) dtw[i] = 0.f;
    const bool interior = valid_q && (kh0 + 4 <= a.gh) && (kw0 + kPW <= a.gw);
    cta_wait(&s_bar, kt & 1, warp);
    tc_fence_after();
#pragma unroll
    for (int i = 0; i < 4; ++i) { pin(th[i]); th[i] = fmaf(th[i], kLog2e, -lse2); }      // first use of the prefetched terms
#pragma unroll
    for (int i = 0; i < kPW; ++i) { pin(tw[i]); tw[i] *= kLog2e; }
    {
      // both 32-column chunks of S and dP in flight before the one wait
      uint32_t rs[2][32], rp[2][32];
#pragma unroll
      for (int cl = 0; cl < 2; ++cl) {
        tmem_ld_32x32(tmem_s + lane_off + (2 * half + cl) * 32, rs[cl]);
        tmem_ld_32x32(tmem_dp + lane_off + (2 * half + cl) * 32, rp[cl]);
      }
      tmem_ld_wait();
#pragma unroll
      for (int cl = 0; cl < 2; ++cl) {
        float ds[32];
        if (interior) {        // uniform branch: no per-key masks away from the grid's last row / column of patches
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float p = ex2(fmaf(__uint_as_float(rs[cl][i]), a.scale_log2, th[2 * cl + (i >> 4)]) + tw[i & 15]);
            ds[i] = p * (__uint_as_float(rp[cl][i]) - delta);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int ph = 2 * cl + (i >> 4), pw = i & 15;
            const float p = ex2(fmaf(__uint_as_float(rs[cl][i]), a.scale_log2, th[ph]) + tw[pw]);
            ds[i] = (valid_q && (kh0 + ph < a.gh) && (kw0 + pw < a.gw)) ? p * (__uint_as_float(rp[cl][i]) - delta) : 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          dth[2 * cl + (i >> 4)] += ds[i];
          dtw[i & 15] += ds[i];
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack2(ds[8 * g], ds[8 * g + 1]);
          v.y = pack2(ds[8 * g + 2], ds[8 * g + 3]);
          v.z = pack2(ds[8 * g + 4], ds[8 * g + 5]);
          v.w = pack2(ds[8 * g + 6], ds[8 * g + 7]);
          *reinterpret_cast<uint4*>(sDS + half * kTileBytes + swz128(r, cl * 4 + g)) = v;
        }
      }
    }
    if (kt + 1 < nkt) load_bias(kt + 1);
    if (dh) {
      if (half) {
#pragma unroll
        for (int i = 0; i < kPW; ++i) sXw[i * kTile + r] = dtw[i];
      }
      if (ktw == a.tiles_w - 1) {      // this band of key rows is complete
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (kh0 + i < a.gh) dh[(long long)(kh0 + i) * tn] = dth[i];
          dth[i] = 0.f;
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();     // dS of this tile is in shared memory; every thread has finished with S and dP
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const uint64_t ad = desc_kmajor(sDS + cc * kTileBytes);
        const uint64_t bd = desc_mnmajor(sK + buf * kTileBytes + cc * 8192, 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_dq, ad + 2 * k, bd + 128 * k, idesc_dq, (kt | cc | k) != 0);      // dQ += dS K
      }
      umma_commit(&o_bar);
      if (kt + 1 < nkt) issue_s(kt + 1);
    }
    __syncwarp();
    if (dh && !half) {      // one owner per (key column, row): plain read-modify-write; the barrier inside the wait below
#pragma unroll                // separates these reads from the next tile's writes of the exchange buffer
      for (int i = 0; i < kPW; ++i) sDtw[(kw0 + i) * kTile + r] += dtw[i] + sXw[i * kTile + r];
    }
    cta_wait(&o_bar, kt & 1, warp);    // dS buffer and this tile's K / V buffers are free again
    if (tid == 0 && kt + 2 < nkt) load_kv(kt + 2);
    __syncwarp();
  }
  tc_fence_after();
  // tcgen05.ld is warp-collective: rows outside the grid load too, only the stores are predicated
  {
    uint32_t raw[32];
    tmem_ld_32x32(tmem_dq + lane_off + half * 32, raw);
    tmem_ld_wait();
    if (valid_q) {
      __nv_bfloat16* dqrow = a.dqkv + (long long)b * a.batch_stride + (long long)qtok * a.row_stride + h * 64 + half * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 v;
        v.x = pack2(__uint_as_float(raw[8 * g]) * a.scale, __uint_as_float(raw[8 * g + 1]) * a.scale);
        v.y = pack2(__uint_as_float(raw[8 * g + 2]) * a.scale, __uint_as_float(raw[8 * g + 3]) * a.scale);
        v.z = pack2(__uint_as_float(raw[8 * g + 4]) * a.scale, __uint_as_float(raw[8 * g + 5]) * a.scale);
        v.w = pack2(__uint_as_float(raw[8 * g + 6]) * a.scale, __uint_as_float(raw[8 * g + 7]) * a.scale);
        *reinterpret_cast<uint4*>(dqrow + 8 * g) = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();      // all shared-memory column sums are final
  if (dw) {
    const int k0 = half ? a.gw / 2 : 0, k1 = half ? a.gw : a.gw / 2;
    for (int kw = k0; kw < k1; ++kw) dw[(long long)kw * tn] = sDtw[kw * kTile + r];
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward, key side: dk, dv   (CTA = key tile, loop over query tiles; the two threads of a row split the CTA's keys)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kDkvSmem = 10 * kTileBytes + 1024;   // K, V, Q x2, dO x2, P (2 chunks), dS (2 chunks)

__global__ void __launch_bounds__(kCompute, 1)
attn_bwd_dkv_tc(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sK = smem;
  uint8_t* sV = sK + kTileBytes;
  uint8_t* sQ = sV + kTileBytes;        // [2]
  uint8_t* sDO = sQ + 2 * kTileBytes;   // [2]
  uint8_t* sP = sDO + 2 * kTileBytes;   // [2 chunks of 64 keys]
  uint8_t* sDS = sP + 2 * kTileBytes;   // [2 chunks]
  __shared__ __align__(8) uint64_t k_bar, qd_bar[2], s_bar, o_bar;
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int kh0 = (kt / a.tiles_w) * kPH, kw0 = (kt % a.tiles_w) * kPW;
  const int nqt = a.tiles_h * a.tiles_w;
  constexpr uint32_t idesc_s = make_idesc_bf16(kTile, 128, 0, 0);
  constexpr uint32_t idesc_t = make_idesc_bf16(kTile, 64, 1, 1);

  if (tid == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmDO);
    mbar_init(&k_bar, 1);
    mbar_init(&qd_bar[0], 1);
    mbar_init(&qd_bar[1], 1);
    mbar_init(&s_bar, 1);
    mbar_init(&o_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128, tmem_dv = tmem_base + 256, tmem_dk = tmem_base + 320;

  auto load_q = [&](int qt) {
    const int buf = qt & 1, qh0 = (qt / a.tiles_w) * kPH, qw0 = (qt % a.tiles_w) * kPW;
    mbar_expect_tx(&qd_bar[buf], 2 * kTileBytes);
    tma_load_4d(sQ + buf * kTileBytes, &tmQKV, &qd_bar[buf], h * 64, qw0, qh0, b);
    tma_load_4d(sDO + buf * kTileBytes, &tmDO, &qd_bar[buf], h * 64, qw0, qh0, b);
  };
  auto issue_s = [&](int qt) {       // S = Q K^T and dP = dO V^T of query tile qt
    const int buf = qt & 1;
    mbar_wait(&qd_bar[buf], (qt >> 1) & 1);
    tc_fence_after();
    const uint64_t qd = desc_kmajor(sQ + buf * kTileBytes), dod = desc_kmajor(sDO + buf * kTileBytes);
    const uint64_t kd = desc_kmajor(sK), vd = desc_kmajor(sV);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_dp, dod + 2 * k, vd + 2 * k, idesc_s, k != 0);
    umma_commit(&s_bar);
  };
  if (tid == 0) {
    mbar_expect_tx(&k_bar, 2 * kTileBytes);
    tma_load_4d(sK, &tmQKV, &k_bar, a.dim + h * 64, kw0, kh0, b);
    tma_load_4d(sV, &tmQKV, &k_bar, 2 * a.dim + h * 64, kw0, kh0, b);
    load_q(0);
    if (nqt > 1) load_q(1);
    mbar_wait(&k_bar, 0);
    issue_s(0);
  }
  __syncwarp();

  const int r = tid & 127, half = tid >> 7;
  const int tn = a.gh * a.gw;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const int khh = kh0 + 4 * half;            // this thread's 4 x 16 keys of the CTA's key tile
  float th[4], tw[kPW], delta, lse2;
  bool valid_q;
  // per query tile: this row's validity, delta and relative-position terms (x log2 e, minus the row's log-sum-exp) for the
  // CTA's keys -- loaded one tile ahead
  auto load_row = [&](int qt) {
    const int qh = (qt / a.tiles_w) * kPH + r / kPW, qw = (qt % a.tiles_w) * kPW + r % kPW;
    valid_q = qh < a.gh && qw < a.gw;
    const int qtok = valid_q ? qh * a.gw + qw : 0;
    const long long bh = (long long)b * a.heads + h;
    const float* rh = (a.rel_h && valid_q) ? a.rel_h + bh * a.gh * tn + qtok : nullptr;
    const float* rw = (a.rel_h && valid_q) ? a.rel_w + bh * a.gw * tn + qtok : nullptr;
    lse2 = __ldg(a.lse + bh * tn + qtok);
    delta = __ldg(a.delta + bh * tn + qtok);
#pragma unroll
    for (int i = 0; i < 4; ++i) th[i] = (rh && khh + i < a.gh) ? __ldg(rh + (long long)(khh + i) * tn) : 0.f;
#pragma unroll
    for (int i = 0; i < kPW; ++i) tw[i] = (rh && kw0 + i < a.gw) ? __ldg(rw + (long long)(kw0 + i) * tn) : 0.f;
  };
  load_row(0);
  for (int qt = 0; qt < nqt; ++qt) {
    const int buf = qt & 1;
    const bool interior = valid_q && (khh + 4 <= a.gh) && (kw0 + kPW <= a.gw);
    const bool vq = valid_q;
    cta_wait(&s_bar, qt & 1, warp);
    tc_fence_after();
    pin(lse2);
    pin(delta);
    {
      const float l2 = lse2 * kLog2e;
#pragma unroll
      for (int i = 0; i < 4; ++i) { pin(th[i]); th[i] = fmaf(th[i], kLog2e, -l2); }      // first use of the prefetched terms
#pragma unroll
      for (int i = 0; i < kPW; ++i) { pin(tw[i]); tw[i] *= kLog2e; }
    }
    {
      uint32_t rs[2][32], rp[2][32];
#pragma unroll
      for (int cl = 0; cl < 2; ++cl) {
        tmem_ld_32x32(tmem_s + lane_off + (2 * half + cl) * 32, rs[cl]);
        tmem_ld_32x32(tmem_dp + lane_off + (2 * half + cl) * 32, rp[cl]);
      }
      tmem_ld_wait();
#pragma unroll
      for (int cl = 0; cl < 2; ++cl) {
        float p[32], ds[32];
        if (interior) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            p[i] = ex2(fmaf(__uint_as_float(rs[cl][i]), a.scale_log2, th[2 * cl + (i >> 4)]) + tw[i & 15]);
            ds[i] = p[i] * (__uint_as_float(rp[cl][i]) - delta);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int ph = 2 * cl + (i >> 4), pw = i & 15;
            const float pv = ex2(fmaf(__uint_as_float(rs[cl][i]), a.scale_log2, th[ph]) + tw[pw]);
            p[i] = (vq && (khh + ph < a.gh) && (kw0 + pw < a.gw)) ? pv : 0.f;
            ds[i] = p[i] * (__uint_as_float(rp[cl][i]) - delta);
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v, w;
          v.x = pack2(p[8 * g], p[8 * g + 1]);
          v.y = pack2(p[8 * g + 2], p[8 * g + 3]);
          v.z = pack2(p[8 * g + 4], p[8 * g + 5]);
          v.w = pack2(p[8 * g + 6], p[8 * g + 7]);
          w.x = pack2(ds[8 * g], ds[8 * g + 1]);
          w.y = pack2(ds[8 * g + 2], ds[8 * g + 3]);
          w.z = pack2(ds[8 * g + 4], ds[8 * g + 5]);
          w.w = pack2(ds[8 * g + 6], ds[8 * g + 7]);
          const uint32_t off = half * kTileBytes + swz128(r, cl * 4 + g);
          *reinterpret_cast<uint4*>(sP + off) = v;
          *reinterpret_cast<uint4*>(sDS + off) = w;
        }
      }
    }
    if (qt + 1 < nqt) load_row(qt + 1);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();     // P and dS of this tile are in shared memory; every thread has finished with S and dP
    if (tid == 0) {
      tc_fence_after();
      // A = P^T / dS^T: the [query row][key] tiles read MN-major (K = query rows, M = keys: two 64-key groups one chunk apart)
      const uint64_t pd = desc_mnmajor(sP, kTileBytes), dsd = desc_mnmajor(sDS, kTileBytes);
      const uint64_t dod = desc_mnmajor(sDO + buf * kTileBytes, 8192), qd = desc_mnmajor(sQ + buf * kTileBytes, 8192);
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_bf16(tmem_dv, pd + 128 * k, dod + 128 * k, idesc_t, (qt | k) != 0);   // dV += P^T dO
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_bf16(tmem_dk, dsd + 128 * k, qd + 128 * k, idesc_t, (qt | k) != 0);   // dK += dS^T Q
      umma_commit(&o_bar);
      if (qt + 1 < nqt) issue_s(qt + 1);
    }
    __syncwarp();
    cta_wait(&o_bar, qt & 1, warp);
    if (tid == 0 && qt + 2 < nqt) load_q(qt + 2);
    __syncwarp();
  }
  tc_fence_after();
  // TMEM lane = key row r of the CTA's key tile; threads 0-127 write dV, threads 128-255 dK
  {
    const int kh = kh0 + r / kPW, kw = kw0 + r % kPW;
    const bool valid_k = kh < a.gh && kw < a.gw;
    const float sc = half ? a.scale : 1.f;
    __nv_bfloat16* dst = a.dqkv + (long long)b * a.batch_stride + (long long)(valid_k ? kh * a.gw + kw : 0) * a.row_stride + h * 64 +
                         (half ? a.dim : 2 * a.dim);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32((half ? tmem_dk : tmem_dv) + lane_off + c * 32, raw);
      tmem_ld_wait();
      if (valid_k) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack2(__uint_as_float(raw[8 * g]) * sc, __uint_as_float(raw[8 * g + 1]) * sc);
          v.y = pack2(__uint_as_float(raw[8 * g + 2]) * sc, __uint_as_float(raw[8 * g + 3]) * sc);
          v.z = pack2(__uint_as_float(raw[8 * g + 4]) * sc, __uint_as_float(raw[8 * g + 5]) * sc);
          v.w = pack2(__uint_as_float(raw[8 * g + 6]) * sc, __uint_as_float(raw[8 * g + 7]) * sc);
          *reinterpret_cast<uint4*>(dst + c * 32 + 8 * g) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
int check_tc(const aldi_attn_params* p, const char* who) {
  ALDI_CHECK_ARG((reinterpret_cast<uintptr_t>(p->qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->out) & 15) == 0,
                 "%s: qkv / out must be 16-byte aligned", who);
  ALDI_CHECK_ARG(p->row_stride % 8 == 0 && p->batch_stride % 8 == 0 && p->out_stride % 8 == 0 && p->out_batch_stride % 8 == 0,
                 "%s: strides must be multiples of 8 elements", who);
  ALDI_CHECK_ARG(p->batch <= 65535 && p->heads <= 65535, "%s: batch / heads exceed the grid limits", who);
  return ALDI_OK;
}

TcArgs make_tc_args(const aldi_attn_params* p) {
  TcArgs a;
  a.gh = p->gh; a.gw = p->gw; a.heads = p->heads; a.dim = p->heads * 64;
  a.tiles_h = aldi_div_up(p->gh, kPH); a.tiles_w = aldi_div_up(p->gw, kPW);
  a.scale = p->scale; a.scale_log2 = p->scale * kLog2e;
  a.rel_h = p->rel_h; a.rel_w = p->rel_w; a.drel_h = p->drel_h; a.drel_w = p->drel_w; a.lse = p->lse; a.delta = p->delta;
  a.out = reinterpret_cast<__nv_bfloat16*>(p->out);
  a.dout = reinterpret_cast<const __nv_bfloat16*>(p->dout);
  a.dqkv = reinterpret_cast<__nv_bfloat16*>(p->dqkv);
  a.row_stride = p->row_stride; a.batch_stride = p->batch_stride;
  a.out_stride = p->out_stride; a.out_batch_stride = p->out_batch_stride;
  return a;
}

int make_token_map(CUtensorMap* tm, const void* base, int width, int gw, int gh, int batch, long long row_stride,
                   long long batch_stride) {
  uint64_t dims[4] = {(uint64_t)width, (uint64_t)gw, (uint64_t)gh, (uint64_t)batch};
  uint64_t strides[3] = {(uint64_t)row_stride * 2, (uint64_t)row_stride * gw * 2, (uint64_t)batch_stride * 2};
  uint32_t box[4] = {64, kPW, kPH, 1};
  return aldi_make_tmap_bf16(tm, base, 4, dims, strides, box);
}

template <typename K>
int set_smem(K kernel, int bytes, const char* who) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    aldi_set_error("%s: cudaFuncSetAttribute(%d bytes) failed: %s", who, bytes, cudaGetErrorString(e));
    return ALDI_ERR_CUDA;
  }
  return ALDI_OK;
}

}  // namespace

int aldi_attention_forward_tc(const aldi_attn_params* p, cudaStream_t stream) {
  int rc = check_tc(p, "aldi_attention_forward");
  if (rc) return rc;
  const TcArgs a = make_tc_args(p);
  CUtensorMap tmQKV;
  rc = make_token_map(&tmQKV, p->qkv, 3 * a.dim, p->gw, p->gh, p->batch, p->row_stride, p->batch_stride);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    rc = set_smem(attn_fwd_tc, kFwdSmem, "aldi_attention_forward");
    if (rc) return rc;
    attr = true;
  }
  const dim3 grid(a.tiles_h * a.tiles_w, p->heads, p->batch);
  attn_fwd_tc<<<grid, kCompute, kFwdSmem, stream>>>(tmQKV, a);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_attention_forward");
  return ALDI_OK;
}

int aldi_attention_backward_tc(const aldi_attn_params* p, cudaStream_t stream) {
  int rc = check_tc(p, "aldi_attention_backward");
  if (rc) return rc;
  ALDI_CHECK_ARG((reinterpret_cast<uintptr_t>(p->dout) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->dqkv) & 15) == 0,
                 "aldi_attention_backward: dout / dqkv must be 16-byte aligned");
  const TcArgs a = make_tc_args(p);
  CUtensorMap tmQKV, tmDO;
  rc = make_token_map(&tmQKV, p->qkv, 3 * a.dim, p->gw, p->gh, p->batch, p->row_stride, p->batch_stride);
  if (rc) return rc;
  rc = make_token_map(&tmDO, p->dout, a.dim, p->gw, p->gh, p->batch, p->out_stride, p->out_batch_stride);
  if (rc) return rc;
  const int dq_smem = dq_smem_bytes(a.tiles_w);
  ALDI_CHECK_ARG(dq_smem <= 227 * 1024, "aldi_attention_backward: token grid %d wide needs %d bytes of shared memory", p->gw, dq_smem);
  static int dq_attr = 0;
  if (dq_attr < dq_smem) {
    rc = set_smem(attn_bwd_dq_tc, dq_smem, "aldi_attention_backward(dq)");
    if (rc) return rc;
    dq_attr = dq_smem;
  }
  static bool dkv_attr = false;
  if (!dkv_attr) {
    rc = set_smem(attn_bwd_dkv_tc, kDkvSmem, "aldi_attention_backward(dkv)");
    if (rc) return rc;
    dkv_attr = true;
  }
  const dim3 grid(a.tiles_h * a.tiles_w, p->heads, p->batch);
  attn_bwd_dq_tc<<<grid, kCompute, dq_smem, stream>>>(tmQKV, tmDO, a);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_attention_backward(dq)");
  attn_bwd_dkv_tc<<<grid, kCompute, kDkvSmem, stream>>>(tmQKV, tmDO, a);
  ALDI_COUNT_LAUNCH();
  ALDI_CUDA_LAUNCH_CHECK("aldi_attention_backward(dkv)");
  return ALDI_OK;
}
