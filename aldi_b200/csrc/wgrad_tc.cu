// Weight-gradient GEMM on tcgen05:  dW[cout, tap, cin] += scale[cout] * sum_pixels dY[pixel, cout] * X[pixel+tap, cin]
//
// The reduction (K) dimension is the PIXEL index, which is the slow axis of both channels-last operands,
// so both UMMA operands are MN-major: a TMA box {64 ch, TW, TH, 1} (TW*TH = 64 pixels) lands as 64 rows
// (pixels = K) x 128 B (64 channels = M or N), 128-byte swizzled -> canonical MN-major SW128 layout with
// SBO = 1024 B (next 8 pixels) and LBO = 8192 B (next 64-channel box).  The tap shift of X and the
// conv zero padding are again TMA coordinates + OOB zero fill.  Split-K over pixel tiles, partial
// results reduced with vector fp32 atomics straight into the flat gradient buffer.
//
// The kernel is bound by L2 -> shared-memory traffic (every (tap, cout tile) item re-streams x), so layers
// with >= 256 output channels use a 256-row M tile: two 128 x BLOCK_N accumulators share each x stage, which
// halves the number of times x crosses the L2 (TMEM then holds 2 x 256 columns, single-buffered: items are
// thousands of pixels long, the un-overlapped epilogue is noise).
//
// Replaces cuDNN wgrad / cuBLAS (linear weight grad) under autograd of detectron2 layers
// (reached from aldi/trainer.py:79 `trainer.do_backward`).
#include "common.cuh"
#include "sm100.cuh"
#include "tmap.h"
#include "../../include/aldi_b200.h"
#include <stdlib.h>

using namespace sm100;

namespace {

constexpr int kBlockM = 128;   // cout rows per MMA (one or two of them per item)
constexpr int kPix = 64;       // pixels (K) per stage
constexpr int kBoxBytes = kPix * 128;  // 8 KB per 64-channel box
constexpr int kNumThreads = 320;   // warps: 0 TMA, 1 MMA, 2-5 epilogue, 6-9 bias-gradient column sums
constexpr int kSumWarps = 4;

struct WgradArgs {
  int tiles_h, tiles_w, th, tw, pix_tiles;
  int taps, taps_w, pad_h, pad_w;
  int m_tiles, n_tiles, ksplit, num_items;
  const float* scale;
  float* dw;
  int cout_store, cin_store;
  float* dbias;   // bias gradient fused into this pass (nullable): column sums of the dy stages already in shared memory
  int tap_pair;   // cin == 128: one N = 256 MMA covers TWO taps (x boxes of tap 2i | tap 2i+1 side by side) — a
                  // 128 x N x 16 tcgen05.mma costs the same for N = 128 and N = 256; `taps` then counts tap PAIRS
  int taps_real;
};

template <int BLOCK_N, int MH>
struct WCfg {
  static constexpr int kNB = BLOCK_N / 64;
  static constexpr int kABoxes = 2 * MH;                      // 64-channel dy boxes per stage
  static constexpr int kStageBytes = (kABoxes + kNB) * kBoxBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kAccBufs = (2 * MH * BLOCK_N <= 512) ? 2 : 1;
  static constexpr int kColsUsed = kAccBufs * MH * BLOCK_N;
  static constexpr int kTmemCols = (kColsUsed <= 128) ? 128 : (kColsUsed <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;
};

__device__ __forceinline__ void decode_item(const WgradArgs& a, int item, int& tap, int& nt, int& mt, int& ks) {
  tap = item % a.taps;
  item /= a.taps;
  nt = item % a.n_tiles;
  item /= a.n_tiles;
  mt = item % a.m_tiles;
  ks = item / a.m_tiles;
}

template <int BLOCK_N, int MH>
__global__ void __launch_bounds__(kNumThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, const WgradArgs a) {
  using C = WCfg<BLOCK_N, MH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  __shared__ __align__(8) uint64_t full_bar[C::kStages];
  __shared__ __align__(8) uint64_t empty_bar[C::kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmY);
    prefetch_tmap(&tmX);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], a.dbias ? 1 + kSumWarps : 1);   // the MMA commit + (if present) the column-sum warps
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();               // prologue above overlapped the previous kernel's tail
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x) {
        int tap, nt, mt, ks;
        decode_item(a, item, tap, nt, mt, ks);
        const int r = tap / a.taps_w, s = tap - r * a.taps_w;
        const int pt0 = (int)((long long)a.pix_tiles * ks / a.ksplit);
        const int pt1 = (int)((long long)a.pix_tiles * (ks + 1) / a.ksplit);
        for (int pt = pt0; pt < pt1; ++pt) {
          int t = pt;
          const int twi = t % a.tiles_w;
          t /= a.tiles_w;
          const int thi = t % a.tiles_h;
          const int img = t / a.tiles_h;
          const int h0 = thi * a.th, w0 = twi * a.tw;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABoxes * kBoxBytes;
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
#pragma unroll
          for (int b = 0; b < C::kABoxes; ++b)
            tma_load_4d(sa + b * kBoxBytes, &tmY, &full_bar[stage], mt * kBlockM * MH + b * 64, w0, h0, img);
          if (a.tap_pair) {
            // boxes 0,1: channels 0-63 / 64-127 of tap 2*tap; boxes 2,3: the same of tap 2*tap+1 (clamped: the odd
            // ninth tap is computed twice and stored once)
#pragma unroll
            for (int b = 0; b < C::kNB; ++b) {
              int t2 = 2 * tap + (b >> 1);
              if (t2 >= a.taps_real) t2 = a.taps_real - 1;
              const int r2 = t2 / a.taps_w, s2 = t2 - r2 * a.taps_w;
              tma_load_4d(sb + b * kBoxBytes, &tmX, &full_bar[stage], (b & 1) * 64, w0 + s2 - a.pad_w, h0 + r2 - a.pad_h,
                          img);
            }
          } else {
#pragma unroll
            for (int b = 0; b < C::kNB; ++b)
              tma_load_4d(sb + b * kBoxBytes, &tmX, &full_bar[stage], nt * BLOCK_N + b * 64, w0 + s - a.pad_w,
                          h0 + r - a.pad_h, img);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++it) {
        int tap, nt, mt, ks;
        decode_item(a, item, tap, nt, mt, ks);
        const int pt0 = (int)((long long)a.pix_tiles * ks / a.ksplit);
        const int pt1 = (int)((long long)a.pix_tiles * (ks + 1) / a.ksplit);
        const int acc = it % C::kAccBufs;
        const uint32_t acc_phase = (it / C::kAccBufs) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * MH * BLOCK_N;
        for (int pt = pt0; pt < pt1; ++pt) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint64_t bdesc = make_smem_desc_sw128(sa + C::kABoxes * kBoxBytes, kBoxBytes, 1024);
#pragma unroll
          for (int half = 0; half < MH; ++half) {
            const uint64_t adesc = make_smem_desc_sw128(sa + half * 2 * kBoxBytes, kBoxBytes, 1024);
#pragma unroll
            for (int k = 0; k < kPix / 16; ++k) {
              // 16 pixels (K) = two 8-row swizzle atoms = 2048 B -> +128 in the (addr>>4) field
              umma_bf16(d_tmem + half * BLOCK_N, adesc + 128 * k, bdesc + 128 * k, idesc, (pt > pt0) || (k != 0));
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else if (warp >= 6) {
    // ---- bias gradient: dbias[co] += sum over pixels of dy[pixel, co], from the dy boxes the MMA is consuming anyway.
    // Items with tap 0 and cin tile 0 cover every (cout tile, pixel range) exactly once.  A box is 64 pixel rows x
    // 128 B (64 channels), 128-byte swizzled: lane l reads 16-byte chunk (l & 7) of rows (l >> 3) + 4 i.
    if (a.dbias) {
      const int sw = warp - 6;   // this warp sums boxes sw, sw + kSumWarps, ...
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x) {
        int tap, nt, mt, ks;
        decode_item(a, item, tap, nt, mt, ks);
        const bool mine = tap == 0 && nt == 0;
        const int pt0 = (int)((long long)a.pix_tiles * ks / a.ksplit);
        const int pt1 = (int)((long long)a.pix_tiles * (ks + 1) / a.ksplit);
        float acc[C::kABoxes][8];
#pragma unroll
        for (int b = 0; b < C::kABoxes; ++b)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[b][k] = 0.f;
        for (int pt = pt0; pt < pt1; ++pt) {
          mbar_wait(&full_bar[stage], phase);
          if (mine) {
            const uint8_t* sa = smem + stage * C::kStageBytes;
            const int chunk = lane & 7, r0 = lane >> 3;
#pragma unroll
            for (int b = 0; b < C::kABoxes; ++b) {
              if ((b % kSumWarps) != sw) continue;
#pragma unroll 4
              for (int i = 0; i < 16; ++i) {
                const int r = r0 + 4 * i;
                const uint4 q4 = *reinterpret_cast<const uint4*>(sa + b * kBoxBytes + r * 128 + ((chunk ^ (r & 7)) << 4));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 f = __bfloat1622float2(h[k]);
                  acc[b][2 * k] += f.x;
                  acc[b][2 * k + 1] += f.y;
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (mine) {
#pragma unroll
          for (int b = 0; b < C::kABoxes; ++b) {
            if ((b % kSumWarps) != sw) continue;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float v = acc[b][k];
              v += __shfl_xor_sync(0xffffffffu, v, 8);
              v += __shfl_xor_sync(0xffffffffu, v, 16);
              const int cout = mt * kBlockM * MH + b * 64 + (lane & 7) * 8 + k;
              if (lane < 8 && cout < a.cout_store) atomicAdd(a.dbias + cout, v);
            }
          }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool vec_ok = (a.cin_store % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.dw) & 15) == 0);
    int it = 0;
    for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++it) {
      int tap, nt, mt, ks;
      decode_item(a, item, tap, nt, mt, ks);
      const int acc = it % C::kAccBufs;
      const uint32_t acc_phase = (it / C::kAccBufs) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int hc = 0; hc < MH * (BLOCK_N / 32); ++hc) {
        const int half = hc / (BLOCK_N / 32), c0 = (hc - half * (BLOCK_N / 32)) * 32;
        const int cout = (mt * MH + half) * kBlockM + row;
        const bool valid = cout < a.cout_store;
        const float sc = (valid && a.scale) ? __ldg(a.scale + cout) : 1.f;
        int tap_out = tap, cbase = nt * BLOCK_N + c0;
        bool store = true;
        if (a.tap_pair) {
          tap_out = 2 * tap + (c0 >= 128 ? 1 : 0);
          cbase = c0 & 127;
          store = tap_out < a.taps_real;
        }
        float* drow = a.dw + ((long long)cout * a.taps_real + tap_out) * a.cin_store;
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MH + half) * BLOCK_N + c0), raw);
        tmem_ld_wait();
        if (valid && store && cbase < a.cin_store) {
          if (vec_ok && cbase + 32 <= a.cin_store) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float4 v = make_float4(__uint_as_float(raw[4 * g]) * sc, __uint_as_float(raw[4 * g + 1]) * sc,
                                     __uint_as_float(raw[4 * g + 2]) * sc, __uint_as_float(raw[4 * g + 3]) * sc);
              atomicAdd(reinterpret_cast<float4*>(drow + cbase) + g, v);
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (cbase + j < a.cin_store) atomicAdd(drow + cbase + j, __uint_as_float(raw[j]) * sc);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N, int MH>
int launch_wgrad(const CUtensorMap& tmY, const CUtensorMap& tmX, const WgradArgs& a, cudaStream_t stream) {
  using C = WCfg<BLOCK_N, MH>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BLOCK_N, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::kSmemBytes);
    if (e != cudaSuccess) {
      aldi_set_error("aldi_wgrad_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ALDI_ERR_CUDA;
    }
    attr_set = true;
  }
  int grid = a.num_items < aldi_num_sms() ? a.num_items : aldi_num_sms();
  cudaError_t le = aldi_launch_pdl(wgrad_tc_kernel<BLOCK_N, MH>, dim3(grid), dim3(kNumThreads), (size_t)C::kSmemBytes,
                                   stream, tmY, tmX, a);
  ALDI_COUNT_LAUNCH();
  if (le != cudaSuccess) {
    aldi_set_error("aldi_wgrad_tc: launch failed: %s", cudaGetErrorString(le));
    return ALDI_ERR_CUDA;
  }
  ALDI_CUDA_LAUNCH_CHECK("aldi_wgrad_tc");
  return ALDI_OK;
}

void pick_patch64(int ho, int wo, int* th, int* tw) {
  long best = -1;
  for (int t = 64; t >= 8; t >>= 1) {
    int hh = 64 / t;
    long cover = (long)((wo + t - 1) / t) * t * (long)((ho + hh - 1) / hh) * hh;
    if (best < 0 || cover < best) {
      best = cover;
      *tw = t;
      *th = hh;
    }
  }
}

}  // namespace

extern "C" int aldi_wgrad_tc(const aldi_wgrad_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(p && p->x && p->dy && p->dw, "aldi_wgrad_tc: null pointer");
  ALDI_CHECK_ARG(p->x_c > 0 && p->x_c % 64 == 0, "aldi_wgrad_tc: x_c=%d must be a multiple of 64", p->x_c);
  ALDI_CHECK_ARG(p->dy_c > 0 && p->dy_c % 64 == 0, "aldi_wgrad_tc: dy_c=%d must be a multiple of 64", p->dy_c);
  ALDI_CHECK_ARG(p->cout_store > 0 && p->cout_store <= p->dy_c && p->cin_store > 0 && p->cin_store <= p->x_c,
                 "aldi_wgrad_tc: bad cout_store/cin_store");
  ALDI_CHECK_ARG(p->n > 0 && p->ho > 0 && p->wo > 0, "aldi_wgrad_tc: empty");
  ALDI_CHECK_ARG(p->stride <= 1, "aldi_wgrad_tc: stride %d unsupported, pass a strided view of x", p->stride);
  ALDI_CHECK_ARG((p->x_sw % 8) == 0 && (p->x_sh % 8) == 0 && (p->x_sn % 8) == 0 && (p->dy_sw % 8) == 0 &&
                     (p->dy_sh % 8) == 0 && (p->dy_sn % 8) == 0,
                 "aldi_wgrad_tc: strides must be multiples of 8 elements");
  ALDI_CHECK_ARG((reinterpret_cast<uintptr_t>(p->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->dy) & 15) == 0,
                 "aldi_wgrad_tc: x/dy must be 16-byte aligned");

  int block_n = (p->x_c % 256 == 0) ? 256 : (p->x_c % 128 == 0) ? 128 : 64;
  static const char* no_pair = getenv("ALDI_WGRAD_NO_TAP_PAIR");
  const bool tap_pair = !no_pair && p->x_c == 128 && p->cin_store == 128 && p->taps_h * p->taps_w > 1;
  if (tap_pair) block_n = 256;
  int th, tw;
  pick_patch64(p->ho, p->wo, &th, &tw);

  WgradArgs a;
  a.th = th; a.tw = tw;
  a.tiles_h = aldi_div_up(p->ho, th);
  a.tiles_w = aldi_div_up(p->wo, tw);
  a.pix_tiles = p->n * a.tiles_h * a.tiles_w;
  a.taps_real = p->taps_h * p->taps_w;
  a.tap_pair = tap_pair ? 1 : 0;
  a.taps = tap_pair ? (a.taps_real + 1) / 2 : a.taps_real;
  a.taps_w = p->taps_w;
  a.pad_h = p->pad_h; a.pad_w = p->pad_w;
  // 256-row M tile when cout allows and the items are long enough (split-K atomics dominate short ones)
  const int pix_tiles0 = p->n * aldi_div_up(p->ho, th) * aldi_div_up(p->wo, tw);
  static const char* force_mh1 = getenv("ALDI_WGRAD_MH1");
  static const char* force_mh2 = getenv("ALDI_WGRAD_MH2");  // test knob: the 256-row tile at any size
  const bool long_items = force_mh2 || ((p->taps_h * p->taps_w > 1) ? pix_tiles0 >= 1024 : pix_tiles0 >= 4096);
  const int mh = (!force_mh1 && p->dy_c % 256 == 0 && p->cout_store > kBlockM && long_items) ? 2 : 1;
  a.m_tiles = aldi_div_up(p->cout_store, kBlockM * mh);
  a.n_tiles = tap_pair ? 1 : aldi_div_up(p->cin_store, block_n);
  int base_items = a.taps * a.m_tiles * a.n_tiles;
  // split K so that there are ~2 waves of work items but each keeps >= 8 pixel tiles
  static const char* waves_env = getenv("ALDI_WGRAD_WAVES");  // perf bisection: work-item waves the split-K aims at
  static const char* waves2_env = getenv("ALDI_WGRAD_WAVES_MH2");
  int waves = (waves_env && atoi(waves_env) >= 1) ? atoi(waves_env) : 2;
  // measured: one wave of longer items for the 256-row tile (whose epilogue cannot overlap the next main loop) is
  // slower than two (wgrad 7.36 vs 6.84 ms per step), so both tile heights aim at two waves
  if (mh == 2 && waves2_env && atoi(waves2_env) >= 1) waves = atoi(waves2_env);
  int want = aldi_div_up(waves * aldi_num_sms(), base_items);
  int max_split = a.pix_tiles / 8 > 0 ? a.pix_tiles / 8 : 1;
  a.ksplit = want < 1 ? 1 : (want > max_split ? max_split : want);
  a.num_items = base_items * a.ksplit;
  a.scale = p->scale;
  a.dw = p->dw;
  a.dbias = p->dbias;
  a.cout_store = p->cout_store;
  a.cin_store = p->cin_store;

  CUtensorMap tmY, tmX;
  {
    uint64_t dims[4] = {(uint64_t)p->dy_c, (uint64_t)p->wo, (uint64_t)p->ho, (uint64_t)p->n};
    uint64_t strides[3] = {(uint64_t)p->dy_sw * 2, (uint64_t)p->dy_sh * 2, (uint64_t)p->dy_sn * 2};
    uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, 1};
    int rc = aldi_make_tmap_bf16(&tmY, p->dy, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)p->x_c, (uint64_t)p->x_w, (uint64_t)p->x_h, (uint64_t)p->x_n};
    uint64_t strides[3] = {(uint64_t)p->x_sw * 2, (uint64_t)p->x_sh * 2, (uint64_t)p->x_sn * 2};
    uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, 1};
    int rc = aldi_make_tmap_bf16(&tmX, p->x, 4, dims, strides, box);
    if (rc) return rc;
  }
  if (mh == 2) {
    switch (block_n) {
      case 256: return launch_wgrad<256, 2>(tmY, tmX, a, stream);
      case 128: return launch_wgrad<128, 2>(tmY, tmX, a, stream);
      default: return launch_wgrad<64, 2>(tmY, tmX, a, stream);
    }
  }
  switch (block_n) {
    case 256: return launch_wgrad<256, 1>(tmY, tmX, a, stream);
    case 128: return launch_wgrad<128, 1>(tmY, tmX, a, stream);
    default: return launch_wgrad<64, 1>(tmY, tmX, a, stream);
  }
}
