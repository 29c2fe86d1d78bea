// Implicit-GEMM convolution for layers with FEW OUTPUT CHANNELS (<= 128), pixels on the MMA's N dimension.
//
// Why: on this tensor pipe a tcgen05.mma of shape 128 x N x 16 occupies the pipe for ~128 cycles whatever N is
// (measured: the 3x3 layers run at 0.27 / 0.52 / 1.0 of the bf16 peak for 64 / 128 / 256 output channels, with
// operands from L2 or from a resident shared-memory tile alike).  conv_tc.cu puts the 128 pixels of a tile on M and
// the output channels on N, so a 64- or 128-channel layer can use a quarter / half of the pipe at best.  Here the
// roles are swapped:  D[cout, pixel] = sum_k W[cout, k] * X[pixel, k]
//   A (M = 128 rows)  = weights of one (tap, 64-channel chunk) K block, 16 KB; a layer with 64 output channels reads
//                       its missing rows as TMA out-of-bounds zeros
//   B (N = 256 rows)  = a 16 x 16 pixel patch of the channels-last activation, shifted by the tap: ONE 4-D TMA box,
//                       32 KB, out-of-bounds zero fill = the convolution padding (as in conv_tc.cu)
// so every instruction is a full-rate 128 x 256 x 16.  The accumulator comes out channel-major (TMEM lane = output
// channel, column = pixel); the epilogue transposes it while staging: thread = channel reads 32 pixels with one
// tcgen05.ld, applies its channel's FrozenBN scale / shift (+ ReLU, + ReLU-backward mask from a TMA-loaded tile) and
// writes 2-byte elements into a 128-byte-swizzled [pixel][channel] tile — the 32 lanes of a warp cover 64 contiguous
// bytes of one pixel row, conflict-free — which leaves as ONE TMA store per 64 channels (clipping partial tiles).
// Persistent, one CTA per SM: warp 0 TMA producer, warp 1 MMA issuer, warps 4-11 epilogue (two per TMEM lane
// quarter, each taking half of the tile's pixels); two 256-column accumulators overlap epilogue and main loop.
//
// Used by aldi_conv_tc for 3x3 / 4x1 layers with <= 128 output channels and bf16 output (res2 / res3 bottleneck 3x3s,
// the stem); same replacement targets as conv_tc.cu.
#include "common.cuh"
#include "sm100.cuh"
#include "tmap.h"
#include "../../include/aldi_b200.h"
#include <stdlib.h>

using namespace sm100;

namespace {

constexpr int kPixTile = 256;                 // pixels per tile (N), a 16 x 16 patch
constexpr int kPatch = 16;
constexpr int kBlockK = 64;
constexpr int kWBytes = 128 * kBlockK * 2;    // A: 16 KB
constexpr int kXBytes = kPixTile * kBlockK * 2;  // B: 32 KB
constexpr int kStageBytes = kWBytes + kXBytes;
constexpr int kChunkBytes = kPixTile * 128;   // one [256 pixels][64 channels] bf16 tile: 32 KB
constexpr int kThreads = 384;
constexpr int kMaxStages = 4;
constexpr int kSmemLimit = 226 * 1024;

struct TnArgs {
  int tiles_h, tiles_w, num_tiles;
  int kchunks, taps_w, num_kb;
  int pad_h, pad_w;
  const float* scale;
  const float* bias;
  int relu, has_mask;
  int cout_chunks;     // 64-channel chunks of the output (1 or 2)
  int cout_store;
  int stages;
};

__device__ __forceinline__ uint32_t swz_elem(int row, int col) {   // byte offset of bf16 element (row, col) in a SW128 tile
  return (uint32_t)(row * 128 + ((((col >> 3) ^ (row & 7))) << 4) + ((col & 7) << 1));
}

__global__ void __launch_bounds__(kThreads, 1)
conv_tn_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmM, const TnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_out = smem + a.stages * kStageBytes;                 // cout_chunks x 32 KB
  uint8_t* s_mask = s_out + a.cout_chunks * kChunkBytes;          // cout_chunks x 32 KB (dgrad with ReLU mask)

  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t mask_full_bar, mask_empty_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmO);
    if (a.has_mask) prefetch_tmap(&tmM);
    for (int i = 0; i < kMaxStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 8); }
    mbar_init(&mask_full_bar, 1);
    mbar_init(&mask_empty_bar, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();
  pdl_launch_dependents();

  auto tile_origin = [&](int tile, int& img, int& h0, int& w0) {
    const int twi = tile % a.tiles_w;
    tile /= a.tiles_w;
    const int thi = tile % a.tiles_h;
    img = tile / a.tiles_h;
    h0 = thi * kPatch;
    w0 = twi * kPatch;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        int img, h0, w0;
        tile_origin(tile, img, h0, w0);
        for (int kb = 0; kb < a.num_kb; ++kb) {
          const int tap = kb / a.kchunks, kc = kb - tap * a.kchunks;
          const int r = tap / a.taps_w, s = tap - r * a.taps_w;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sw = smem + stage * kStageBytes;
          mbar_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_2d(sw, &tmW, &full_bar[stage], kb * kBlockK, 0);
          tma_load_4d(sw + kWBytes, &tmX, &full_bar[stage], kc * kBlockK, w0 + s - a.pad_w, h0 + r - a.pad_h, img);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // (an M = 64 instruction was timed too: same duration as M = 128, so 64-channel layers gain nothing from it)
      constexpr uint32_t idesc = make_idesc_bf16(128, kPixTile, 0, 0);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kPixTile;
        for (int kb = 0; kb < a.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sw = smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc = make_smem_desc_sw128(sw, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sw + kWBytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else if (warp == 2) {
    // ReLU-backward mask tiles (data gradients): one [256 pixels][64 channels] box per output chunk and tile
    if (lane == 0 && a.has_mask) {
      int it = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
        int img, h0, w0;
        tile_origin(tile, img, h0, w0);
        mbar_wait(&mask_empty_bar, (it & 1) ^ 1);
        mbar_expect_tx(&mask_full_bar, (uint32_t)a.cout_chunks * kChunkBytes);
        for (int c = 0; c < a.cout_chunks; ++c) tma_load_4d(s_mask + c * kChunkBytes, &tmM, &mask_full_bar, c * 64, w0, h0, img);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                 // TMEM lane quarter: output channels 32q .. 32q+31
    const int half = (warp - 4) >> 2;       // which 128 pixels of the tile
    const int ch = q * 32 + lane;
    const bool ch_ok = ch < a.cout_store;
    const float sc = (ch_ok && a.scale) ? __ldg(a.scale + ch) : 1.f;
    const float bi = (ch_ok && a.bias) ? __ldg(a.bias + ch) : 0.f;
    const bool chunk_ok = (ch >> 6) < a.cout_chunks;
    uint8_t* my_out = s_out + (ch >> 6) * kChunkBytes;
    const uint8_t* my_mask = s_mask + (ch >> 6) * kChunkBytes;
    const int col = ch & 63;
    int it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      int img, h0, w0;
      tile_origin(tile, img, h0, w0);
      mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
      tc_fence_after();
      // the previous tile's TMA stores must have finished READING the staging tile before it is overwritten
      if (warp == 4 && lane == 0) bulk_wait_group_read<0>();
      named_bar_sync(1, 256);
      if (a.has_mask) mbar_wait(&mask_full_bar, it & 1);
#pragma unroll 1
      for (int c4 = 0; c4 < 4; ++c4) {
        const int n0 = half * 128 + c4 * 32;
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kPixTile + n0), raw);
        tmem_ld_wait();
        if (chunk_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = fmaf(__uint_as_float(raw[j]), sc, bi);
            if (a.relu) v = fmaxf(v, 0.f);
            const uint32_t off = swz_elem(n0 + j, col);
            if (a.has_mask) {
              const float m = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(my_mask + off));
              v = m > 0.f ? v : 0.f;
            }
            *reinterpret_cast<__nv_bfloat16*>(my_out + off) = __float2bfloat16_rn(ch_ok ? v : 0.f);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tmem_empty_bar[acc]);
        if (a.has_mask) mbar_arrive(&mask_empty_bar);
      }
      named_bar_sync(2, 256);
      if (warp == 4 && lane == 0) {
        for (int c = 0; c < a.cout_chunks; ++c) tma_store_4d(&tmO, s_out + c * kChunkBytes, c * 64, w0, h0, img);
        bulk_commit_group();
      }
    }
    if (warp == 4 && lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int make_cl_tmap(CUtensorMap* tm, const void* base, int c, int w, int h, int n, long long sw, long long sh, long long sn) {
  uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
  uint64_t strides[3] = {(uint64_t)sw * 2, (uint64_t)sh * 2, (uint64_t)sn * 2};
  uint32_t box[4] = {64, kPatch, kPatch, 1};
  return aldi_make_tmap_bf16(tm, base, 4, dims, strides, box);
}

}  // namespace

// Returns ALDI_OK with *handled = 1 when the layer was launched here, *handled = 0 when conv_tc.cu should take it.
int aldi_conv_tn_try(const aldi_conv_params* p, cudaStream_t stream, int* handled) {
  *handled = 0;
  static const char* off = getenv("ALDI_NO_CONV_TN");
  if (off && off[0] == '1') return ALDI_OK;
  const int taps = p->taps_h * p->taps_w;
  // 64 output channels: half of every M = 128 instruction would be zero padding (measured 370 TFLOP/s effective against
  // 430-460 for conv_tc's halo mode), so those layers stay in conv_tc.cu unless asked for (ALDI_CONV_TN_C64=1)
  static const char* c64 = getenv("ALDI_CONV_TN_C64");
  static const char* stem_env = getenv("ALDI_CONV_TN_STEM");
  const bool stem_like = p->taps_h == 4 && p->taps_w == 1 && stem_env && stem_env[0] == '1';
  if (p->cout_p < 128 && !(c64 && c64[0] == '1') && !stem_like) return ALDI_OK;
  if (p->out_dtype != ALDI_DTYPE_BF16 || p->cout_p > 128 || p->cout_store != p->cout_p || taps < 4 || p->res_mode ||
      p->accumulate || p->stride > 1 || p->x_c % 64 != 0)
    return ALDI_OK;
  if ((p->out_sw % 8) || (p->out_sh % 8) || (p->out_sn % 8) || (reinterpret_cast<uintptr_t>(p->out) & 15)) return ALDI_OK;
  if (p->mask && ((p->mask_sw % 8) || (p->mask_sh % 8) || (p->mask_sn % 8) || (reinterpret_cast<uintptr_t>(p->mask) & 15)))
    return ALDI_OK;
  TnArgs a;
  a.tiles_h = aldi_div_up(p->ho, kPatch);
  a.tiles_w = aldi_div_up(p->wo, kPatch);
  a.num_tiles = p->n * a.tiles_h * a.tiles_w;
  a.kchunks = p->x_c / 64;
  a.taps_w = p->taps_w;
  a.num_kb = taps * a.kchunks;
  a.pad_h = p->pad_h; a.pad_w = p->pad_w;
  a.scale = p->scale; a.bias = p->bias;
  a.relu = p->relu;
  a.has_mask = p->mask ? 1 : 0;
  a.cout_chunks = p->cout_p / 64;
  a.cout_store = p->cout_store;
  const int epi = a.cout_chunks * kChunkBytes * (a.has_mask ? 2 : 1);
  int stages = (kSmemLimit - 1024 - epi) / kStageBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return ALDI_OK;
  a.stages = stages;
  const int smem_bytes = stages * kStageBytes + epi + 1024;

  CUtensorMap tmW, tmX, tmO, tmM;
  {
    const uint64_t ktot = (uint64_t)taps * p->x_c;
    uint64_t dims[2] = {ktot, (uint64_t)p->cout_p};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {64, 128};   // rows beyond cout_p are out of bounds -> zeros
    int rc = aldi_make_tmap_bf16(&tmW, p->w, 2, dims, strides, box);
    if (rc) return rc;
  }
  int rc = make_cl_tmap(&tmX, p->x, p->x_c, p->x_w, p->x_h, p->x_n, p->x_sw, p->x_sh, p->x_sn);
  if (rc) return rc;
  rc = make_cl_tmap(&tmO, p->out, p->cout_store, p->wo, p->ho, p->n, p->out_sw, p->out_sh, p->out_sn);
  if (rc) return rc;
  tmM = tmO;
  if (p->mask) {
    rc = make_cl_tmap(&tmM, p->mask, p->cout_store, p->wo, p->ho, p->n, p->mask_sw, p->mask_sh, p->mask_sn);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) {
      aldi_set_error("aldi_conv_tc(tn): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ALDI_ERR_CUDA;
    }
    attr_set = true;
  }
  const int grid = a.num_tiles < aldi_num_sms() ? a.num_tiles : aldi_num_sms();
  cudaError_t le = aldi_launch_pdl(conv_tn_kernel, dim3(grid), dim3(kThreads), (size_t)smem_bytes, stream, tmW, tmX, tmO, tmM, a);
  ALDI_COUNT_LAUNCH();
  if (le != cudaSuccess) {
    aldi_set_error("aldi_conv_tc(tn): launch failed: %s", cudaGetErrorString(le));
    return ALDI_ERR_CUDA;
  }
  ALDI_CUDA_LAUNCH_CHECK("aldi_conv_tc(tn)");
  *handled = 1;
  return ALDI_OK;
}
