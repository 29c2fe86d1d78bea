// Implicit-GEMM convolution / linear layer on Blackwell tensor cores (tcgen05 + TMEM + TMA), bf16 -> fp32.
//
// GEMM view:  D[pixel, cout] = sum_{tap, cin} X[pixel shifted by tap, cin] * W[cout, tap, cin]
//   M tile = 128 output pixels arranged as a TH x TW spatial patch of ONE image (TH*TW == 128)
//   N tile = BLOCK_N output channels, K block = 64 input channels of one filter tap.
// The A operand of every (tap, cin-chunk) K block is ONE 4-D TMA box {64 ch, TW, TH, 1} of the
// channels-last activation view, shifted by the tap; TMA zero-fills out-of-bounds coordinates, which
// is exactly the convolution's zero padding, so no im2col buffer and no predication exist anywhere.
// The box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle = the canonical
// K-major UMMA operand layout.  Warp roles (persistent CTA, one per SM):
//   warp 0: TMA producer (A/B stages)      warp 1: TMEM alloc + tcgen05.mma issuer
//   warp 2: TMA producer of the epilogue operands (residual / mask tiles)       warp 3: idle
//   warps 4-11: epilogue — two warps per TMEM lane quarter (= per scheduler).  They take ALTERNATE 64-channel
//   chunks and never synchronise with each other: each warp reads its 32 pixel rows from TMEM (tcgen05.ld is
//   ~64 B/clk/SM plus a ~100-cycle wait::ld, the real floor of the K-small layers), stages a 32 x 64 bf16 slab
//   in its own 4 KB of swizzled shared memory and stores it with its own TMA store, while its sibling on the
//   same scheduler covers the latencies with the neighbouring chunk
// Two TMEM accumulator buffers let the epilogue of tile i overlap the main loop of tile i+1.
//
// Epilogue (bf16 outputs): most ResNet layers here are HBM-bound (1x1 convs with a residual), so the
// epilogue is built like a copy kernel: residual / ReLU-mask tiles are TMA-loaded (warp 6, double buffered,
// 64 channels = 128-byte rows at a time) while the main loop runs, the finished 128 x 64 bf16 chunk is staged
// in 128B-swizzled shared memory and leaves with ONE TMA store, which also clips partial tiles.  `accumulate`
// is the same path with the output tile as the residual.  fp32 outputs (RPN head / box predictor rows,
// 15 / 41 channels) keep direct per-thread stores.
//
// Replaces the cuDNN/cuBLAS calls detectron2 makes for ResNet/FPN/RPN/box-head layers
// (reached from aldi/trainer.py:87, aldi/distill.py:157,162, aldi/pseudolabeler.py:21).
#include "common.cuh"
#include "sm100.cuh"
#include "tmap.h"
#include "../../include/aldi_b200.h"
#include <stdlib.h>

using namespace sm100;

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                      // bf16 elements = 128 B = swizzle span
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kNumThreads = 384;
constexpr int kEpiWarps = 8;

constexpr int kEpiChunk = 64;                    // channels per epilogue chunk (128-byte rows)
constexpr int kEpiBytes = kBlockM * kEpiChunk * 2;  // 16 KB
constexpr int kMaxStages = 8;
constexpr int kMaxEpiBufs = 8;
constexpr int kHaloW = 16;                         // halo row pitch in pixels: 8-row swizzle groups stay 1024-aligned
constexpr int kHaloH = 18;                         // 16 output rows + 2
constexpr int kHaloBytes = kHaloH * kHaloW * 128;  // 36 KB per 64-channel chunk
constexpr int kMaxHaloStages = 4;
constexpr int kSmemLimit = 226 * 1024;     // dynamic budget (static barriers etc. live in the remaining 1 KB)

struct ConvArgs {
  int n, ho, wo;
  int tiles_h, tiles_w, th, tw, tw_shift;
  int num_n_tiles, num_tiles;
  int kchunks, taps_w, num_kb;
  int pad_h, pad_w;
  const float* scale;
  const float* bias;
  const __nv_bfloat16* residual;
  int res_mode;
  long long res_sw, res_sh, res_sn;
  const __nv_bfloat16* mask;
  long long mask_sw, mask_sh, mask_sn;
  void* out;
  int out_f32;
  int cout_store;
  long long out_sw, out_sh, out_sn;
  int relu, accumulate;
  int stages;        // A/B pipeline depth (runtime: the epilogue buffers share the 227 KB)
  int tma_epi;       // bf16 output through shared memory + TMA store
  int epi_res, epi_mask;  // residual / mask tiles arrive by TMA (tma_epi only)
  int epi_bufs;      // depth of the residual / mask tile ring (2..4): bytes in flight for the HBM-bound layers
  int halo;          // 3x3 small-channel mode: the A operand of all 9 taps comes from ONE halo tile per 64-channel chunk
  int halo_stages;   // depth of the halo ring (each 18 x 16 pixels x 128 B = 36 KB)
  int halo_bo;       // descriptor base-offset convention (1: tap column shift, 2: zero) — bring-up knob
  int b_resident;    // halo mode, one N tile: all 9 x kchunks weight blocks stay in shared memory for the whole kernel
  int epi_prefetch;  // tiles ahead whose residual / mask boxes warp 2 prefetches into the L2 (0 = off)
  int a_prefetch;    // K blocks ahead of the smem ring whose A boxes warp 0 prefetches into the L2 (0 = off)
  int debug;         // ALDI_CONV_DEBUG (perf bisection only): 1 no TMA store, 2 also no smem staging, 3 empty epilogue
};

template <int BLOCK_N>
struct Cfg {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
};

// (x0, x1) = (x0, x1) * (s0, s1) + (b0, b1) in one packed FFMA2 (sm_100 fma.rn.f32x2)
__device__ __forceinline__ void ffma2(float& x0, float& x1, float s0, float s1, float b0, float b1) {
  unsigned long long x, sc, bi;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(sc) : "f"(s0), "f"(s1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bi) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(sc), "l"(bi));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
}

__device__ __forceinline__ uint32_t swz128(int row, int chunk16) {  // byte offset inside a 128B-swizzled tile
  return (uint32_t)(row * 128 + ((chunk16 ^ (row & 7)) << 4));
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& q, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float* f) {
  uint4 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return q;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmM, const ConvArgs a) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // [stages x (A | B)] [out x2] [res x2] [mask x2]   (all 16 KB pieces, 1024-byte aligned)
  // halo mode: [halo_stages x 36 KB] [stages x B] [out] [res] [mask]
  uint8_t* s_bring = smem + (a.halo ? a.halo_stages * kHaloBytes : 0);
  uint8_t* s_out = a.halo ? s_bring + a.stages * C::kBBytes : smem + a.stages * C::kStageBytes;
  uint8_t* s_res = s_out + 2 * kEpiBytes;
  uint8_t* s_mask = s_res + (a.epi_res ? a.epi_bufs * kEpiBytes : 0);

  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t halo_full_bar[kMaxHaloStages];
  __shared__ __align__(8) uint64_t halo_empty_bar[kMaxHaloStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t epi_full_bar[kMaxEpiBufs];
  __shared__ __align__(8) uint64_t epi_empty_bar[kMaxEpiBufs];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool epi_loads = a.epi_res || a.epi_mask;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (a.tma_epi) prefetch_tmap(&tmO);
    if (a.epi_res) prefetch_tmap(&tmR);
    if (a.epi_mask) prefetch_tmap(&tmM);
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kMaxHaloStages; ++i) {
      mbar_init(&halo_full_bar[i], 1);
      mbar_init(&halo_empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      // one arrive per epilogue warp that reads the tile (N = 64 TMA path: one chunk per tile -> one warp per quarter)
      mbar_init(&tmem_empty_bar[i], (a.tma_epi && BLOCK_N == 64) ? 4 : kEpiWarps);
    }
    for (int i = 0; i < kMaxEpiBufs; ++i) {
      mbar_init(&epi_full_bar[i], 1);
      mbar_init(&epi_empty_bar[i], 4);  // a chunk is consumed by one warp per lane quarter
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // everything above overlapped the previous kernel's tail (programmatic dependent launch); its results are
  // visible after this wait, and the next kernel may begin ITS prologue as soon as our CTAs retire
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0, hstage = 0;
      uint32_t phase = 0, hphase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int n_tile = tile % a.num_n_tiles;
        int m_tile = tile / a.num_n_tiles;
        const int twi = m_tile % a.tiles_w;
        m_tile /= a.tiles_w;
        const int thi = m_tile % a.tiles_h;
        const int img = m_tile / a.tiles_h;
        const int h0 = thi * a.th, w0 = twi * a.tw;
        if (a.halo) {
          // one halo tile per 64-channel chunk feeds all nine taps; the weights stream through the B ring, or are
          // loaded once when all of them fit (a stage of 8 KB is 128 cycles of MMA work: no ring hides TMA latency)
          if (a.b_resident && tile == (int)blockIdx.x) {
            mbar_expect_tx(&full_bar[0], (uint32_t)(9 * a.kchunks) * C::kBBytes);
            for (int i = 0; i < 9 * a.kchunks; ++i)
              tma_load_2d(s_bring + i * C::kBBytes, &tmB, &full_bar[0], i * kBlockK, n_tile * BLOCK_N);
          }
          for (int kc = 0; kc < a.kchunks; ++kc) {
            mbar_wait(&halo_empty_bar[hstage], hphase ^ 1);
            mbar_expect_tx(&halo_full_bar[hstage], kHaloBytes);
            tma_load_4d(smem + hstage * kHaloBytes, &tmA, &halo_full_bar[hstage], kc * kBlockK, w0 - a.pad_w, h0 - a.pad_h, img);
            if (++hstage == a.halo_stages) { hstage = 0; hphase ^= 1; }
            for (int tap = 0; tap < 9 && !a.b_resident; ++tap) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], C::kBBytes);
              tma_load_2d(s_bring + stage * C::kBBytes, &tmB, &full_bar[stage], (tap * a.kchunks + kc) * kBlockK, n_tile * BLOCK_N);
              if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
          }
          continue;
        }
        if (a.a_prefetch) {
          // 1x1 layers stream x straight from HBM: warm the L2 with the NEXT tile's A boxes while this tile runs
          const int nt = tile + gridDim.x;
          if (nt < a.num_tiles) {
            int mt = nt / a.num_n_tiles;
            const int pw = mt % a.tiles_w;
            mt /= a.tiles_w;
            const int ph = mt % a.tiles_h;
            const int pi = mt / a.tiles_h;
            for (int kc = 0; kc < a.kchunks; ++kc)
              tma_prefetch_4d(&tmA, kc * kBlockK, pw * a.tw - a.pad_w, ph * a.th - a.pad_h, pi);
          }
        }
        for (int kb = 0; kb < a.num_kb; ++kb) {
          const int tap = kb / a.kchunks, kc = kb - tap * a.kchunks;
          const int r = tap / a.taps_w, s = tap - r * a.taps_w;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          tma_load_4d(sa, &tmA, &full_bar[stage], kc * kBlockK, w0 + s - a.pad_w, h0 + r - a.pad_h, img);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * kBlockK, n_tile * BLOCK_N);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
      int stage = 0, hstage = 0;
      uint32_t phase = 0, hphase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        if (a.halo) {
          if (a.b_resident && it == 0) {
            mbar_wait(&full_bar[0], 0);
            tc_fence_after();
          }
          for (int kc = 0; kc < a.kchunks; ++kc) {
            mbar_wait(&halo_full_bar[hstage], hphase);
            tc_fence_after();
            const uint32_t hbase = smem_u32(smem + hstage * kHaloBytes);
            for (int tap = 0; tap < 9; ++tap) {
              const int r = tap / 3, sx = tap - r * 3;
              if (a.b_resident) {
                uint64_t adesc = make_smem_desc_sw128(hbase + (uint32_t)((r * kHaloW + sx) * 128), 16, kHaloW * 128);
                const uint64_t bdesc =
                    make_smem_desc_sw128(smem_u32(s_bring + (tap * a.kchunks + kc) * C::kBBytes), 16, 1024);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | tap | k) != 0);
                continue;
              }
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              // rows of the shifted operand: pixel (ty + r, tx + sx) of the halo; 8 pixels of one tile row are one
              // swizzle group (stride 128 B), the next tile row is kHaloW pixels = 2048 B further
              uint64_t adesc = make_smem_desc_sw128(hbase + (uint32_t)((r * kHaloW + sx) * 128), 16, kHaloW * 128);
              if (a.halo_bo == 1) adesc |= (uint64_t)(sx & 7) << 49;
              const uint64_t bdesc = make_smem_desc_sw128(smem_u32(s_bring + stage * C::kBBytes), 16, 1024);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | tap | k) != 0);
              umma_commit(&empty_bar[stage]);
              if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(&halo_empty_bar[hstage]);
            if (++hstage == a.halo_stages) { hstage = 0; hphase ^= 1; }
          }
          umma_commit(&tmem_full_bar[acc]);
          continue;
        }
        for (int kb = 0; kb < a.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sa + kABytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 (32 B) along K inside the 128-B swizzle atom: +2 in the (addr>>4) field
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp == 2) {
    // ===================== TMA producer of the epilogue operands (residual / mask tiles) =====================
    if (lane == 0 && epi_loads) {
      int buf = 0;
      uint32_t phase = 0;
      const uint32_t res_bytes = a.res_mode == 2 ? kEpiBytes / 4 : kEpiBytes;
      const uint32_t bytes = (a.epi_res ? res_bytes : 0u) + (a.epi_mask ? (uint32_t)kEpiBytes : 0u);
      // the residual / mask tiles are cold in HBM and the ring is only 2-4 chunks deep: warm the L2 `epi_prefetch`
      // tiles ahead so the ring's loads see L2 latency instead of DRAM latency
      auto prefetch_tile = [&](int t) {
        if (t >= a.num_tiles) return;
        const int nt = t % a.num_n_tiles;
        int mt = t / a.num_n_tiles;
        const int pw = mt % a.tiles_w;
        mt /= a.tiles_w;
        const int ph = mt % a.tiles_h;
        const int pi = mt / a.tiles_h;
        for (int c0 = 0; c0 < BLOCK_N; c0 += kEpiChunk) {
          const int cb = nt * BLOCK_N + c0;
          if (a.epi_res) {
            if (a.res_mode == 2) tma_prefetch_4d(&tmR, cb, (pw * a.tw) >> 1, (ph * a.th) >> 1, pi);
            else tma_prefetch_4d(&tmR, cb, pw * a.tw, ph * a.th, pi);
          }
          if (a.epi_mask) tma_prefetch_4d(&tmM, cb, pw * a.tw, ph * a.th, pi);
        }
      };
      for (int d = 1; d < a.epi_prefetch; ++d) prefetch_tile(blockIdx.x + d * gridDim.x);
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        if (a.epi_prefetch) prefetch_tile(tile + a.epi_prefetch * gridDim.x);
        const int n_tile = tile % a.num_n_tiles;
        int m_tile = tile / a.num_n_tiles;
        const int twi = m_tile % a.tiles_w;
        m_tile /= a.tiles_w;
        const int thi = m_tile % a.tiles_h;
        const int img = m_tile / a.tiles_h;
        const int h0 = thi * a.th, w0 = twi * a.tw;
        for (int c0 = 0; c0 < BLOCK_N; c0 += kEpiChunk) {
          const int cbase = n_tile * BLOCK_N + c0;
          mbar_wait(&epi_empty_bar[buf], phase ^ 1);
          mbar_expect_tx(&epi_full_bar[buf], bytes);
          if (a.epi_res) {
            if (a.res_mode == 2)
              tma_load_4d(s_res + buf * kEpiBytes, &tmR, &epi_full_bar[buf], cbase, w0 >> 1, h0 >> 1, img);
            else
              tma_load_4d(s_res + buf * kEpiBytes, &tmR, &epi_full_bar[buf], cbase, w0, h0, img);
          }
          if (a.epi_mask) tma_load_4d(s_mask + buf * kEpiBytes, &tmM, &epi_full_bar[buf], cbase, w0, h0, img);
          if (++buf == a.epi_bufs) { buf = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps (4..11) =====================
    const int q = warp & 3;              // TMEM lane quarter this warp may touch (== its scheduler)
    const int par = (warp - 4) >> 2;     // which of the quarter's two warps: takes chunks with (global index & 1) == par
    const int colhalf = par;             // fp32 direct path: alternate 32-column groups
    const int row = q * 32 + lane;
    const int hl = row >> a.tw_shift, wl = row & (a.tw - 1);
    // residual row of this thread inside the TMA-loaded tile (res_mode 2: the (TH/2 x TW/2) coarse tile)
    const int rrow = (a.res_mode == 2) ? ((hl >> 1) * (a.tw >> 1) + (wl >> 1)) : row;
    // this warp's 32-row slab inside the tile and its private 4 KB staging buffer
    const int slab_h = (q * 32) >> a.tw_shift, slab_w = (q * 32) & (a.tw - 1);
    uint8_t* my_out = s_out + (warp - 4) * 4096;
    constexpr int kChunks = BLOCK_N / kEpiChunk;
    int it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n_tile = tile % a.num_n_tiles;
      int m_tile = tile / a.num_n_tiles;
      const int twi = m_tile % a.tiles_w;
      m_tile /= a.tiles_w;
      const int thi = m_tile % a.tiles_h;
      const int img = m_tile / a.tiles_h;
      const int h0 = thi * a.th, w0 = twi * a.tw;
      const int h = h0 + hl, w = w0 + wl;

      if (a.tma_epi) {
        // ---------- bf16 output: per-warp 32 x 64 slabs through swizzled shared memory, TMA in / TMA out ----------
        bool waited = false;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          const int g = it * kChunks + c;  // global chunk index: residual/mask ring slot and warp assignment
          if ((g & 1) != par) continue;
          if (!waited) {
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            waited = true;
          }
          const int cbase = n_tile * BLOCK_N + c * kEpiChunk;
          const int slot = g % a.epi_bufs;
          const uint32_t ephase = (uint32_t)(g / a.epi_bufs) & 1u;
          if (a.debug == 3) continue;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int cw = cbase + half * 32;  // first channel of this half
            uint32_t raw[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c * kEpiChunk + half * 32),
                          raw);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
            if (a.scale && a.bias) {  // FrozenBN: v = v * scale + shift, two channels per FFMA2
              const float4* sp = reinterpret_cast<const float4*>(a.scale + cw);
              const float4* bp = reinterpret_cast<const float4*>(a.bias + cw);
#pragma unroll
              for (int gq = 0; gq < 8; ++gq) {
                const float4 s4 = __ldg(sp + gq), b4 = __ldg(bp + gq);
                ffma2(v[4 * gq], v[4 * gq + 1], s4.x, s4.y, b4.x, b4.y);
                ffma2(v[4 * gq + 2], v[4 * gq + 3], s4.z, s4.w, b4.z, b4.w);
              }
            } else if (a.bias) {
              const float4* bp = reinterpret_cast<const float4*>(a.bias + cw);
#pragma unroll
              for (int gq = 0; gq < 8; ++gq) {
                const float4 t = __ldg(bp + gq);
                v[4 * gq] += t.x; v[4 * gq + 1] += t.y; v[4 * gq + 2] += t.z; v[4 * gq + 3] += t.w;
              }
            } else if (a.scale) {
              const float4* sp = reinterpret_cast<const float4*>(a.scale + cw);
#pragma unroll
              for (int gq = 0; gq < 8; ++gq) {
                const float4 t = __ldg(sp + gq);
                v[4 * gq] *= t.x; v[4 * gq + 1] *= t.y; v[4 * gq + 2] *= t.z; v[4 * gq + 3] *= t.w;
              }
            }
            if (half == 0 && epi_loads) mbar_wait(&epi_full_bar[slot], ephase);
            if (a.epi_res && !a.accumulate) {
              const uint8_t* rb = s_res + slot * kEpiBytes;
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                float f[8];
                unpack_bf16x8(*reinterpret_cast<const uint4*>(rb + swz128(rrow, half * 4 + gq)), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[gq * 8 + j] += f[j];
              }
            }
            uint4 packed[4];
            if (!a.accumulate) {
              // ReLU and the ReLU-backward mask commute with the bf16 rounding: do them two channels at a time
              __nv_bfloat162* pk = reinterpret_cast<__nv_bfloat162*>(packed);
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              if (a.relu) {
                const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[j] = __hmax2(pk[j], z);
              }
              if (a.epi_mask) {
                const uint8_t* mb = s_mask + slot * kEpiBytes;
                const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
                uint32_t* pw = reinterpret_cast<uint32_t*>(packed);
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                  const uint4 m4 = *reinterpret_cast<const uint4*>(mb + swz128(row, half * 4 + gq));
                  const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&m4);
#pragma unroll
                  for (int j = 0; j < 4; ++j) pw[gq * 4 + j] &= __hgt2_mask(mh[j], z);
                }
              }
            } else {
              if (a.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
              }
              if (a.epi_mask) {
                const uint8_t* mb = s_mask + slot * kEpiBytes;
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                  float f[8];
                  unpack_bf16x8(*reinterpret_cast<const uint4*>(mb + swz128(row, half * 4 + gq)), f);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[gq * 8 + j] = (f[j] > 0.f) ? v[gq * 8 + j] : 0.f;
                }
              }
              // out += v: the old output tile arrived as the "residual"
              const uint8_t* rb = s_res + slot * kEpiBytes;
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                float f[8];
                unpack_bf16x8(*reinterpret_cast<const uint4*>(rb + swz128(row, half * 4 + gq)), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[gq * 8 + j] += f[j];
                packed[gq] = pack_bf16x8(v + gq * 8);
              }
            }
            if (a.debug == 2) {
              if (packed[0].x == 0x12345678u && packed[3].w == 0x9abcdef0u) atomicAdd(reinterpret_cast<int*>(a.out), 1);
              continue;
            }
            if (half == 0) {
              // this warp's previous TMA store must have finished reading the slab buffer
              if (lane == 0) bulk_wait_group_read<0>();
              __syncwarp();
            }
#pragma unroll
            for (int gq = 0; gq < 4; ++gq)
              *reinterpret_cast<uint4*>(my_out + swz128(lane, half * 4 + gq)) = packed[gq];
          }
          if (epi_loads) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&epi_empty_bar[slot]);
          }
          if (a.debug == 2) continue;
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && a.debug != 1) {
            tma_store_4d(&tmO, my_out, cbase, w0 + slab_w, h0 + slab_h, img);
            bulk_commit_group();
          }
        }
        if (waited) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        continue;
      } else {
        // ---------- fp32 (or unaligned) output: direct per-thread row-segment stores ----------
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        const bool valid = (h < a.ho) && (w < a.wo);
        const long long out_off = (long long)img * a.out_sn + (long long)h * a.out_sh + (long long)w * a.out_sw;
        long long res_off = 0, mask_off = 0;
        if (a.res_mode == 1)
          res_off = (long long)img * a.res_sn + (long long)h * a.res_sh + (long long)w * a.res_sw;
        else if (a.res_mode == 2)
          res_off = (long long)img * a.res_sn + (long long)(h >> 1) * a.res_sh + (long long)(w >> 1) * a.res_sw;
        if (a.mask) mask_off = (long long)img * a.mask_sn + (long long)h * a.mask_sh + (long long)w * a.mask_sw;
#pragma unroll 1
        for (int c0 = colhalf * 32; c0 < BLOCK_N; c0 += 64) {
          uint32_t raw[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0), raw);
          tmem_ld_wait();
          const int cbase = n_tile * BLOCK_N + c0;
          if (valid && cbase < a.cout_store) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
            if (a.scale) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= __ldg(a.scale + cbase + j);
            }
            if (a.bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += __ldg(a.bias + cbase + j);
            }
            if (a.res_mode) {
              const uint4* rp = reinterpret_cast<const uint4*>(a.residual + res_off + cbase);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float f[8];
                unpack_bf16x8(__ldg(rp + g), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[g * 8 + j] += f[j];
              }
            }
            if (a.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (a.mask) {
              const uint4* mp = reinterpret_cast<const uint4*>(a.mask + mask_off + cbase);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float f[8];
                unpack_bf16x8(__ldg(mp + g), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[g * 8 + j] = (f[j] > 0.f) ? v[g * 8 + j] : 0.f;
              }
            }
            if (a.out_f32) {
              float* op = reinterpret_cast<float*>(a.out) + out_off + cbase;
              if (cbase + 32 <= a.cout_store && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                  float4 o = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                  if (a.accumulate) {
                    float4 old = reinterpret_cast<float4*>(op)[g];
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                  }
                  reinterpret_cast<float4*>(op)[g] = o;
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (cbase + j < a.cout_store) op[j] = a.accumulate ? op[j] + v[j] : v[j];
              }
            } else {
              __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + out_off + cbase;
              if (cbase + 32 <= a.cout_store) {
                uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  if (a.accumulate) {
                    float f[8];
                    unpack_bf16x8(o4[g], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[g * 8 + j] += f[j];
                  }
                  o4[g] = pack_bf16x8(v + g * 8);
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (cbase + j < a.cout_store) {
                    float o = v[j];
                    if (a.accumulate) o += __bfloat162float(op[j]);
                    op[j] = __float2bfloat16_rn(o);
                  }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
    if (a.tma_epi && lane == 0) bulk_wait_group<0>();  // all output slabs written before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N>
int launch_conv(const CUtensorMap* tm, const ConvArgs& a, int smem_bytes, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemLimit);
    if (e != cudaSuccess) {
      aldi_set_error("aldi_conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ALDI_ERR_CUDA;
    }
    attr_set = true;
  }
  int grid = a.num_tiles < aldi_num_sms() ? a.num_tiles : aldi_num_sms();
  cudaError_t le = aldi_launch_pdl(conv_tc_kernel<BLOCK_N>, dim3(grid), dim3(kNumThreads), (size_t)smem_bytes, stream,
                                   tm[0], tm[1], tm[2], tm[3], tm[4], a);
  ALDI_COUNT_LAUNCH();
  if (le != cudaSuccess) {
    aldi_set_error("aldi_conv_tc: launch failed: %s", cudaGetErrorString(le));
    return ALDI_ERR_CUDA;
  }
  ALDI_CUDA_LAUNCH_CHECK("aldi_conv_tc");
  return ALDI_OK;
}

// spatial patch (TH x TW = 128) that wastes the fewest pixels for an (ho, wo) output
void pick_patch(int ho, int wo, int max_tw, int* th, int* tw) {
  long best = -1;
  for (int t = max_tw; t >= 8; t >>= 1) {
    int hh = 128 / t;
    long cover = (long)((wo + t - 1) / t) * t * (long)((ho + hh - 1) / hh) * hh;
    if (best < 0 || cover < best) {
      best = cover;
      *tw = t;
      *th = hh;
    }
  }
}

// 4-D channels-last tensor map {c, w, h, n} over a strided view, box {64, bw, bh, 1}
int make_cl_tmap(CUtensorMap* tm, const void* base, int c, int w, int h, int n, long long sw, long long sh,
                 long long sn, int bw, int bh) {
  uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
  uint64_t strides[3] = {(uint64_t)sw * 2, (uint64_t)sh * 2, (uint64_t)sn * 2};
  uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
  return aldi_make_tmap_bf16(tm, base, 4, dims, strides, box);
}

}  // namespace

int aldi_conv_tn_try(const aldi_conv_params* p, cudaStream_t stream, int* handled);   // conv_tn.cu

extern "C" int aldi_conv_tc(const aldi_conv_params* p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(p && p->x && p->w && p->out, "aldi_conv_tc: null pointer");
  ALDI_CHECK_ARG(p->x_c > 0 && p->x_c % 64 == 0, "aldi_conv_tc: x_c=%d must be a positive multiple of 64", p->x_c);
  ALDI_CHECK_ARG(p->cout_p > 0 && p->cout_p % 64 == 0, "aldi_conv_tc: cout_p=%d must be a multiple of 64", p->cout_p);
  ALDI_CHECK_ARG(p->cout_store > 0 && p->cout_store <= p->cout_p, "aldi_conv_tc: bad cout_store");
  ALDI_CHECK_ARG(p->n > 0 && p->ho > 0 && p->wo > 0, "aldi_conv_tc: empty output");
  ALDI_CHECK_ARG(p->taps_h > 0 && p->taps_w > 0 && p->taps_h * p->taps_w <= 64, "aldi_conv_tc: bad taps");
  ALDI_CHECK_ARG(p->stride <= 1, "aldi_conv_tc: stride %d unsupported, pass a strided view of x", p->stride);
  ALDI_CHECK_ARG((p->x_sw % 8) == 0 && (p->x_sh % 8) == 0 && (p->x_sn % 8) == 0,
                 "aldi_conv_tc: activation strides must be multiples of 8 elements (16 B)");
  ALDI_CHECK_ARG((reinterpret_cast<uintptr_t>(p->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w) & 15) == 0,
                 "aldi_conv_tc: x/w must be 16-byte aligned");
  if (p->out_dtype == ALDI_DTYPE_BF16) {
    ALDI_CHECK_ARG((p->out_sw % 8) == 0 && (p->out_sh % 8) == 0 && (p->out_sn % 8) == 0 &&
                       (reinterpret_cast<uintptr_t>(p->out) & 15) == 0,
                   "aldi_conv_tc: bf16 output must be 16-byte aligned with strides multiple of 8");
  }
  if (p->res_mode) {
    ALDI_CHECK_ARG(p->residual && (p->res_sw % 8) == 0 && (p->res_sh % 8) == 0 && (p->res_sn % 8) == 0 &&
                       (reinterpret_cast<uintptr_t>(p->residual) & 15) == 0,
                   "aldi_conv_tc: residual must be 16-byte aligned with strides multiple of 8");
  }
  if (p->mask) {
    ALDI_CHECK_ARG((p->mask_sw % 8) == 0 && (p->mask_sh % 8) == 0 && (p->mask_sn % 8) == 0 &&
                       (reinterpret_cast<uintptr_t>(p->mask) & 15) == 0,
                   "aldi_conv_tc: mask must be 16-byte aligned with strides multiple of 8");
  }

  // bf16 outputs whose channel extent is a whole number of 64-channel chunks leave through TMA; `accumulate`
  // becomes "residual = the output tile itself" (never combined with another residual by the callers)
  {
    // few output channels, several taps: pixels go on the MMA's N dimension (conv_tn.cu)
    int handled = 0;
    const int rc = aldi_conv_tn_try(p, stream, &handled);
    if (rc || handled) return rc;
  }
  const bool tma_epi = p->out_dtype == ALDI_DTYPE_BF16 && p->cout_store == p->cout_p && !(p->accumulate && p->res_mode);
  const bool epi_res = tma_epi && (p->res_mode || p->accumulate);
  const bool epi_mask = tma_epi && p->mask;
  const int res_mode = (tma_epi && p->accumulate) ? 1 : p->res_mode;

  int block_n = (p->cout_p % 256 == 0) ? 256 : (p->cout_p % 128 == 0) ? 128 : 64;
  static const char* force_bn = getenv("ALDI_CONV_BN");  // perf bisection only: cap the N tile
  if (force_bn && atoi(force_bn) >= 64 && atoi(force_bn) < block_n && p->cout_p % atoi(force_bn) == 0) block_n = atoi(force_bn);
  // shared-memory plan: [stages x (A | B)] [2 output chunks] [epi_bufs x (residual chunk, mask chunk)].
  // Keep >= 3 pipeline stages (narrowing the N tile if needed), then deepen the residual/mask ring up to 4 —
  // for the HBM-bound 1x1 layers the bytes in flight per SM are what sets the achieved bandwidth.
  const int num_kb = p->taps_h * p->taps_w * (p->x_c / 64);
  const int set_bytes = ((epi_res ? 1 : 0) + (epi_mask ? 1 : 0)) * kEpiBytes;
  const int out_bytes = tma_epi ? 2 * kEpiBytes : 0;
  while (block_n > 64 && (kSmemLimit - 1024 - out_bytes - 2 * set_bytes) / (kABytes + block_n * kBlockK * 2) < 3) block_n >>= 1;
  const int stage_bytes = kABytes + block_n * kBlockK * 2;
  const int min_stages = num_kb >= 4 ? 4 : 3;
  int epi_bufs = 2;
  static const char* force_bufs2 = getenv("ALDI_EPI_BUFS2");
  while (!force_bufs2 && set_bytes && epi_bufs < kMaxEpiBufs &&
         (kSmemLimit - 1024 - out_bytes - (epi_bufs + 1) * set_bytes) / stage_bytes >= min_stages)
    ++epi_bufs;
  const int epi_bytes = out_bytes + (set_bytes ? epi_bufs * set_bytes : 0);
  int stages = (kSmemLimit - 1024 - epi_bytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  static const char* force_stages = getenv("ALDI_CONV_STAGES");  // perf bisection only
  if (force_stages && atoi(force_stages) >= 1 && atoi(force_stages) < stages) stages = atoi(force_stages);
  // whatever the stage granularity leaves over deepens the residual / mask ring for free
  int epi_total = epi_bytes;
  while (set_bytes && epi_bufs < kMaxEpiBufs && stages * stage_bytes + epi_total + set_bytes + 1024 <= kSmemLimit) {
    ++epi_bufs;
    epi_total += set_bytes;
  }
  int smem_bytes = stages * stage_bytes + epi_total + 1024;

  int th, tw;
  pick_patch(p->ho, p->wo, (tma_epi && res_mode == 2) ? 64 : 128, &th, &tw);
  // 3x3 layers with few channels: load ONE 18 x 16-pixel halo tile per 64-channel chunk and address the nine shifted
  // A operands inside it (UMMA swizzles on absolute shared-memory address bits: a start address shifted by whole
  // 128-byte rows needs NO descriptor base offset as long as the 8-row groups stay 1024-byte aligned — measured,
  // tools/gpu_diag.py --case halo), with all 9 weight blocks resident in shared memory.  Measured 111.7 -> 89.1 us for
  // res2's 3x3 64 -> 64 at 4 x 256 x 512; what is left is the tensor pipe itself: a 128 x N x 16 tcgen05.mma costs the
  // same ~128 cycles for N = 64, 128 or 256 (the 3x3 layers scale 1 : 2 : 4 with N at every size), so N = 64 tiles
  // cannot pass 25 % of the bf16 peak.  With streamed weights (128 channels) the halo path is no faster, so it is
  // taken only when the weights fit.  ALDI_CONV_HALO: 0 off, 2 default, 3 also with streamed weights, 1 = base-offset
  // = tap column (wrong results; kept as the record of the descriptor experiment).
  static const char* halo_env = getenv("ALDI_CONV_HALO");
  const int halo_mode = halo_env ? atoi(halo_env) : 2;
  int halo = 0, halo_stages = 0, b_resident = 0;
  if (halo_mode > 0 && p->taps_h == 3 && p->taps_w == 3 && p->x_c <= 128 && res_mode != 2 && block_n <= 128) {
    const int bbytes = block_n * kBlockK * 2;
    const int nb = 9 * (p->x_c / 64);
    // weights resident when they fit beside two halo stages (64 -> 64: 72 KB); else a deeper ring
    const bool resident = p->cout_p == block_n && nb * bbytes + 2 * kHaloBytes + out_bytes + 2 * set_bytes + 1024 <= kSmemLimit;
    b_resident = resident ? 1 : 0;
    int bst = resident ? nb : 6;
    while (!resident && bst > 3 && (kSmemLimit - 1024 - out_bytes - 2 * set_bytes - bst * bbytes) / kHaloBytes < 2) --bst;
    int hst = (kSmemLimit - 1024 - out_bytes - 2 * set_bytes - bst * bbytes) / kHaloBytes;
    if (hst > kMaxHaloStages) hst = kMaxHaloStages;
    if (hst >= 2 && (resident || halo_mode == 3)) {
      halo = 1; halo_stages = hst; th = 16; tw = 8;
      stages = bst;
      epi_bufs = 2;
      int used = halo_stages * kHaloBytes + bst * bbytes + out_bytes + 2 * set_bytes + 1024;
      while (set_bytes && epi_bufs < kMaxEpiBufs && used + set_bytes <= kSmemLimit) { ++epi_bufs; used += set_bytes; }
    }
  }

  if (halo)
    smem_bytes = halo_stages * kHaloBytes + stages * block_n * kBlockK * 2 + out_bytes + (set_bytes ? epi_bufs * set_bytes : 0) + 1024;
  ConvArgs a;
  a.halo = halo; a.halo_stages = halo_stages; a.halo_bo = halo_mode; a.b_resident = halo ? b_resident : 0;
  a.n = p->n; a.ho = p->ho; a.wo = p->wo;
  a.th = th; a.tw = tw;
  a.tw_shift = 0;
  while ((1 << a.tw_shift) < tw) ++a.tw_shift;
  a.tiles_h = aldi_div_up(p->ho, th);
  a.tiles_w = aldi_div_up(p->wo, tw);
  a.num_n_tiles = p->cout_p / block_n;
  a.num_tiles = p->n * a.tiles_h * a.tiles_w * a.num_n_tiles;
  a.kchunks = p->x_c / 64;
  a.taps_w = p->taps_w;
  a.num_kb = p->taps_h * p->taps_w * a.kchunks;
  a.pad_h = p->pad_h; a.pad_w = p->pad_w;
  a.scale = p->scale; a.bias = p->bias;
  a.residual = reinterpret_cast<const __nv_bfloat16*>(p->residual);
  a.res_mode = res_mode;
  a.res_sw = p->res_sw; a.res_sh = p->res_sh; a.res_sn = p->res_sn;
  a.mask = reinterpret_cast<const __nv_bfloat16*>(p->mask);
  a.mask_sw = p->mask_sw; a.mask_sh = p->mask_sh; a.mask_sn = p->mask_sn;
  a.out = p->out;
  a.out_f32 = (p->out_dtype == ALDI_DTYPE_F32);
  a.cout_store = p->cout_store;
  a.out_sw = p->out_sw; a.out_sh = p->out_sh; a.out_sn = p->out_sn;
  a.relu = p->relu; a.accumulate = p->accumulate;
  a.stages = stages;
  a.tma_epi = tma_epi; a.epi_res = epi_res; a.epi_mask = epi_mask;
  a.epi_bufs = epi_bufs;
  static const char* epf = getenv("ALDI_EPI_PREFETCH");
  static const char* apf = getenv("ALDI_A_PREFETCH");
  // measured (tools/gpu_diag.py perf_res4): L2 prefetching costs 5-15 % on these layers — they are bound by L2
  // throughput, not latency, and a prefetch is one more pass through the L2 — so both stay off unless asked for
  a.epi_prefetch = (epi_res || epi_mask) ? (epf ? atoi(epf) : 0) : 0;
  a.a_prefetch = (p->taps_h * p->taps_w == 1) ? (apf ? atoi(apf) : 0) : 0;
  static const char* dbg = getenv("ALDI_CONV_DEBUG");
  a.debug = dbg ? atoi(dbg) : 0;

  CUtensorMap tm[5];
  int rc = make_cl_tmap(&tm[0], p->x, p->x_c, p->x_w, p->x_h, p->x_n, p->x_sw, p->x_sh, p->x_sn, halo ? kHaloW : tw,
                        halo ? kHaloH : th);
  if (rc) return rc;
  {
    uint64_t ktot = (uint64_t)p->taps_h * p->taps_w * p->x_c;
    uint64_t dims[2] = {ktot, (uint64_t)p->cout_p};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {64, (uint32_t)block_n};
    rc = aldi_make_tmap_bf16(&tm[1], p->w, 2, dims, strides, box);
    if (rc) return rc;
  }
  tm[2] = tm[0]; tm[3] = tm[0]; tm[4] = tm[0];  // placeholders when unused
  if (tma_epi) {
    // output leaves as per-warp 32-pixel slabs: box {64 ch, min(TW,32), 32 / min(TW,32), 1}
    const int bw = tw < 32 ? tw : 32;
    rc = make_cl_tmap(&tm[2], p->out, p->cout_store, p->wo, p->ho, p->n, p->out_sw, p->out_sh, p->out_sn, bw, 32 / bw);
    if (rc) return rc;
    if (epi_res) {
      if (p->accumulate)
        rc = make_cl_tmap(&tm[3], p->out, p->cout_store, p->wo, p->ho, p->n, p->out_sw, p->out_sh, p->out_sn, tw, th);
      else if (res_mode == 2)
        rc = make_cl_tmap(&tm[3], p->residual, p->cout_store, (p->wo + 1) / 2, (p->ho + 1) / 2, p->n, p->res_sw,
                          p->res_sh, p->res_sn, tw / 2, th / 2);
      else
        rc = make_cl_tmap(&tm[3], p->residual, p->cout_store, p->wo, p->ho, p->n, p->res_sw, p->res_sh, p->res_sn, tw,
                          th);
      if (rc) return rc;
    }
    if (epi_mask) {
      rc = make_cl_tmap(&tm[4], p->mask, p->cout_store, p->wo, p->ho, p->n, p->mask_sw, p->mask_sh, p->mask_sn, tw, th);
      if (rc) return rc;
    }
  }
  switch (block_n) {
    case 256: return launch_conv<256>(tm, a, smem_bytes, stream);
    case 128: return launch_conv<128>(tm, a, smem_bytes, stream);
    default: return launch_conv<64>(tm, a, smem_bytes, stream);
  }
}
