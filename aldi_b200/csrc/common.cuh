// Shared helpers for the aldi_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

#define ALDI_OK 0
#define ALDI_ERR_INVALID (-1)
#define ALDI_ERR_CUDA (-2)
#define ALDI_ERR_UNSUPPORTED (-3)

// dtype codes used across the C ABI
#define ALDI_F32 0
#define ALDI_BF16 1

void aldi_set_error(const char* fmt, ...);

#define ALDI_CHECK_ARG(cond, ...)                       \
  do {                                                  \
    if (!(cond)) {                                      \
      aldi_set_error(__VA_ARGS__);                      \
      return ALDI_ERR_INVALID;                          \
    }                                                   \
  } while (0)

#define ALDI_CUDA_LAUNCH_CHECK(name)                                          \
  do {                                                                        \
    cudaError_t _e = cudaGetLastError();                                      \
    if (_e != cudaSuccess) {                                                  \
      aldi_set_error("%s: launch failed: %s", name, cudaGetErrorString(_e));  \
      return ALDI_ERR_CUDA;                                                   \
    }                                                                         \
  } while (0)

static inline int aldi_div_up(long a, long b) { return (int)((a + b - 1) / b); }

// every launch of one of OUR kernels bumps this (bench.py reports it as gpu_launches)
extern unsigned long long g_aldi_launch_count;
#define ALDI_COUNT_LAUNCH() (++g_aldi_launch_count)

int aldi_num_sms();

// Programmatic dependent launch: the kernel may start (barrier init, TMEM alloc, descriptor prefetch) while the
// previous kernel on the stream drains; it must execute pdl_wait() before touching global memory.  ALDI_NO_PDL=1
// falls back to plain stream order (A/B knob).
bool aldi_pdl_enabled();
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t aldi_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                   Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = aldi_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// wait for the upstream grid's memory, then let the downstream grid begin its own prologue
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum (blockDim.x multiple of 32, <= 1024); result valid in every thread
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? smem32[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) smem32[0] = r;
  __syncthreads();
  return smem32[0];
}

#endif  // __CUDACC__
