// Library bookkeeping for the C ABI: error string, launch counter, SM count, tensor-map encoder.
#include "common.cuh"
#include "tmap.h"
#include "../../include/aldi_b200.h"
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

static thread_local char g_err[512] = "";
unsigned long long g_aldi_launch_count = 0;

void aldi_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int aldi_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

static int g_pdl_on = -1;
bool aldi_pdl_enabled() {
  if (g_pdl_on < 0) {
    // off unless asked for (ALDI_PDL=1): measured on the round-2 step it changes nothing (27.38 ms with, 27.32 ms without:
    // the step replays as CUDA graphs and every kernel is a persistent grid whose successor cannot start its main loop
    // before the predecessor's last CTA retires anyway), and early-launched dependents only park on SMs that concurrent
    // NCCL kernels could use
    const char* e = getenv("ALDI_PDL");
    const char* off = getenv("ALDI_NO_PDL");
    g_pdl_on = (e && e[0] == '1' && !(off && off[0] == '1')) ? 1 : 0;
  }
  return g_pdl_on != 0;
}
extern "C" void aldi_set_pdl(int on) { g_pdl_on = on ? 1 : 0; }

extern "C" const char* aldi_last_error(void) { return g_err; }
extern "C" int aldi_abi_version(void) { return 1; }
extern "C" unsigned long long aldi_launch_count(void) { return g_aldi_launch_count; }
extern "C" void aldi_reset_launch_count(void) { g_aldi_launch_count = 0; }

int aldi_make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box) {
  return aldi_make_tmap_bf16_sw(out, base, rank, dims, strides_bytes, box, 1);
}

int aldi_make_tmap_bf16_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      aldi_set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
      return ALDI_ERR_CUDA;
    }
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                      bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    aldi_set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]",
                   (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                   (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                   rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return ALDI_ERR_CUDA;
  }
  return ALDI_OK;
}
