// Strong augmentation of a weakly-augmented image ON THE DEVICE: the weak view is already in HBM for
// the teacher, so the student's strong view is derived there instead of being built on dataloader workers with
// NumPy / scipy / cv2 and copied over PCIe a second time.
//
// Replaces, for planar uint8 images (3, H, W), what aldi/aug.py:39-60 `build_strong_augmentation` + the MIC option
// (aldi/aug.py:31-35) do per image, given parameters drawn on the host in the reference's RNG order
// (aldi_b200/augment.py):  colour jitter (detectron2 RandomContrast / RandomBrightness / RandomSaturation through
// fvcore BlendTransform: float32 / float64 blend, clip, truncating cast, after EACH op) -> random grayscale ->
// RandomBlurTransform (aldi/aug.py:81-93: scipy gaussian_filter over H, W AND the channel axis, reflect boundary,
// double accumulation in scipy's symmetric-kernel order, float32 between passes) -> up to three
// RandomEraseTransform rectangles (aldi/aug.py:106-143; the fill noise is a counter-based hash, not NumPy's stream)
// -> MICTransform block mask (aldi/aug.py:154-176, cv2 INTER_NEAREST index rule).
// Bit-exact against oracle/aug_ref.py (itself pinned to the reference's classes) outside the erased rectangles.
// Every kernel is a coalesced byte / float stream: HBM-bound, ~ (3 + 3 + 12 + 12 + 12 + 3 + 6) B per pixel-channel
// with blur, 8 B without.
#include "common.cuh"
#include "../../include/aldi_b200.h"

namespace {

struct Img {
  const unsigned char* src;
  unsigned char* dst;
  int h, w;
  long long sp, sr, dp, dr;   // plane / row strides (elements)
};

__global__ void __launch_bounds__(256) aug_sum_kernel(Img im, unsigned long long* out) {
  const long long total = 3LL * im.h * im.w;
  unsigned long long acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % im.w);
    const long long r = i / im.w;
    const int y = (int)(r % im.h), c = (int)(r / im.h);
    acc += im.src[c * im.sp + y * im.sr + x];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ unsigned long long part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(out, t);
  }
}

__device__ __forceinline__ unsigned char clip_u8(float v) { return (unsigned char)fminf(fmaxf(v, 0.f), 255.f); }
__device__ __forceinline__ unsigned char clip_u8(double v) { return (unsigned char)fmin(fmax(v, 0.0), 255.0); }

// one thread per pixel, all three channels (saturation / grayscale mix them)
__global__ void __launch_bounds__(256)
aug_color_kernel(Img im, const unsigned long long* sum, int do_color, double cw, double bw, double sw, int do_gray) {
  const long long total = (long long)im.h * im.w;
  float c_con = 0.f;
  if (do_color) {
    const double mean = (double)(*sum) / (double)(3LL * im.h * im.w);
    c_con = (float)((1.0 - cw) * mean);
  }
  const float fcw = (float)cw, fbw = (float)bw, fsw = (float)sw;
  const float c_bri = (float)((1.0 - bw) * 0.0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % im.w), y = (int)(i / im.w);
    unsigned char v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = im.src[c * im.sp + y * im.sr + x];
    if (do_color) {
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = clip_u8(__fadd_rn(c_con, __fmul_rn(fcw, (float)v[c])));   // RandomContrast
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = clip_u8(__fadd_rn(c_bri, __fmul_rn(fbw, (float)v[c])));   // RandomBrightness
      const double gray = __dadd_rn(__dadd_rn(__dmul_rn((double)v[0], 0.299), __dmul_rn((double)v[1], 0.587)),
                                    __dmul_rn((double)v[2], 0.114));
      const double sg = __dmul_rn(1.0 - sw, gray);
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = clip_u8(__dadd_rn(sg, (double)__fmul_rn(fsw, (float)v[c])));  // RandomSaturation
    }
    if (do_gray) {
      const double gray = __dadd_rn(__dadd_rn(__dmul_rn((double)v[0], 0.299), __dmul_rn((double)v[1], 0.587)),
                                    __dmul_rn((double)v[2], 0.114));
      const unsigned char g = clip_u8(__dadd_rn(__dmul_rn(1.0, gray), (double)__fmul_rn(0.f, (float)v[0])));
      v[0] = v[1] = v[2] = g;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) im.dst[c * im.dp + y * im.dr + x] = v[c];
  }
}

struct Taps {
  double w[2 * ALDI_AUG_MAX_RADIUS + 1];
  int r;
};

__device__ __forceinline__ int reflect_idx(int i, int n) {   // scipy 'reflect': d c b a | a b c d | d c b a
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

// 1-D correlation along one axis of a (3, h, w) volume; AXIS 0 = rows (H), 1 = columns (W), 2 = channels.
// scipy.ndimage correlate1d, symmetric branch: tmp = x[l]*w0; for j = -r..-1: tmp += w[j] * (x[l+j] + x[l-j]), double.
template <typename TIN, int AXIS, bool FINAL>
__global__ void __launch_bounds__(256)
aug_blur_kernel(const TIN* __restrict__ in, long long ip, long long ir, void* __restrict__ out_, long long op, long long orow,
                int h, int w, Taps t) {
  const long long total = 3LL * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const long long rr = i / w;
    const int y = (int)(rr % h), c = (int)(rr / h);
    const int len = AXIS == 0 ? h : (AXIS == 1 ? w : 3);
    const int pos = AXIS == 0 ? y : (AXIS == 1 ? x : c);
    auto at = [&](int q) -> double {
      const int qq = reflect_idx(q, len);
      const int yy = AXIS == 0 ? qq : y, xx = AXIS == 1 ? qq : x, cc = AXIS == 2 ? qq : c;
      return (double)in[cc * ip + yy * ir + xx];
    };
    double tmp = __dmul_rn(at(pos), t.w[t.r]);
    for (int j = -t.r; j < 0; ++j) tmp = __dadd_rn(tmp, __dmul_rn(t.w[j + t.r], __dadd_rn(at(pos + j), at(pos - j))));
    const float f = (float)tmp;
    if (FINAL)
      reinterpret_cast<unsigned char*>(out_)[c * op + y * orow + x] = clip_u8(f);
    else
      reinterpret_cast<float*>(out_)[c * op + y * orow + x] = f;
  }
}

__device__ __forceinline__ unsigned int hash3(unsigned int a, unsigned int b, unsigned int c) {
  unsigned int h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return h;
}

struct Post {
  int num_erase;
  int rect[3][4];          // h0, w0, h, w
  unsigned int seed[3];
  const unsigned char* mic;
  int mic_h, mic_w;
};

__global__ void __launch_bounds__(256) aug_post_kernel(Img im, Post p) {
  const long long total = 3LL * im.h * im.w;
  const double ify = p.mic ? 1.0 / ((double)im.h / (double)p.mic_h) : 0.0;
  const double ifx = p.mic ? 1.0 / ((double)im.w / (double)p.mic_w) : 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % im.w);
    const long long rr = i / im.w;
    const int y = (int)(rr % im.h), c = (int)(rr / im.h);
    unsigned char* q = im.dst + c * im.dp + y * im.dr + x;
    unsigned char v = *q;
    bool touched = false;
    for (int e = 0; e < p.num_erase; ++e) {
      if (y >= p.rect[e][0] && y < p.rect[e][0] + p.rect[e][2] && x >= p.rect[e][1] && x < p.rect[e][1] + p.rect[e][3]) {
        const float u = (float)(hash3(p.seed[e], (unsigned)(y * im.w + x), (unsigned)c) >> 8) * (1.0f / 16777216.0f);
        v = clip_u8(__fmul_rn(u, 255.f));
        touched = true;
      }
    }
    if (p.mic) {
      const int my = min((int)floor((double)y * ify), p.mic_h - 1), mx = min((int)floor((double)x * ifx), p.mic_w - 1);
      if (!p.mic[my * p.mic_w + mx]) { v = 0; touched = true; }
    }
    if (touched) *q = v;
  }
}

int grid_for_elems(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)aldi_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" size_t aldi_strong_augment_workspace_bytes(int h, int w) {
  return 256 + 2 * (size_t)3 * h * w * sizeof(float);
}

extern "C" int aldi_strong_augment(const unsigned char* src, unsigned char* dst, const aldi_aug_params* p, void* workspace,
                                   size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ALDI_CHECK_ARG(src && dst && p && workspace, "aldi_strong_augment: null pointer");
  ALDI_CHECK_ARG(p->h > 0 && p->w > 0, "aldi_strong_augment: empty image");
  ALDI_CHECK_ARG(workspace_bytes >= aldi_strong_augment_workspace_bytes(p->h, p->w), "aldi_strong_augment: workspace too small");
  ALDI_CHECK_ARG(p->blur_radius <= ALDI_AUG_MAX_RADIUS, "aldi_strong_augment: blur radius %d > %d", p->blur_radius,
                 ALDI_AUG_MAX_RADIUS);
  ALDI_CHECK_ARG(p->num_erase >= 0 && p->num_erase <= 3, "aldi_strong_augment: at most 3 erase rectangles");
  ALDI_CHECK_ARG(!p->mic_mask || (p->mic_h > 0 && p->mic_w > 0), "aldi_strong_augment: bad MIC mask shape");
  Img im{src, dst, p->h, p->w, p->src_plane, p->src_row, p->dst_plane, p->dst_row};
  unsigned long long* sum = reinterpret_cast<unsigned long long*>(workspace);
  float* fa = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  float* fb = fa + (size_t)3 * p->h * p->w;
  const long long px = (long long)p->h * p->w;
  if (p->do_color) {
    cudaMemsetAsync(sum, 0, sizeof(unsigned long long), stream);
    aug_sum_kernel<<<grid_for_elems(3 * px), 256, 0, stream>>>(im, sum);
    ALDI_COUNT_LAUNCH();
  }
  aug_color_kernel<<<grid_for_elems(px), 256, 0, stream>>>(im, sum, p->do_color, p->contrast_w, p->brightness_w,
                                                          p->saturation_w, p->do_gray);
  ALDI_COUNT_LAUNCH();
  if (p->blur_radius >= 0) {
    Taps t;
    t.r = p->blur_radius;
    for (int i = 0; i < 2 * ALDI_AUG_MAX_RADIUS + 1; ++i) t.w[i] = i <= 2 * t.r ? p->blur_taps[i] : 0.0;
    const int g = grid_for_elems(3 * px);
    const long long plane = px, row = p->w;
    aug_blur_kernel<unsigned char, 0, false><<<g, 256, 0, stream>>>(dst, p->dst_plane, p->dst_row, fa, plane, row, p->h, p->w, t);
    aug_blur_kernel<float, 1, false><<<g, 256, 0, stream>>>(fa, plane, row, fb, plane, row, p->h, p->w, t);
    aug_blur_kernel<float, 2, true><<<g, 256, 0, stream>>>(fb, plane, row, dst, p->dst_plane, p->dst_row, p->h, p->w, t);
    ALDI_COUNT_LAUNCH(); ALDI_COUNT_LAUNCH(); ALDI_COUNT_LAUNCH();
  }
  if (p->num_erase > 0 || p->mic_mask) {
    Post q;
    q.num_erase = p->num_erase;
    for (int e = 0; e < 3; ++e) {
      for (int k = 0; k < 4; ++k) q.rect[e][k] = p->erase_rect[e][k];
      q.seed[e] = p->erase_seed[e];
    }
    q.mic = p->mic_mask; q.mic_h = p->mic_h; q.mic_w = p->mic_w;
    aug_post_kernel<<<grid_for_elems(3 * px), 256, 0, stream>>>(im, q);
    ALDI_COUNT_LAUNCH();
  }
  ALDI_CUDA_LAUNCH_CHECK("aldi_strong_augment");
  return ALDI_OK;
}
