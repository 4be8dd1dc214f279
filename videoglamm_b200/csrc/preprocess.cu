// GPU pre-processing of the SAM2 frames: the reference resizes every uint8 frame on the CPU with PIL (ResizeLongestSide(1024),
// utils/sam_transforms.py:26-65, model/segment_anything/utils/transforms.py:26-36), normalises in fp32, squashes to 1024x1024 with a second bilinear
// interpolation and uploads 12.6 MB of fp32 per frame (chat.py:447-456). Here the uint8 frames are uploaded as they are (H*W*3 bytes) and
//   vg_resize_u8_pil         reproduces PIL's 8-bit two-pass resampling BIT FOR BIT: horizontal pass to an 8-bit intermediate, then vertical pass,
//                            fixed-point coefficients (22 fractional bits, rounding constant 1 << 21, clip to [0, 255]); the coefficient / bound
//                            tables are PIL's precompute_coeffs + normalize_coeffs_8bpc evaluated on the host (videoglamm_b200/preprocess.py);
//   vg_sam_normalize_squash  (x - mean) / std per channel followed by F.interpolate(bilinear, align_corners=False) to S x S, HWC uint8 -> CHW fp32.
// Integer / byte work, HBM-bound: one thread per output element, channel-contiguous (coalesced) accesses.
#include <cstdint>

#include "common.h"
#include "pdl.cuh"

namespace vg {

constexpr int PIL_PRECISION_BITS = 22;

// one pass along x (axis = 0: [n][h][in][c] -> [n][h][out][c]) or along y (axis = 1: [n][in][w][c] -> [n][out][w][c])
__global__ void resize_u8_pass_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long total, int h_or_w, int in_size, int out_size, int c,
                                      const int* __restrict__ kk, const int* __restrict__ bounds, int ksize, int axis) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int ch = static_cast<int>(i % c);
  long long t = i / c;
  int o, other;
  long long img;
  if (axis == 0) {   // dst [n][h][out][c]
    o = static_cast<int>(t % out_size); t /= out_size;
    other = static_cast<int>(t % h_or_w); img = t / h_or_w;
  } else {           // dst [n][out][w][c]
    other = static_cast<int>(t % h_or_w); t /= h_or_w;
    o = static_cast<int>(t % out_size); img = t / out_size;
  }
  const int lo = bounds[2 * o], n = bounds[2 * o + 1];
  const int* k = kk + static_cast<long long>(o) * ksize;
  int ss = 1 << (PIL_PRECISION_BITS - 1);
  if (axis == 0) {
    const uint8_t* p = src + ((img * h_or_w + other) * in_size + lo) * c + ch;
    for (int x = 0; x < n; ++x) ss += static_cast<int>(p[static_cast<long long>(x) * c]) * k[x];
  } else {
    const uint8_t* p = src + ((img * in_size + lo) * h_or_w + other) * c + ch;
    for (int y = 0; y < n; ++y) ss += static_cast<int>(p[static_cast<long long>(y) * h_or_w * c]) * k[y];
  }
  ss >>= PIL_PRECISION_BITS;
  dst[i] = static_cast<uint8_t>(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
}

__global__ void sam_normalize_squash_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int n, int h, int w, int S, float m0, float m1, float m2,
                                            float s0, float s1, float s2) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(n) * 3 * S * S) return;
  const int ox = static_cast<int>(i % S), oy = static_cast<int>((i / S) % S);
  const int ch = static_cast<int>((i / (static_cast<long long>(S) * S)) % 3);
  const long long img = i / (3LL * S * S);
  const float mean = ch == 0 ? m0 : (ch == 1 ? m1 : m2), sd = ch == 0 ? s0 : (ch == 1 ? s1 : s2);
  // PyTorch upsample_bilinear2d, align_corners=False
  const float ry = static_cast<float>(h) / S, rx = static_cast<float>(w) / S;
  const float sy = fmaxf((oy + 0.5f) * ry - 0.5f, 0.f), sx = fmaxf((ox + 0.5f) * rx - 0.5f, 0.f);
  const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float ly = sy - y0, lx = sx - x0;
  const uint8_t* p = src + img * h * w * 3 + ch;
  auto px = [&](int y, int x) { return (static_cast<float>(p[(static_cast<long long>(y) * w + x) * 3]) - mean) / sd; };
  dst[i] = (1.f - ly) * ((1.f - lx) * px(y0, x0) + lx * px(y0, x1)) + ly * ((1.f - lx) * px(y1, x0) + lx * px(y1, x1));
}

// cv::resize(INTER_LINEAR) on 8-bit images: horizontal pass in 11-bit fixed point (two taps), vertical pass
// ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2 — one thread per output value, both passes fused.
__global__ void resize_u8_cv_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n, int h, int w, int c, int oh, int ow,
                                    const int* __restrict__ xi, const int* __restrict__ xc, const int* __restrict__ yi, const int* __restrict__ yc) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(n) * oh * ow * c) return;
  const int ch = static_cast<int>(i % c);
  long long t = i / c;
  const int ox = static_cast<int>(t % ow);
  t /= ow;
  const int oy = static_cast<int>(t % oh);
  const long long img = t / oh;
  const int x0 = xi[ox], x1 = min(x0 + 1, w - 1), a0 = xc[2 * ox], a1 = xc[2 * ox + 1];
  const int sy = yi[oy], y0 = min(max(sy, 0), h - 1), y1 = min(max(sy + 1, 0), h - 1), b0 = yc[2 * oy], b1 = yc[2 * oy + 1];
  const uint8_t* p = src + img * h * w * c + ch;
  const int s0 = p[(static_cast<long long>(y0) * w + x0) * c] * a0 + p[(static_cast<long long>(y0) * w + x1) * c] * a1;
  const int s1 = p[(static_cast<long long>(y1) * w + x0) * c] * a0 + p[(static_cast<long long>(y1) * w + x1) * c] * a1;
  const int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
  dst[i] = static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// uint8 HWC (cropped at (top, left)) -> fp32 CHW. mode 0: float32(x * (1/255) in double), then (v - mean) / std in float32 (HF rescale + normalize);
// mode 1: (x / 255 - mean) / std in double, rounded to float32 once (VideoTrainProcessor.normalize in numpy float64).
__global__ void u8_normalize_chw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int n, int h, int w, int oh, int ow, int top, int left, int mode,
                                        double m0, double m1, double m2, double s0, double s1, double s2) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(n) * 3 * oh * ow) return;
  const int ox = static_cast<int>(i % ow), oy = static_cast<int>((i / ow) % oh);
  const int ch = static_cast<int>((i / (static_cast<long long>(oh) * ow)) % 3);
  const long long img = i / (3LL * oh * ow);
  const double mean = ch == 0 ? m0 : (ch == 1 ? m1 : m2), sd = ch == 0 ? s0 : (ch == 1 ? s1 : s2);
  const double x = src[((img * h + oy + top) * w + ox + left) * 3 + ch];
  if (mode == 0) {
    const float v = static_cast<float>(x * 0.00392156862745098);
    dst[i] = __fdiv_rn(__fsub_rn(v, static_cast<float>(mean)), static_cast<float>(sd));
  } else {
    dst[i] = static_cast<float>((x / 255.0 - mean) / sd);
  }
}

}  // namespace vg

using namespace vg;

extern "C" int vg_resize_u8_cv(const unsigned char* src, unsigned char* dst, int n, int h, int w, int c, int oh, int ow, const int* xi, const int* xc, const int* yi,
                               const int* yc, void* stream) {
  if (!src || !dst || !xi || !xc || !yi || !yc || n <= 0 || h <= 0 || w <= 0 || c <= 0 || oh <= 0 || ow <= 0) return set_error(VG_EINVAL, "vg_resize_u8_cv: bad argument");
  const long long total = static_cast<long long>(n) * oh * ow * c;
  launch_k(resize_u8_cv_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), src, dst, n, h, w, c, oh, ow, xi, xc,
           yi, yc);
  count_launch();
  return check_launch("resize_u8_cv_kernel");
}

extern "C" int vg_u8_normalize_chw(const unsigned char* src, float* dst, int n, int h, int w, int oh, int ow, int top, int left, int mode, const double* mean3,
                                   const double* std3, void* stream) {
  if (!src || !dst || !mean3 || !std3 || n <= 0 || top < 0 || left < 0 || top + oh > h || left + ow > w || (mode != 0 && mode != 1))
    return set_error(VG_EINVAL, "vg_u8_normalize_chw: bad argument (the crop must lie inside the image)");
  const long long total = static_cast<long long>(n) * 3 * oh * ow;
  launch_k(u8_normalize_chw_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), src, dst, n, h, w, oh, ow, top,
           left, mode, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  count_launch();
  return check_launch("u8_normalize_chw_kernel");
}

extern "C" int vg_resize_u8_pil(const unsigned char* src, unsigned char* tmp, unsigned char* dst, int n, int h, int w, int c, int oh, int ow, const int* kh,
                                const int* bh, int ksize_h, const int* kv, const int* bv, int ksize_v, void* stream) {
  if (!src || !tmp || !dst || !kh || !bh || !kv || !bv || n <= 0 || h <= 0 || w <= 0 || c <= 0 || oh <= 0 || ow <= 0)
    return set_error(VG_EINVAL, "vg_resize_u8_pil: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long t1 = static_cast<long long>(n) * h * ow * c, t2 = static_cast<long long>(n) * oh * ow * c;
  launch_k(resize_u8_pass_kernel, dim3(static_cast<unsigned>((t1 + 255) / 256)), dim3(256), 0, s, src, tmp, t1, h, w, ow, c, kh, bh, ksize_h, 0);
  count_launch();
  launch_k(resize_u8_pass_kernel, dim3(static_cast<unsigned>((t2 + 255) / 256)), dim3(256), 0, s, static_cast<const uint8_t*>(tmp), dst, t2, ow, h, oh, c, kv, bv, ksize_v, 1);
  count_launch();
  return check_launch("resize_u8_pass_kernel");
}

extern "C" int vg_sam_normalize_squash(const unsigned char* src, float* dst, int n, int h, int w, int S, const float* mean3, const float* std3, void* stream) {
  if (!src || !dst || !mean3 || !std3 || n <= 0 || h <= 0 || w <= 0 || S <= 0) return set_error(VG_EINVAL, "vg_sam_normalize_squash: bad argument");
  const long long total = static_cast<long long>(n) * 3 * S * S;
  launch_k(sam_normalize_squash_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), src, dst, n, h, w, S,
           mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  count_launch();
  return check_launch("sam_normalize_squash_kernel");
}
