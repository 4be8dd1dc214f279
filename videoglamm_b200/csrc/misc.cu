// Normalisation, elementwise, pooling, resampling, small convolutions, RoPE, embedding, argmax.
// All HBM-bound: coalesced channel-contiguous access, fp32 math, dtype (bf16/fp32) chosen at run time.
#include "common.h"
#include "pdl.cuh"
#include "ptx.cuh"

namespace vg {

__device__ __forceinline__ float ld(const void* p, int dt, long long i) {
  return dt == VG_F32 ? reinterpret_cast<const float*>(p)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ float ld_(const void* p, int dt, long long i) { return ld(p, dt, i); }  // `ld` is shadowed by stride parameters in some kernels
__device__ __forceinline__ void st(void* p, int dt, long long i, float v) {
  if (dt == VG_F32)
    reinterpret_cast<float*>(p)[i] = v;
  else
    reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16(v);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }  // exact: small tensors only

static inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------ norm
// one warp per row; two passes over registers-cached row when cols <= 32*MAXV, else re-read.
template <int MAXV>
__global__ void norm_kernel(const void* x, int xdt, long long ldx, void* y, int ydt, long long ldy, const float* w, const float* b,
                            int rows, int cols, float eps, int rms) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long long xo = static_cast<long long>(row) * ldx, yo = static_cast<long long>(row) * ldy;
  float v[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    v[i] = c < cols ? ld(x, xdt, xo + c) : 0.f;
    sum += v[i];
  }
  float mean = 0.f;
  if (!rms) mean = warp_sum(sum) / cols;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    const float d = c < cols ? v[i] - mean : 0.f;
    sq += d * d;
  }
  const float rstd = rsqrtf(warp_sum(sq) / cols + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < cols) {
      float o = (v[i] - mean) * rstd * w[c];
      if (b) o += b[c];
      st(y, ydt, yo + c, o);
    }
  }
}

// fp32-in fast path: one warp per row, 128-bit loads (lane owns float4 chunks lane, lane+32, ...), 64/128-bit stores
template <int MAXC, bool XBF16 = false, bool GELU = false>
__global__ void norm_vec_kernel(const void* __restrict__ xv, long long ldx, void* y, int ydt, long long ldy, const float* __restrict__ w,
                                const float* __restrict__ b, int rows, int cols, float eps, int rms) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(static_cast<const float*>(xv) + static_cast<long long>(row) * ldx);
  const uint2* xh = reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(xv) + static_cast<long long>(row) * ldx);  // 4 bf16 per chunk
  const int nch = cols >> 2;
  float4 v[MAXC];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = lane + i * 32;
    if (XBF16) {
      const uint2 u = c < nch ? xh[c] : make_uint2(0u, 0u);
      v[i] = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
    } else {
      v[i] = c < nch ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    sum += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  float mean = 0.f;
  if (!rms) mean = warp_sum(sum) / cols;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    if (lane + i * 32 < nch) {
      const float a = v[i].x - mean, c2 = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
      sq += a * a + c2 * c2 + d * d + e * e;
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / cols + eps);
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = lane + i * 32;
    if (c < nch) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + c);
      float4 o = make_float4((v[i].x - mean) * rstd * ww.x, (v[i].y - mean) * rstd * ww.y, (v[i].z - mean) * rstd * ww.z, (v[i].w - mean) * rstd * ww.w);
      if (b) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + c);
        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
      }
      if (GELU) {
        o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w);
      }
      if (ydt == VG_F32)
        *(reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + static_cast<long long>(row) * ldy) + c) = o;
      else
        *(reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + static_cast<long long>(row) * ldy) + c) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
  }
}

// one 256-thread block per row (few, wide rows: LLM decode), values cached in registers, block reduction through smem
template <int MAXV>
__global__ void __launch_bounds__(256) norm_block_kernel(const void* x, int xdt, long long ldx, void* y, int ydt, long long ldy, const float* w,
                                                         const float* b, int cols, float eps, int rms) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  __shared__ float red[8];
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long xo = static_cast<long long>(row) * ldx, yo = static_cast<long long>(row) * ldy;
  float v[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + i * 256;
    v[i] = c < cols ? ld(x, xdt, xo + c) : 0.f;
    sum += v[i];
  }
  auto block_sum = [&](float t) {
    t = warp_sum(t);
    __syncthreads();
    if (lane == 0) red[wid] = t;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += red[i];
    return r;
  };
  float mean = 0.f;
  if (!rms) mean = block_sum(sum) / cols;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + i * 256;
    const float d = c < cols ? v[i] - mean : 0.f;
    sq += d * d;
  }
  const float rstd = rsqrtf(block_sum(sq) / cols + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + i * 256;
    if (c < cols) {
      float o = (v[i] - mean) * rstd * w[c];
      if (b) o += b[c];
      st(y, ydt, yo + c, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------ add / cast
__global__ void add_kernel(const void* a, int adt, long long lda, int a_rows, const void* b, int bdt, long long ldb, int b_rows, float alpha, void* out,
                           int odt, long long ldo, long long rows, int cols) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = static_cast<int>(i - r * cols);
  float v = ld(a, adt, (r % a_rows) * lda + c);
  if (b) v += alpha * ld(b, bdt, (r % b_rows) * ldb + c);
  st(out, odt, r * ldo + c, v);
}

// fp32 + fp32 -> fp32, four columns per thread (16-byte loads / stores); ncu on the scalar kernel above at 524288 x 144 (Hiera position
// embedding): 396 us for 868 MB of DRAM traffic = 2.2 TB/s (64-bit division per element, 4-byte accesses).
__global__ void add_vec4_kernel(const float* __restrict__ a, long long lda, int a_rows, const float* __restrict__ b, long long ldb, int b_rows, float alpha,
                                float* __restrict__ out, long long ldo, long long rows, int cols4) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= rows * cols4) return;
  const long long r = i / cols4;
  const int c = static_cast<int>(i - r * cols4) * 4;
  float4 v = *reinterpret_cast<const float4*>(a + (r % a_rows) * lda + c);
  if (b) {
    const float4 w = *reinterpret_cast<const float4*>(b + (r % b_rows) * ldb + c);
    v.x += alpha * w.x; v.y += alpha * w.y; v.z += alpha * w.z; v.w += alpha * w.w;
  }
  *reinterpret_cast<float4*>(out + r * ldo + c) = v;
}

// out[i] = resid[i] + sum_p parts[p][i] (fp32): reduction of the K-split partial products of the batched-decode projections
__global__ void sum_parts_kernel(const float* __restrict__ parts, long long n, int nparts, const float* resid, float* out) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float v = resid ? resid[i] : 0.f;
  for (int p = 0; p < nparts; ++p) v += parts[p * n + i];
  out[i] = v;
}

__global__ void pool2x2_kernel(const void* x, int xdt, long long ldx, void* y, int ydt, long long ldy, int n, int h, int w, int c, int is_max) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int oh = h / 2, ow = w / 2;
  if (i >= static_cast<long long>(n) * oh * ow * c) return;
  const int ch = static_cast<int>(i % c);
  long long t = i / c;
  const int ox = static_cast<int>(t % ow);
  t /= ow;
  const int oy = static_cast<int>(t % oh);
  const int img = static_cast<int>(t / oh);
  const long long base = (static_cast<long long>(img) * h + 2 * oy) * w + 2 * ox;
  const float v00 = ld(x, xdt, base * ldx + ch), v01 = ld(x, xdt, (base + 1) * ldx + ch);
  const float v10 = ld(x, xdt, (base + w) * ldx + ch), v11 = ld(x, xdt, (base + w + 1) * ldx + ch);
  const float v = is_max ? fmaxf(fmaxf(v00, v01), fmaxf(v10, v11)) : 0.25f * (v00 + v01 + v10 + v11);
  st(y, ydt, ((static_cast<long long>(img) * oh + oy) * ow + ox) * ldy + ch, v);
}

// 8 channels (bf16) or 4 channels (fp32) per thread, 16-byte loads and stores; the scalar kernel above ran at ~1/5 of the HBM roofline on the
// Hiera stage transitions (130 us for 150 MB in).
template <bool BF16>
__global__ void pool2x2_vec_kernel(const void* x, long long ldx, void* y, long long ldy, int n, int h, int w, int c, int is_max) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int V = BF16 ? 8 : 4;
  const int cv = c / V;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int oh = h / 2, ow = w / 2;
  if (i >= static_cast<long long>(n) * oh * ow * cv) return;
  const int ch = static_cast<int>(i % cv) * V;
  long long t = i / cv;
  const int ox = static_cast<int>(t % ow);
  t /= ow;
  const int oy = static_cast<int>(t % oh);
  const int img = static_cast<int>(t / oh);
  const long long base = (static_cast<long long>(img) * h + 2 * oy) * w + 2 * ox;
  const long long o00 = base * ldx + ch, o01 = o00 + ldx, o10 = o00 + w * ldx, o11 = o10 + ldx;
  const long long oo = ((static_cast<long long>(img) * oh + oy) * ow + ox) * ldy + ch;
  float r[8];
  if (BF16) {
    const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
    const uint4 a = *reinterpret_cast<const uint4*>(xp + o00), b = *reinterpret_cast<const uint4*>(xp + o01);
    const uint4 cc = *reinterpret_cast<const uint4*>(xp + o10), d = *reinterpret_cast<const uint4*>(xp + o11);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, cw[4] = {cc.x, cc.y, cc.z, cc.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a0 = __uint_as_float(aw[k] << 16), a1 = __uint_as_float(aw[k] & 0xffff0000u), b0 = __uint_as_float(bw[k] << 16), b1 = __uint_as_float(bw[k] & 0xffff0000u);
      const float c0 = __uint_as_float(cw[k] << 16), c1 = __uint_as_float(cw[k] & 0xffff0000u), d0 = __uint_as_float(dw[k] << 16), d1 = __uint_as_float(dw[k] & 0xffff0000u);
      r[2 * k] = is_max ? fmaxf(fmaxf(a0, b0), fmaxf(c0, d0)) : 0.25f * (a0 + b0 + c0 + d0);
      r[2 * k + 1] = is_max ? fmaxf(fmaxf(a1, b1), fmaxf(c1, d1)) : 0.25f * (a1 + b1 + c1 + d1);
    }
    uint32_t pk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat162 v2 = __floats2bfloat162_rn(r[2 * k], r[2 * k + 1]);
      pk[k] = *reinterpret_cast<uint32_t*>(&v2);
    }
    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(y) + oo) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  } else {
    const float* xp = static_cast<const float*>(x);
    const float4 a = *reinterpret_cast<const float4*>(xp + o00), b = *reinterpret_cast<const float4*>(xp + o01);
    const float4 cc = *reinterpret_cast<const float4*>(xp + o10), d = *reinterpret_cast<const float4*>(xp + o11);
    float4 o;
    o.x = is_max ? fmaxf(fmaxf(a.x, b.x), fmaxf(cc.x, d.x)) : 0.25f * (a.x + b.x + cc.x + d.x);
    o.y = is_max ? fmaxf(fmaxf(a.y, b.y), fmaxf(cc.y, d.y)) : 0.25f * (a.y + b.y + cc.y + d.y);
    o.z = is_max ? fmaxf(fmaxf(a.z, b.z), fmaxf(cc.z, d.z)) : 0.25f * (a.z + b.z + cc.z + d.z);
    o.w = is_max ? fmaxf(fmaxf(a.w, b.w), fmaxf(cc.w, d.w)) : 0.25f * (a.w + b.w + cc.w + d.w);
    *reinterpret_cast<float4*>(static_cast<float*>(y) + oo) = o;
  }
}

__global__ void upsample2x_add_kernel(const float* lat, const float* top, float* out, int n, int h, int w, int c) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int H = 2 * h, W = 2 * w;
  if (i >= static_cast<long long>(n) * H * W * c) return;
  const int ch = static_cast<int>(i % c);
  long long t = i / c;
  const int X = static_cast<int>(t % W);
  t /= W;
  const int Y = static_cast<int>(t % H);
  const int img = static_cast<int>(t / H);
  out[i] = lat[i] + top[((static_cast<long long>(img) * h + Y / 2) * w + X / 2) * c + ch];
}

// one warp per output pixel; cout <= 128 (4 values per lane)
__global__ void convt2x2_epilogue_kernel(const void* g, int gdt, const float* bias, const void* feat, int fdt, const float* ln_w,
                                         const float* ln_b, float eps, void* out, int odt, int n, int h, int w, int cout, int feat_bcast) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const long long pix = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int H = 2 * h, W = 2 * w;
  if (pix >= static_cast<long long>(n) * H * W) return;
  const int X = static_cast<int>(pix % W);
  const int Y = static_cast<int>((pix / W) % H);
  const int img = static_cast<int>(pix / (static_cast<long long>(W) * H));
  const long long src_row = (static_cast<long long>(img) * h + Y / 2) * w + X / 2;
  const int sub = (Y & 1) * 2 + (X & 1);
  float v[4];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = lane + i * 32;
    v[i] = 0.f;
    if (c < cout) {
      v[i] = ld(g, gdt, src_row * (4LL * cout) + sub * cout + c) + bias[c];
      if (feat) v[i] += ld(feat, fdt, (feat_bcast ? pix % (static_cast<long long>(H) * W) : pix) * cout + c);
      sum += v[i];
    }
  }
  if (ln_w) {
    const float mean = warp_sum(sum) / cout;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + i * 32;
      const float d = c < cout ? v[i] - mean : 0.f;
      sq += d * d;
    }
    const float rstd = rsqrtf(warp_sum(sq) / cout + eps);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + i * 32;
      if (c < cout) v[i] = (v[i] - mean) * rstd * ln_w[c] + ln_b[c];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = lane + i * 32;
    if (c < cout) st(out, odt, pix * cout + c, gelu_erf(v[i]));
  }
}

__global__ void im2col_kernel(const void* x, int xdt, void* out, int odt, int n, int cin, int h, int w, int k, int s, int p, int oh, int ow,
                              int kpad) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(n) * oh * ow * kpad) return;
  const int col = static_cast<int>(i % kpad);
  long long t = i / kpad;
  const int ox = static_cast<int>(t % ow);
  t /= ow;
  const int oy = static_cast<int>(t % oh);
  const int img = static_cast<int>(t / oh);
  float v = 0.f;
  if (col < cin * k * k) {
    const int ci = col / (k * k), rem = col - ci * k * k;
    const int ky = rem / k, kx = rem - ky * k;
    const int iy = oy * s - p + ky, ix = ox * s - p + kx;
    if (iy >= 0 && iy < h && ix >= 0 && ix < w) v = ld(x, xdt, ((static_cast<long long>(img) * cin + ci) * h + iy) * w + ix);
  }
  st(out, odt, i, v);
}

// bf16 output, eight consecutive columns of one patch row per thread (one 16-byte store; the scalar kernel above wrote 2 bytes per thread:
// 604 us for the 8-frame 7x7/4 patch embedding, 218 MB of traffic)
template <int KC>   // KC > 0: compile-time kernel size (the column -> (channel, ky, kx) split is two integer divisions per element)
__global__ void im2col_bf16x8_kernel(const void* __restrict__ x, int xdt, __nv_bfloat16* __restrict__ out, int n, int cin, int h, int w, int k_rt, int s, int p,
                                     int oh, int ow, int kpad) {
  const int k = KC > 0 ? KC : k_rt;
  pdl_wait();
  pdl_launch_dependents();
  const int chunks = kpad >> 3;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(n) * oh * ow * chunks) return;
  const int c8 = static_cast<int>(i % chunks);
  long long t = i / chunks;
  const int ox = static_cast<int>(t % ow);
  t /= ow;
  const int oy = static_cast<int>(t % oh);
  const int img = static_cast<int>(t / oh);
  const int kk = k * k, valid = cin * kk;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = c8 * 8 + j;
    v[j] = 0.f;
    if (col < valid) {
      const int ci = col / kk, rem = col - ci * kk;
      const int ky = rem / k, kx = rem - ky * k;
      const int iy = oy * s - p + ky, ix = ox * s - p + kx;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) v[j] = ld(x, xdt, ((static_cast<long long>(img) * cin + ci) * h + iy) * w + ix);
    }
  }
  uint32_t pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    pk[j] = *reinterpret_cast<uint32_t*>(&h2);
  }
  *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

__global__ void dwconv_kernel(const void* x, int xdt, const float* wt, const float* bias, void* y, int ydt, int n, int h, int w, int c, int k) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(n) * h * w * c) return;
  const int ch = static_cast<int>(i % c);
  long long t = i / c;
  const int px = static_cast<int>(t % w);
  t /= w;
  const int py = static_cast<int>(t % h);
  const int img = static_cast<int>(t / h);
  const int r = k / 2;
  float acc = bias ? bias[ch] : 0.f;
  for (int ky = 0; ky < k; ++ky) {
    const int iy = py + ky - r;
    if (iy < 0 || iy >= h) continue;
    for (int kx = 0; kx < k; ++kx) {
      const int ix = px + kx - r;
      if (ix < 0 || ix >= w) continue;
      acc = fmaf(ld(x, xdt, ((static_cast<long long>(img) * h + iy) * w + ix) * c + ch), wt[(ch * k + ky) * k + kx], acc);
    }
  }
  st(y, ydt, i, acc);
}

// fp32 channels-last depthwise KxK: a thread owns 4 channels x PX consecutive pixels of one row: per kernel row it loads the PX + K - 1 input
// float4s once and slides the K taps over them in registers (the scalar kernel above: 49 dependent loads per output, 68 us for the 64x64x256
// maps of the SAM2 memory fuser).
template <int K, int PX>
__global__ void dwconv_vec_kernel(const float* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ bias, void* y, int ydt, int n, int h, int w,
                                  int c) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int R = K / 2;
  const int cgs = c >> 2, xts = w / PX;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(n) * h * xts * cgs) return;
  const int cg = static_cast<int>(idx % cgs);
  long long t = idx / cgs;
  const int xt = static_cast<int>(t % xts);
  t /= xts;
  const int py = static_cast<int>(t % h);
  const int img = static_cast<int>(t / h);
  const int ch = cg * 4, px0 = xt * PX;
  float4 acc[PX];
  const float4 b4 = bias ? *reinterpret_cast<const float4*>(bias + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int p = 0; p < PX; ++p) acc[p] = b4;
#pragma unroll 1
  for (int ky = 0; ky < K; ++ky) {
    const int iy = py + ky - R;
    if (iy < 0 || iy >= h) continue;
    const float* xr = x + (static_cast<long long>(img) * h + iy) * w * c + ch;
    float4 row[PX + K - 1];
#pragma unroll
    for (int q = 0; q < PX + K - 1; ++q) {
      const int ix = px0 + q - R;
      row[q] = (ix >= 0 && ix < w) ? *reinterpret_cast<const float4*>(xr + static_cast<long long>(ix) * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float wk[4][K];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int kx = 0; kx < K; ++kx) wk[j][kx] = __ldg(wt + ((ch + j) * K + ky) * K + kx);
#pragma unroll
    for (int kx = 0; kx < K; ++kx)
#pragma unroll
      for (int p = 0; p < PX; ++p) {
        acc[p].x = fmaf(row[p + kx].x, wk[0][kx], acc[p].x);
        acc[p].y = fmaf(row[p + kx].y, wk[1][kx], acc[p].y);
        acc[p].z = fmaf(row[p + kx].z, wk[2][kx], acc[p].z);
        acc[p].w = fmaf(row[p + kx].w, wk[3][kx], acc[p].w);
      }
  }
  const long long o0 = ((static_cast<long long>(img) * h + py) * w + px0) * c + ch;
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    if (ydt == VG_F32) {
      *reinterpret_cast<float4*>(static_cast<float*>(y) + o0 + static_cast<long long>(p) * c) = acc[p];
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(acc[p].x, acc[p].y), hi = __floats2bfloat162_rn(acc[p].z, acc[p].w);
      *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(y) + o0 + static_cast<long long>(p) * c) =
          make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
}

// one warp per output pixel; lanes stride over cout (<= 256 -> 8 per lane)
__global__ void conv3x3s2_ln_gelu_kernel(const void* x, int xdt, const float* wt, const float* bias, const float* ln_w, const float* ln_b,
                                         float eps, void* y, int ydt, int n, int h, int w, int cin, int cout) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const long long pix = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int oh = h / 2, ow = w / 2;
  if (pix >= static_cast<long long>(n) * oh * ow) return;
  const int ox = static_cast<int>(pix % ow);
  const int oy = static_cast<int>((pix / ow) % oh);
  const int img = static_cast<int>(pix / (static_cast<long long>(ow) * oh));
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int co = lane + i * 32;
    acc[i] = co < cout ? bias[co] : 0.f;
  }
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = 2 * oy - 1 + ky;
    if (iy < 0 || iy >= h) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = 2 * ox - 1 + kx;
      if (ix < 0 || ix >= w) continue;
      const long long xb = ((static_cast<long long>(img) * h + iy) * w + ix) * cin;
      const float* wk = wt + static_cast<long long>((ky * 3 + kx) * cin) * cout;
      for (int ci = 0; ci < cin; ++ci) {
        const float xv = ld(x, xdt, xb + ci);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int co = lane + i * 32;
          if (co < cout) acc[i] = fmaf(xv, wk[static_cast<long long>(ci) * cout + co], acc[i]);
        }
      }
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += (lane + i * 32 < cout) ? acc[i] : 0.f;
  const float mean = warp_sum(sum) / cout;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float d = (lane + i * 32 < cout) ? acc[i] - mean : 0.f;
    sq += d * d;
  }
  const float rstd = rsqrtf(warp_sum(sq) / cout + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int co = lane + i * 32;
    if (co < cout) st(y, ydt, pix * cout + co, gelu_erf((acc[i] - mean) * rstd * ln_w[co] + ln_b[co]));
  }
}

// PX consecutive output pixels of one row per warp: every weight value a lane loads is reused for PX pixels (the one-pixel kernel re-read the
// whole [9*cin, cout] fp32 weight block per pixel: 590 KB x 4096 pixels = 2.4 GB of L2 traffic for the 64 -> 256 stage, 215 us per launch)
template <int PX>
__global__ void __launch_bounds__(256) conv3x3s2_ln_gelu_px_kernel(const void* x, int xdt, const float* __restrict__ wt, const float* __restrict__ bias,
                                                                   const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, void* y,
                                                                   int ydt, int n, int h, int w, int cin, int cout) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const int oh = h / 2, ow = w / 2, gpr = ow / PX;  // pixel groups per output row
  const long long grp = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (grp >= static_cast<long long>(n) * oh * gpr) return;
  const int ox0 = static_cast<int>(grp % gpr) * PX;
  const int oy = static_cast<int>((grp / gpr) % oh);
  const int img = static_cast<int>(grp / (static_cast<long long>(gpr) * oh));
  float acc[PX][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int co = lane + i * 32;
    const float bv = co < cout ? bias[co] : 0.f;
#pragma unroll
    for (int q = 0; q < PX; ++q) acc[q][i] = bv;
  }
  const int nco = (cout + 31) / 32;  // live accumulator slots per lane (warp-uniform)
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = 2 * oy - 1 + ky;
    if (iy < 0 || iy >= h) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const float* wk = wt + static_cast<long long>((ky * 3 + kx) * cin) * cout;
      const long long rowb = (static_cast<long long>(img) * h + iy) * w;
      for (int ci = 0; ci < cin; ++ci) {
        float wv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int co = lane + i * 32;
          wv[i] = (i < nco && co < cout) ? wk[static_cast<long long>(ci) * cout + co] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < PX; ++q) {
          const int ix = 2 * (ox0 + q) - 1 + kx;
          const float xv = (ix >= 0 && ix < w) ? ld(x, xdt, (rowb + ix) * cin + ci) : 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < nco) acc[q][i] = fmaf(xv, wv[i], acc[q][i]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PX; ++q) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += (lane + i * 32 < cout) ? acc[q][i] : 0.f;
    const float mean = warp_sum(sum) / cout;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = (lane + i * 32 < cout) ? acc[q][i] - mean : 0.f;
      sq += d * d;
    }
    const float rstd = rsqrtf(warp_sum(sq) / cout + eps);
    const long long pix = (static_cast<long long>(img) * oh + oy) * ow + ox0 + q;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int co = lane + i * 32;
      if (co < cout) st(y, ydt, pix * cout + co, gelu_erf((acc[q][i] - mean) * rstd * ln_w[co] + ln_b[co]));
    }
  }
}

// small-channel variant: one THREAD per output pixel, all COUT channels + LayerNorm + GELU in registers (the warp-per-pixel kernel
// leaves 28 of 32 lanes idle when cout = 4)
template <int CIN, int COUT>
__global__ void conv3x3s2_ln_gelu_small_kernel(const void* x, int xdt, const float* __restrict__ wt, const float* __restrict__ bias,
                                               const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, void* y, int ydt, int n,
                                               int h, int w) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const long long pix = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int oh = h / 2, ow = w / 2;
  if (pix >= static_cast<long long>(n) * oh * ow) return;
  const int ox = static_cast<int>(pix % ow);
  const int oy = static_cast<int>((pix / ow) % oh);
  const int img = static_cast<int>(pix / (static_cast<long long>(ow) * oh));
  float acc[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) acc[co] = bias[co];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = 2 * oy - 1 + ky;
    if (iy < 0 || iy >= h) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = 2 * ox - 1 + kx;
      if (ix < 0 || ix >= w) continue;
      const long long xb = ((static_cast<long long>(img) * h + iy) * w + ix) * CIN;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float xv = ld(x, xdt, xb + ci);
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[co] = fmaf(xv, wt[((ky * 3 + kx) * CIN + ci) * COUT + co], acc[co]);
      }
    }
  }
  float mean = 0.f;
#pragma unroll
  for (int co = 0; co < COUT; ++co) mean += acc[co];
  mean /= COUT;
  float var = 0.f;
#pragma unroll
  for (int co = 0; co < COUT; ++co) var += (acc[co] - mean) * (acc[co] - mean);
  const float rstd = rsqrtf(var / COUT + eps);
#pragma unroll
  for (int co = 0; co < COUT; ++co) st(y, ydt, pix * COUT + co, gelu_erf((acc[co] - mean) * rstd * ln_w[co] + ln_b[co]));
}

__global__ void bilinear_kernel(const float* x, void* y, int n, int ih, int iw, int oh, int ow, int mode, float scale, float bias) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(n) * oh * ow) return;
  const int ox = static_cast<int>(i % ow);
  const int oy = static_cast<int>((i / ow) % oh);
  const int img = static_cast<int>(i / (static_cast<long long>(ow) * oh));
  // PyTorch upsample_bilinear2d, align_corners=False: src = max(0, (dst + 0.5) * (in/out) - 0.5)
  const float ry = static_cast<float>(ih) / oh, rx = static_cast<float>(iw) / ow;
  const float sy = fmaxf((oy + 0.5f) * ry - 0.5f, 0.f), sx = fmaxf((ox + 0.5f) * rx - 0.5f, 0.f);
  const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
  const int y1 = min(y0 + 1, ih - 1), x1 = min(x0 + 1, iw - 1);
  const float ly = sy - y0, lx = sx - x0;
  const float* p = x + static_cast<long long>(img) * ih * iw;
  const float v = (1.f - ly) * ((1.f - lx) * p[y0 * iw + x0] + lx * p[y0 * iw + x1]) + ly * ((1.f - lx) * p[y1 * iw + x0] + lx * p[y1 * iw + x1]);
  if (mode == 3) {
    reinterpret_cast<uint8_t*>(y)[i] = v > 0.f ? 1 : 0;
    return;
  }
  float o = v;
  if (mode == 1) o = scale / (1.f + expf(-v)) + bias;
  if (mode == 2) o = (v > 0.f ? scale : 0.f) + bias;
  reinterpret_cast<float*>(y)[i] = o;
}

// mode 4 of vg_bilinear: threshold (> 0) and bit-pack, one thread per output BYTE = 8 consecutive pixels of one plane (MSB first, like
// numpy.packbits); each plane is padded to a whole number of bytes. Divides the mask bytes that leave the GPU by 8.
__global__ void bilinear_pack_kernel(const float* x, uint8_t* y, int n, int ih, int iw, int oh, int ow) {
  pdl_wait();
  pdl_launch_dependents();
  const long long bpp = (static_cast<long long>(oh) * ow + 7) / 8;  // bytes per plane
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n * bpp) return;
  const int img = static_cast<int>(i / bpp);
  const long long byte = i - img * bpp;
  const float ry = static_cast<float>(ih) / oh, rx = static_cast<float>(iw) / ow;
  const float* p = x + static_cast<long long>(img) * ih * iw;
  unsigned bits = 0;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const long long px = byte * 8 + b;
    if (px >= static_cast<long long>(oh) * ow) break;
    const int ox = static_cast<int>(px % ow), oy = static_cast<int>(px / ow);
    const float sy = fmaxf((oy + 0.5f) * ry - 0.5f, 0.f), sx = fmaxf((ox + 0.5f) * rx - 0.5f, 0.f);
    const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
    const int y1 = min(y0 + 1, ih - 1), x1 = min(x0 + 1, iw - 1);
    const float ly = sy - y0, lx = sx - x0;
    const float v = (1.f - ly) * ((1.f - lx) * p[y0 * iw + x0] + lx * p[y0 * iw + x1]) + ly * ((1.f - lx) * p[y1 * iw + x0] + lx * p[y1 * iw + x1]);
    if (v > 0.f) bits |= 0x80u >> b;
  }
  y[i] = static_cast<uint8_t>(bits);
}

__global__ void rope2d_kernel(void* x, int xdt, long long ld, int nrows, int dim, int gw, int gh) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int half = dim / 2;  // complex pairs per row
  if (i >= static_cast<long long>(nrows) * half) return;
  const int pr = static_cast<int>(i % half);
  const long long r = i / half;
  const int quarter = dim / 4;
  const int px = static_cast<int>(r % gw), py = static_cast<int>((r / gw) % gh);
  const int fi = pr < quarter ? pr : pr - quarter;
  const float freq = 1.0f / powf(10000.0f, static_cast<float>(4 * fi) / dim);
  const float ang = (pr < quarter ? px : py) * freq;
  float sn, cs;
  sincosf(ang, &sn, &cs);
  const long long e = r * ld + 2 * pr;
  const float a = ld_(x, xdt, e), b = ld_(x, xdt, e + 1);
  st(x, xdt, e, a * cs - b * sn);
  st(x, xdt, e + 1, a * sn + b * cs);
}

// The same rotation with the (cos, sin) of every (grid position, pair) read from a table filled once by rope2d_table_kernel (the same powf /
// sincosf, so the bits are those of rope2d_kernel): the per-element transcendental functions made the 4096 x 256 calls of the SAM2 memory attention
// ~10 us each, 112 per clip. `nblk` adjacent column blocks of width dim (q and k of a packed qkv row) share one launch.
__global__ void rope2d_table_kernel(float2* table, int dim, int gw, int gh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2, quarter = dim / 4;
  if (i >= gw * gh * half) return;
  const int pr = i % half, r = i / half;
  const int px = r % gw, py = (r / gw) % gh;
  const int fi = pr < quarter ? pr : pr - quarter;
  const float freq = 1.0f / powf(10000.0f, static_cast<float>(4 * fi) / dim);
  const float ang = (pr < quarter ? px : py) * freq;
  float sn, cs;
  sincosf(ang, &sn, &cs);
  table[i] = make_float2(cs, sn);
}

__global__ void rope2d_cached_kernel(void* x, int xdt, long long ld, int nrows, int dim, int nblk, const float2* __restrict__ table, int period) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int half = dim / 2;
  if (i >= static_cast<long long>(nrows) * half) return;
  const int pr = static_cast<int>(i % half);
  const long long r = i / half;
  const float2 cssn = __ldg(table + (r % period) * half + pr);
  for (int b = 0; b < nblk; ++b) {
    const long long e = r * ld + static_cast<long long>(b) * dim + 2 * pr;
    const float a0 = ld_(x, xdt, e), a1 = ld_(x, xdt, e + 1);
    st(x, xdt, e, a0 * cssn.x - a1 * cssn.y);
    st(x, xdt, e + 1, a0 * cssn.y + a1 * cssn.x);
  }
}

__global__ void rope_neox_kernel(void* qkv, int xdt, long long ld, const int* pos, int rows, int nheads, int hd, float theta) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  // one thread per (row, group of 8 heads, rotation pair): the angle (powf + sincosf, ~100 instructions) is computed once per 8 heads, and the
  // 16 loads of the group are issued before any arithmetic (a thread per element was compute-bound on the transcendental functions, a serial
  // loop over all heads latency-bound)
  constexpr int HG = 8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int half = hd / 2, groups = (nheads + HG - 1) / HG;
  if (i >= static_cast<long long>(rows) * groups * half) return;
  const int j = static_cast<int>(i % half);
  const int hg = static_cast<int>((i / half) % groups);
  const long long r = i / (static_cast<long long>(half) * groups);
  const float inv = powf(theta, -2.0f * j / hd);
  float sn, cs;
  sincosf(pos[r] * inv, &sn, &cs);
  const long long e = r * ld + static_cast<long long>(hg) * HG * hd + j;
  const int nh = min(HG, nheads - hg * HG);
  float a[HG], b[HG];
#pragma unroll
  for (int u = 0; u < HG; ++u) {
    if (u < nh) {
      a[u] = ld_(qkv, xdt, e + u * hd);
      b[u] = ld_(qkv, xdt, e + u * hd + half);
    }
  }
#pragma unroll
  for (int u = 0; u < HG; ++u) {
    if (u < nh) {
      st(qkv, xdt, e + u * hd, a[u] * cs - b[u] * sn);
      st(qkv, xdt, e + u * hd + half, b[u] * cs + a[u] * sn);
    }
  }
}

__global__ void embedding_kernel(const __nv_bfloat16* table, long long ldt, const long long* ids, void* out, int odt, long long ldo, int rows,
                                 int cols) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(rows) * cols) return;
  const int c = static_cast<int>(i % cols);
  const long long r = i / cols;
  st(out, odt, r * ldo + c, __bfloat162float(table[ids[r] * ldt + c]));
}

__global__ void row_scatter_kernel(const __nv_bfloat16* src, __nv_bfloat16* cache, long long ld, const int* pos, int col0, int cols) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cols) cache[static_cast<long long>(*pos) * ld + col0 + i] = src[col0 + i];
}

__global__ void rope_kv_append_kernel(__nv_bfloat16* qkv, __nv_bfloat16* cache, long long ld, const int* pos_dev, int H, int hd, float theta) {
  pdl_launch_dependents();
  pdl_wait();  // qkv comes from the preceding GEMV; pos_dev is advanced by the previous decode step
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = hd / 2, D = H * hd;
  const int pos = *pos_dev;
  __nv_bfloat16* crow = cache + static_cast<long long>(pos) * ld;
  if (i < 2 * H * half) {  // rotate pairs of q (heads 0..H-1) and k (heads H..2H-1)
    const int j = i % half, hh = i / half;
    const float inv = powf(theta, -2.0f * j / hd);
    float sn, cs;
    sincosf(pos * inv, &sn, &cs);
    __nv_bfloat16* p = qkv + hh * hd;
    const float a = __bfloat162float(p[j]), b = __bfloat162float(p[j + half]);
    const __nv_bfloat16 ra = __float2bfloat16(a * cs - b * sn), rb = __float2bfloat16(b * cs + a * sn);
    if (hh < H) {
      p[j] = ra;
      p[j + half] = rb;
    } else {
      crow[hh * hd + j] = ra;  // k goes straight to the cache (columns D..2D)
      crow[hh * hd + j + half] = rb;
    }
  } else if (i < 2 * H * half + D) {  // v: plain copy to columns 2D..3D
    const int c = 2 * D + (i - 2 * H * half);
    crow[c] = qkv[c];
  }
}

// one block per object: stability score of mask 0 (|{x > d}| / |{x > -d}|) and best-IoU index among masks 1..3
__global__ void multimask_select_kernel(const float* masks, const float* iou, int hw, float delta, float thresh, long long* out_idx) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const int b = blockIdx.x;
  const float* m0 = masks + static_cast<long long>(b) * 4 * hw;
  int ci = 0, cu = 0;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const float v = m0[i];
    ci += v > delta;
    cu += v > -delta;
  }
  __shared__ int si[256], su[256];
  si[threadIdx.x] = ci;
  su[threadIdx.x] = cu;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      si[threadIdx.x] += si[threadIdx.x + s];
      su[threadIdx.x] += su[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float stab = su[0] > 0 ? static_cast<float>(si[0]) / static_cast<float>(su[0]) : 1.0f;
    int best = 1;
    for (int j = 2; j < 4; ++j)
      if (iou[b * 4 + j] > iou[b * 4 + best]) best = j;  // first maximal index, like torch.argmax
    out_idx[b] = stab >= thresh ? 0 : best;
  }
}

// one block (256 threads) per row
__global__ void __launch_bounds__(1024) argmax_kernel(const void* x, int xdt, long long ldx, long long* out, int cols) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  // one block per row (first-index tie break, like torch.argmax). fp32 rows take 128-bit loads: the 32k-column lm_head row is ~8 independent
  // loads per thread instead of 125 dependent scalar ones (62 us -> a few us per decoded token).
  const int row = blockIdx.x;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  auto upd = [&](float v, int c) {
    if (v > best || (v == best && c < bi)) {
      best = v;
      bi = c;
    }
  };
  const long long base = static_cast<long long>(row) * ldx;
  if (xdt == VG_F32 && (cols & 3) == 0 && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const float4* xr = reinterpret_cast<const float4*>(static_cast<const float*>(x) + base);
    for (int c = threadIdx.x; c < (cols >> 2); c += blockDim.x) {
      const float4 v = xr[c];
      upd(v.x, 4 * c); upd(v.y, 4 * c + 1); upd(v.z, 4 * c + 2); upd(v.w, 4 * c + 3);
    }
  } else {
    for (int c = threadIdx.x; c < cols; c += blockDim.x) upd(ld(x, xdt, base + c), c);
  }
  // warp reduction, then one value per warp through shared memory
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    upd(ov, oi);
  }
  __shared__ float sv[32];
  __shared__ int si[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) {
    sv[warp] = best;
    si[warp] = bi;
  }
  __syncthreads();
  if (warp == 0) {
    best = lane < nw ? sv[lane] : -INFINITY;
    bi = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      upd(ov, oi);
    }
    if (lane == 0) out[row] = bi;
  }
}

// im2col for Conv2d(3x3, stride 2, pad 1) on a channels-last [n][h][w][cin] tensor -> bf16 [n*oh*ow][9*cin], column order (ky, kx, ci):
// one thread per 8 consecutive channels of one tap (16-byte store). Feeds the tensor-core GEMM (the direct conv kernels are
// parallelism-bound at 64 -> 256 channels: 238 us per launch for 1.2 GFLOP).
__global__ void im2col3x3s2_cl_kernel(const void* x, int xdt, __nv_bfloat16* out, int n, int h, int w, int cin) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const int oh = h / 2, ow = w / 2, c8 = cin >> 3;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(n) * oh * ow * 9 * c8;
  if (i >= total) return;
  const int cc = static_cast<int>(i % c8);
  long long t = i / c8;
  const int tap = static_cast<int>(t % 9);
  const long long pix = t / 9;
  const int ox = static_cast<int>(pix % ow), oy = static_cast<int>((pix / ow) % oh), img = static_cast<int>(pix / (static_cast<long long>(ow) * oh));
  const int iy = 2 * oy - 1 + tap / 3, ix = 2 * ox - 1 + tap % 3;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
    const long long src = ((static_cast<long long>(img) * h + iy) * w + ix) * cin + cc * 8;
    if (xdt == VG_F32) {
      const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(x) + src), b = *reinterpret_cast<const float4*>(static_cast<const float*>(x) + src + 4);
      v = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    } else {
      v = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(x) + src);
    }
  }
  *reinterpret_cast<uint4*>(out + pix * (9LL * cin) + tap * cin + cc * 8) = v;
}

}  // namespace vg

using namespace vg;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int vg_norm(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, long long ldy, const float* w, const float* b,
                       int rows, int cols, float eps, int rms, void* stream) {
  if (!x || !y || !w || rows <= 0 || cols <= 0) return set_error(VG_EINVAL, "vg_norm: bad argument");
  const int warps = 8;
  const unsigned grid = (rows + warps - 1) / warps;
  if (cols > 512 && cols <= 256 * 16 && rows < 2048) {  // wide rows, not enough of them to fill the chip with one warp each
    if (cols <= 256 * 4)
      launch_k(norm_block_kernel<4>, dim3(rows), dim3(256), 0, STREAM, x, x_dtype, ldx, y, y_dtype, ldy, w, b, cols, eps, rms);
    else if (cols <= 256 * 8)
      launch_k(norm_block_kernel<8>, dim3(rows), dim3(256), 0, STREAM, x, x_dtype, ldx, y, y_dtype, ldy, w, b, cols, eps, rms);
    else
      launch_k(norm_block_kernel<16>, dim3(rows), dim3(256), 0, STREAM, x, x_dtype, ldx, y, y_dtype, ldy, w, b, cols, eps, rms);
    count_launch();
    return check_launch("norm_block_kernel");
  }
  const int esz_y = y_dtype == VG_F32 ? 4 : 2;
  const bool vec_ok = x_dtype == VG_F32 && (cols % 4) == 0 && (ldx % 4) == 0 && (ldy % 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(y) & (esz_y * 4 - 1)) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                      (b == nullptr || (reinterpret_cast<uintptr_t>(b) & 15) == 0) && cols <= 128 * 12;
  if (vec_ok) {
    const void* xf = x;
    if (cols <= 128 * 2)
      launch_k(norm_vec_kernel<2>, dim3(grid), dim3(warps * 32), 0, STREAM, xf, ldx, y, y_dtype, ldy, w, b, rows, cols, eps, rms);
    else if (cols <= 128 * 5)
      launch_k(norm_vec_kernel<5>, dim3(grid), dim3(warps * 32), 0, STREAM, xf, ldx, y, y_dtype, ldy, w, b, rows, cols, eps, rms);
    else
      launch_k(norm_vec_kernel<12>, dim3(grid), dim3(warps * 32), 0, STREAM, xf, ldx, y, y_dtype, ldy, w, b, rows, cols, eps, rms);
    count_launch();
    return check_launch("norm_vec_kernel");
  }
  const bool vec_bf = x_dtype == VG_BF16 && (cols % 4) == 0 && (ldx % 4) == 0 && (ldy % 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0 &&
                      (reinterpret_cast<uintptr_t>(y) & (esz_y * 4 - 1)) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                      (b == nullptr || (reinterpret_cast<uintptr_t>(b) & 15) == 0) && cols <= 128 * 12;
  if (vec_bf) {  // bf16 rows (IV2 q/k RMSNorm on the packed qkv): 64-bit loads instead of 2-byte ones (30 us -> HBM-bound)
    if (cols <= 128 * 2)
      launch_k(norm_vec_kernel<2, true>, dim3(grid), dim3(warps * 32), 0, STREAM, x, ldx, y, y_dtype, ldy, w, b, rows, cols, eps, rms);
    else if (cols <= 128 * 5)
      launch_k(norm_vec_kernel<5, true>, dim3(grid), dim3(warps * 32), 0, STREAM, x, ldx, y, y_dtype, ldy, w, b, rows, cols, eps, rms);
    else
      launch_k(norm_vec_kernel<12, true>, dim3(grid), dim3(warps * 32), 0, STREAM, x, ldx, y, y_dtype, ldy, w, b, rows, cols, eps, rms);
    count_launch();
    return check_launch("norm_vec_kernel");
  }
  if (cols <= 32 * 4)
    launch_k(norm_kernel<4>, dim3(grid), dim3(warps * 32), 0, STREAM, x, x_dtype, ldx, y, y_dtype, ldy, w, b, rows, cols, eps, rms);
  else if (cols <= 32 * 12)
    launch_k(norm_kernel<12>, dim3(grid), dim3(warps * 32), 0, STREAM, x, x_dtype, ldx, y, y_dtype, ldy, w, b, rows, cols, eps, rms);
  else if (cols <= 32 * 48)
    launch_k(norm_kernel<48>, dim3(grid), dim3(warps * 32), 0, STREAM, x, x_dtype, ldx, y, y_dtype, ldy, w, b, rows, cols, eps, rms);
  else if (cols <= 32 * 100)
    launch_k(norm_kernel<100>, dim3(grid), dim3(warps * 32), 0, STREAM, x, x_dtype, ldx, y, y_dtype, ldy, w, b, rows, cols, eps, rms);
  else
    return set_error(VG_ENOSUP, "vg_norm: cols %d > 3200", cols);
  count_launch();
  return check_launch("norm_kernel");
}

extern "C" int vg_add(const void* a, int a_dtype, long long lda, int a_rows, const void* b, int b_dtype, long long ldb, int b_rows, float alpha, void* out,
                      int out_dtype, long long ldo, int rows, int cols, void* stream) {
  if (!a || !out || rows <= 0 || cols <= 0 || a_rows <= 0 || (b && b_rows <= 0)) return set_error(VG_EINVAL, "vg_add: bad argument");
  if (a_dtype == VG_F32 && out_dtype == VG_F32 && (!b || b_dtype == VG_F32) && cols % 4 == 0 && lda % 4 == 0 && ldo % 4 == 0 && (!b || ldb % 4 == 0) &&
      !((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(b)) & 15)) {
    launch_k(add_vec4_kernel, dim3(blocks_for(static_cast<long long>(rows) * (cols / 4), 256)), dim3(256), 0, STREAM, static_cast<const float*>(a), lda, a_rows,
             static_cast<const float*>(b), ldb, b_rows, alpha, static_cast<float*>(out), ldo, static_cast<long long>(rows), cols / 4);
    count_launch();
    return check_launch("add_vec4_kernel");
  }
  launch_k(add_kernel, dim3(blocks_for(static_cast<long long>(rows) * cols, 256)), dim3(256), 0, STREAM, a, a_dtype, lda, a_rows, b, b_dtype, ldb, b_rows, alpha, out,
                                                                                        out_dtype, ldo, rows, cols);
  count_launch();
  return check_launch("add_kernel");
}

extern "C" int vg_sum_parts(const float* parts, long long n, int nparts, const float* resid, float* out, void* stream) {
  if (!parts || !out || n <= 0 || nparts <= 0) return set_error(VG_EINVAL, "vg_sum_parts: bad argument");
  launch_k(sum_parts_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, STREAM, parts, n, nparts, resid, out);
  count_launch();
  return check_launch("sum_parts_kernel");
}

extern "C" int vg_pool2x2(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, long long ldy, int n, int h, int w, int c, int is_max,
                          void* stream) {
  if (!x || !y || (h & 1) || (w & 1)) return set_error(VG_EINVAL, "vg_pool2x2: bad argument (h, w must be even)");
  const bool al16 = !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15);
  if (x_dtype == y_dtype && x_dtype == VG_BF16 && al16 && c % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0) {
    launch_k(pool2x2_vec_kernel<true>, dim3(blocks_for(static_cast<long long>(n) * (h / 2) * (w / 2) * (c / 8), 256)), dim3(256), 0, STREAM, x, ldx, y, ldy, n, h, w, c,
             is_max);
    count_launch();
    return check_launch("pool2x2_vec_kernel");
  }
  if (x_dtype == y_dtype && x_dtype == VG_F32 && al16 && c % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0) {
    launch_k(pool2x2_vec_kernel<false>, dim3(blocks_for(static_cast<long long>(n) * (h / 2) * (w / 2) * (c / 4), 256)), dim3(256), 0, STREAM, x, ldx, y, ldy, n, h, w, c,
             is_max);
    count_launch();
    return check_launch("pool2x2_vec_kernel");
  }
  launch_k(pool2x2_kernel, dim3(blocks_for(static_cast<long long>(n) * (h / 2) * (w / 2) * c, 256)), dim3(256), 0, STREAM, x, x_dtype, ldx, y, y_dtype, ldy, n, h,
                                                                                                          w, c, is_max);
  count_launch();
  return check_launch("pool2x2_kernel");
}

extern "C" int vg_upsample2x_add(const float* lat, const float* top, float* out, int n, int h, int w, int c, void* stream) {
  if (!lat || !top || !out) return set_error(VG_EINVAL, "vg_upsample2x_add: null");
  launch_k(upsample2x_add_kernel, dim3(blocks_for(static_cast<long long>(n) * 4 * h * w * c, 256)), dim3(256), 0, STREAM, lat, top, out, n, h, w, c);
  count_launch();
  return check_launch("upsample2x_add_kernel");
}

extern "C" int vg_convt2x2_epilogue(const void* g, int g_dtype, const float* bias, const void* feat, int feat_dtype, const float* ln_w,
                                    const float* ln_b, float eps, void* out, int out_dtype, int n, int h, int w, int cout, int feat_bcast, void* stream) {
  if (!g || !bias || !out || cout > 128) return set_error(VG_EINVAL, "vg_convt2x2_epilogue: bad argument (cout <= 128)");
  const long long pix = static_cast<long long>(n) * 4 * h * w;
  launch_k(convt2x2_epilogue_kernel, dim3(blocks_for(pix, 8)), dim3(256), 0, STREAM, g, g_dtype, bias, feat, feat_dtype, ln_w, ln_b, eps, out, out_dtype, n, h, w,
                                                                    cout, feat_bcast);
  count_launch();
  return check_launch("convt2x2_epilogue_kernel");
}

extern "C" int vg_im2col(const void* x, int x_dtype, void* out, int out_dtype, int n, int cin, int h, int w, int k, int s, int p, int oh, int ow, int kpad,
                         void* stream) {
  if (!x || !out || kpad < cin * k * k) return set_error(VG_EINVAL, "vg_im2col: bad argument");
  if (out_dtype == VG_BF16 && kpad % 8 == 0 && !(reinterpret_cast<uintptr_t>(out) & 15)) {
    const dim3 grid(blocks_for(static_cast<long long>(n) * oh * ow * (kpad / 8), 256));
    __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(out);
    if (k == 7)
      launch_k(im2col_bf16x8_kernel<7>, grid, dim3(256), 0, STREAM, x, x_dtype, ob, n, cin, h, w, k, s, p, oh, ow, kpad);
    else if (k == 3)
      launch_k(im2col_bf16x8_kernel<3>, grid, dim3(256), 0, STREAM, x, x_dtype, ob, n, cin, h, w, k, s, p, oh, ow, kpad);
    else
      launch_k(im2col_bf16x8_kernel<0>, grid, dim3(256), 0, STREAM, x, x_dtype, ob, n, cin, h, w, k, s, p, oh, ow, kpad);
    count_launch();
    return check_launch("im2col_bf16x8_kernel");
  }
  launch_k(im2col_kernel, dim3(blocks_for(static_cast<long long>(n) * oh * ow * kpad, 256)), dim3(256), 0, STREAM, x, x_dtype, out, out_dtype, n,
                                                                                                  cin, h, w, k, s, p, oh, ow, kpad);
  count_launch();
  return check_launch("im2col_kernel");
}

extern "C" int vg_dwconv(const void* x, int x_dtype, const float* w, const float* bias, void* y, int y_dtype, int n, int h, int wd, int c, int k,
                         void* stream) {
  if (!x || !w || !y) return set_error(VG_EINVAL, "vg_dwconv: null");
  if (k == 7 && x_dtype == VG_F32 && c % 4 == 0 && wd % 4 == 0 && !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias)) & 15)) {
    const long long threads = static_cast<long long>(n) * h * (wd / 4) * (c / 4);
    launch_k(dwconv_vec_kernel<7, 4>, dim3(blocks_for(threads, 128)), dim3(128), 0, STREAM, static_cast<const float*>(x), w, bias, y, y_dtype, n, h, wd, c);
    count_launch();
    return check_launch("dwconv_vec_kernel");
  }
  launch_k(dwconv_kernel, dim3(blocks_for(static_cast<long long>(n) * h * wd * c, 256)), dim3(256), 0, STREAM, x, x_dtype, w, bias, y, y_dtype, n, h, wd, c, k);
  count_launch();
  return check_launch("dwconv_kernel");
}

extern "C" int vg_conv3x3s2_ln_gelu(const void* x, int x_dtype, const float* w, const float* bias, const float* ln_w, const float* ln_b,
                                    float eps, void* y, int y_dtype, int n, int h, int wd, int cin, int cout, void* stream) {
  if (!x || !w || !bias || !ln_w || !ln_b || !y || cout > 256 || (h & 1) || (wd & 1)) return set_error(VG_EINVAL, "vg_conv3x3s2_ln_gelu: bad argument");
  const long long pix = static_cast<long long>(n) * (h / 2) * (wd / 2);
  if (cin == 1 && cout == 4) {
    launch_k(conv3x3s2_ln_gelu_small_kernel<1, 4>, dim3(blocks_for(pix, 256)), dim3(256), 0, STREAM, x, x_dtype, w, bias, ln_w, ln_b, eps, y, y_dtype, n, h, wd);
    count_launch();
    return check_launch("conv3x3s2_ln_gelu_small_kernel");
  }
  if (cin == 4 && cout == 16) {
    launch_k(conv3x3s2_ln_gelu_small_kernel<4, 16>, dim3(blocks_for(pix, 256)), dim3(256), 0, STREAM, x, x_dtype, w, bias, ln_w, ln_b, eps, y, y_dtype, n, h, wd);
    count_launch();
    return check_launch("conv3x3s2_ln_gelu_small_kernel");
  }
  if ((wd / 2) % 8 == 0 && cout <= 256) {
    launch_k(conv3x3s2_ln_gelu_px_kernel<8>, dim3(blocks_for(pix / 8, 8)), dim3(256), 0, STREAM, x, x_dtype, w, bias, ln_w, ln_b, eps, y, y_dtype, n, h, wd, cin,
             cout);
    count_launch();
    return check_launch("conv3x3s2_ln_gelu_px_kernel");
  }
  launch_k(conv3x3s2_ln_gelu_kernel, dim3(blocks_for(pix, 8)), dim3(256), 0, STREAM, x, x_dtype, w, bias, ln_w, ln_b, eps, y, y_dtype, n, h, wd, cin, cout);
  count_launch();
  return check_launch("conv3x3s2_ln_gelu_kernel");
}

extern "C" int vg_bilinear(const float* x, void* y, int n, int ih, int iw, int oh, int ow, int mode, float scale, float bias, void* stream) {
  if (!x || !y) return set_error(VG_EINVAL, "vg_bilinear: null");
  if (mode == 4) {
    const long long bytes = static_cast<long long>(n) * ((static_cast<long long>(oh) * ow + 7) / 8);
    launch_k(bilinear_pack_kernel, dim3(blocks_for(bytes, 256)), dim3(256), 0, STREAM, x, static_cast<uint8_t*>(y), n, ih, iw, oh, ow);
    count_launch();
    return check_launch("bilinear_pack_kernel");
  }
  launch_k(bilinear_kernel, dim3(blocks_for(static_cast<long long>(n) * oh * ow, 256)), dim3(256), 0, STREAM, x, y, n, ih, iw, oh, ow, mode, scale, bias);
  count_launch();
  return check_launch("bilinear_kernel");
}

extern "C" int vg_rope2d(void* x, int x_dtype, long long ld, int nrows, int dim, int gw, int gh, void* stream) {
  if (!x || dim % 4) return set_error(VG_EINVAL, "vg_rope2d: dim must be a multiple of 4");
  launch_k(rope2d_kernel, dim3(blocks_for(static_cast<long long>(nrows) * dim / 2, 256)), dim3(256), 0, STREAM, x, x_dtype, ld, nrows, dim, gw, gh);
  count_launch();
  return check_launch("rope2d_kernel");
}

extern "C" int vg_rope2d_table(float* table, int dim, int gw, int gh, void* stream) {
  if (!table || dim % 4) return set_error(VG_EINVAL, "vg_rope2d_table: dim must be a multiple of 4");
  launch_k(rope2d_table_kernel, dim3(blocks_for(static_cast<long long>(gw) * gh * dim / 2, 256)), dim3(256), 0, STREAM, reinterpret_cast<float2*>(table), dim, gw, gh);
  count_launch();
  return check_launch("rope2d_table_kernel");
}

extern "C" int vg_rope2d_cached(void* x, int x_dtype, long long ld, int nrows, int dim, int nblk, const float* table, int period, void* stream) {
  if (!x || !table || dim % 4 || nblk < 1 || period < 1) return set_error(VG_EINVAL, "vg_rope2d_cached: bad argument");
  launch_k(rope2d_cached_kernel, dim3(blocks_for(static_cast<long long>(nrows) * dim / 2, 256)), dim3(256), 0, STREAM, x, x_dtype, ld, nrows, dim, nblk,
           reinterpret_cast<const float2*>(table), period);
  count_launch();
  return check_launch("rope2d_cached_kernel");
}

extern "C" int vg_rope_neox(void* qkv, int x_dtype, long long ld, const int* pos, int rows, int n_heads_total, int head_dim, float theta, void* stream) {
  if (!qkv || !pos || head_dim % 2) return set_error(VG_EINVAL, "vg_rope_neox: bad argument");
  launch_k(rope_neox_kernel, dim3(blocks_for(static_cast<long long>(rows) * ((n_heads_total + 7) / 8) * (head_dim / 2), 128)), dim3(128), 0, STREAM, 
      qkv, x_dtype, ld, pos, rows, n_heads_total, head_dim, theta);
  count_launch();
  return check_launch("rope_neox_kernel");
}

extern "C" int vg_embedding(const void* table, long long ldt, const long long* ids, void* out, int out_dtype, long long ldo, int rows, int cols,
                            void* stream) {
  if (!table || !ids || !out) return set_error(VG_EINVAL, "vg_embedding: null");
  launch_k(embedding_kernel, dim3(blocks_for(static_cast<long long>(rows) * cols, 256)), dim3(256), 0, STREAM, static_cast<const __nv_bfloat16*>(table), ldt, ids,
                                                                                              out, out_dtype, ldo, rows, cols);
  count_launch();
  return check_launch("embedding_kernel");
}

extern "C" int vg_argmax(const void* x, int x_dtype, long long ldx, long long* out, int rows, int cols, void* stream) {
  if (!x || !out) return set_error(VG_EINVAL, "vg_argmax: null");
  launch_k(argmax_kernel, dim3(rows), dim3(cols >= 8192 ? 1024 : 256), 0, STREAM, x, x_dtype, ldx, out, cols);
  count_launch();
  return check_launch("argmax_kernel");
}

extern "C" int vg_row_scatter(const void* src, void* cache, long long ld_cache, const int* pos_dev, int col0, int cols, void* stream) {
  if (!src || !cache || !pos_dev || cols <= 0) return set_error(VG_EINVAL, "vg_row_scatter: bad argument");
  launch_k(row_scatter_kernel, dim3(blocks_for(cols, 256)), dim3(256), 0, STREAM, static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(cache), ld_cache,
                                                                pos_dev, col0, cols);
  count_launch();
  return check_launch("row_scatter_kernel");
}

extern "C" int vg_rope_kv_append(void* qkv, void* cache, long long ld_cache, const int* pos_dev, int n_heads, int head_dim, float theta,
                                 void* stream) {
  if (!qkv || !cache || !pos_dev || head_dim % 2) return set_error(VG_EINVAL, "vg_rope_kv_append: bad argument");
  const int n = n_heads * head_dim * 2;  // q+k pairs (H*hd) + v elements (H*hd)
  const cudaError_t e = launch_pdl(rope_kv_append_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, STREAM, static_cast<__nv_bfloat16*>(qkv),
                                   static_cast<__nv_bfloat16*>(cache), ld_cache, pos_dev, n_heads, head_dim, theta);
  if (e != cudaSuccess) return set_error(VG_ECUDA, "rope_kv_append launch: %s", cudaGetErrorString(e));
  count_launch();
  return check_launch("rope_kv_append_kernel");
}

extern "C" int vg_multimask_select(const float* masks, const float* iou, int n_obj, int hw, float delta, float thresh, long long* out_idx, void* stream) {
  if (!masks || !iou || !out_idx || n_obj <= 0) return set_error(VG_EINVAL, "vg_multimask_select: bad argument");
  launch_k(multimask_select_kernel, dim3(n_obj), dim3(256), 0, STREAM, masks, iou, hw, delta, thresh, out_idx);
  count_launch();
  return check_launch("multimask_select_kernel");
}

extern "C" int vg_im2col3x3s2_cl(const void* x, int x_dtype, void* out, int n, int h, int w, int cin, void* stream) {
  if (!x || !out || n <= 0 || (cin % 8) || (h % 2) || (w % 2)) return set_error(VG_EINVAL, "vg_im2col3x3s2_cl: cin must be a multiple of 8, h and w even");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return set_error(VG_EINVAL, "vg_im2col3x3s2_cl: 16-byte alignment");
  const long long total = static_cast<long long>(n) * (h / 2) * (w / 2) * 9 * (cin / 8);
  launch_k(im2col3x3s2_cl_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM, x, x_dtype, static_cast<__nv_bfloat16*>(out), n, h, w, cin);
  count_launch();
  return check_launch("im2col3x3s2_cl_kernel");
}

extern "C" int vg_layernorm_gelu(const float* x, long long ldx, void* y, int y_dtype, long long ldy, const float* w, const float* b, int rows, int cols,
                                 float eps, void* stream) {
  if (!x || !y || !w || !b || rows <= 0 || cols <= 0) return set_error(VG_EINVAL, "vg_layernorm_gelu: bad argument");
  const int esz_y = y_dtype == VG_F32 ? 4 : 2;
  if ((cols % 4) || (ldx % 4) || (ldy % 4) || cols > 128 * 5 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & (esz_y * 4 - 1)) ||
      (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(b) & 15))
    return set_error(VG_ENOSUP, "vg_layernorm_gelu: needs cols %% 4 == 0, cols <= 640 and 16-byte aligned rows");
  const int warps = 8;
  const unsigned grid = (rows + warps - 1) / warps;
  if (cols <= 128 * 2)
    launch_k(norm_vec_kernel<2, false, true>, dim3(grid), dim3(warps * 32), 0, STREAM, static_cast<const void*>(x), ldx, y, y_dtype, ldy, w, b, rows, cols, eps, 0);
  else
    launch_k(norm_vec_kernel<5, false, true>, dim3(grid), dim3(warps * 32), 0, STREAM, static_cast<const void*>(x), ldx, y, y_dtype, ldy, w, b, rows, cols, eps, 0);
  count_launch();
  return check_launch("norm_vec_kernel");
}
