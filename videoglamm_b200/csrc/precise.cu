// fp32-grade arithmetic on the bf16 tensor cores and an exact fp32 attention, for the stages whose rounding the mask logits are sensitive to
// (SAM2 mask decoder: hypernetwork product, two-way attention) and for the engine-wide verification mode.
//   vg_split3        x (fp32) -> bf16 [hi | lo | hi] (A side) or [hi | hi | lo] (B side) along K, so that ONE ordinary tcgen05 GEMM with K' = 3K computes
//                    A_hi B_hi + A_lo B_hi + A_hi B_lo with fp32 accumulation in TMEM (error ~2^-17 relative: the dropped lo*lo term)
//   vg_attention_f32 softmax(scale Q K^T) V on fp32 operands, fp32 FMA accumulation, expf: the mask decoder's small attentions (9 tokens x 4096 image
//                    tokens, 8 heads x 16..32) in the product path; any shape (causal / sliding window / Hiera windows, D <= 256) in verification mode
#include "common.h"
#include <cuda_bf16.h>

#include "pdl.cuh"

namespace vg {

__global__ void split3_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out, long long ldo, long long rows, int cols, int kp,
                              int b_side) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= rows * kp) return;
  const long long r = i / kp;
  const int c = static_cast<int>(i - r * kp);
  const float v = c < cols ? x[r * ldx + c] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16(v);
  const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
  __nv_bfloat16* o = out + r * ldo + c;
  o[0] = hi;
  o[kp] = b_side ? hi : lo;
  o[2 * kp] = b_side ? lo : hi;
}

struct AttnF32Params {
  const float *q, *k, *v;
  float* o;
  long long q_bs, q_hs, q_rs, k_bs, k_hs, k_rs, v_bs, v_hs, v_rs, o_bs, o_hs, o_rs;
  int B, H, Nq, Nk, D;
  float scale;
  int causal, window;
  int q_win, q_gridw, q_nwx, q_nwy, k_win, k_gridw, k_nwx, k_nwy;
  float* ws;  // split partials [split][b][h][row][D + 2] = (m, l, o[D])
  int nsplit;
};

__device__ __forceinline__ long long tok_row_f(int b, int t, int win, int gridw, int nwx, int nwy) {
  if (win == 0) return t;
  const int per_img = nwx * nwy;
  const int img = b / per_img, w = b - img * per_img;
  const int wy = w / nwx, wx = w - wy * nwx;
  const int ty = t / win, tx = t - ty * win;
  return static_cast<long long>(img) * (static_cast<long long>(nwy) * win * gridw) + static_cast<long long>(wy * win + ty) * gridw + wx * win + tx;
}

// 256 threads: 16 query rows x 64-key tiles. Shared memory: Q [16][D+1], K [64][D+1], V [64][D], S/P [16][64], alpha [16].
constexpr int F32_BQ = 16, F32_BK = 64;

__global__ void __launch_bounds__(256) attn_f32_kernel(const AttnF32Params p) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float sm_f32[];
  const int D = p.D, LD = D + 1;
  float* sQ = sm_f32;
  float* sK = sQ + F32_BQ * LD;
  float* sV = sK + F32_BK * LD;
  float* sS = sV + F32_BK * D;
  float* sAlpha = sS + F32_BQ * F32_BK;
  float* sM = sAlpha + F32_BQ;
  float* sL = sM + F32_BQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q_tiles = (p.Nq + F32_BQ - 1) / F32_BQ;
  const int split = blockIdx.x / q_tiles;
  const int q0 = (blockIdx.x - split * q_tiles) * F32_BQ, h = blockIdx.y, b = blockIdx.z;
  const int off = p.Nk - p.Nq;
  int k_begin = 0, k_end = p.Nk;
  if (p.causal) k_end = min(p.Nk, q0 + F32_BQ + off);
  if (p.window > 0) k_begin = max(0, q0 + off - p.window);
  if (p.nsplit > 1) {
    const int per = ((p.Nk + p.nsplit - 1) / p.nsplit + F32_BK - 1) / F32_BK * F32_BK;
    k_begin = max(k_begin, split * per);
    k_end = min(k_end, split * per + per);
  }
  const float* qb = p.q + (p.q_win ? 0 : b * p.q_bs) + h * p.q_hs;
  const float* kb = p.k + (p.k_win ? 0 : b * p.k_bs) + h * p.k_hs;
  const float* vb = p.v + (p.k_win ? 0 : b * p.v_bs) + h * p.v_hs;
  for (int i = tid; i < F32_BQ * D; i += 256) {
    const int r = i / D, d = i - r * D;
    sQ[r * LD + d] = (q0 + r < p.Nq) ? qb[tok_row_f(b, q0 + r, p.q_win, p.q_gridw, p.q_nwx, p.q_nwy) * p.q_rs + d] : 0.f;
  }
  if (tid < F32_BQ) {
    sM[tid] = -INFINITY;
    sL[tid] = 0.f;
  }
  // O accumulators: thread -> column c = tid % DP2, rows rg + i * nrg (DP2 = D rounded up to a power of two, nrg = 256 / DP2 row groups)
  int DP2 = 8;
  while (DP2 < D) DP2 <<= 1;
  const int nrg = 256 / DP2, c = tid % DP2, rg = tid / DP2;
  const int nacc = (F32_BQ + nrg - 1) / nrg;  // <= 16
  float acc[F32_BQ];
#pragma unroll
  for (int i = 0; i < F32_BQ; ++i) acc[i] = 0.f;
  const int sj = tid & 63, sr0 = (tid >> 6) * 4;  // S entries of this thread: key sj, rows sr0 .. sr0 + 3
  for (int k0 = k_begin; k0 < k_end; k0 += F32_BK) {
    __syncthreads();
    for (int i = tid; i < F32_BK * D; i += 256) {
      const int r = i / D, d = i - r * D;
      const bool ok = k0 + r < k_end;
      const long long row = ok ? tok_row_f(b, k0 + r, p.k_win, p.k_gridw, p.k_nwx, p.k_nwy) : 0;
      sK[r * LD + d] = ok ? kb[row * p.k_rs + d] : 0.f;
      sV[r * D + d] = ok ? vb[row * p.v_rs + d] : 0.f;
    }
    __syncthreads();
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const float* kr = sK + sj * LD;
    for (int d = 0; d < D; ++d) {
      const float kv = kr[d];
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i] = fmaf(sQ[(sr0 + i) * LD + d], kv, s[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + sr0 + i, kj = k0 + sj;
      bool ok = kj < k_end && qi < p.Nq;
      if (p.causal && kj > qi + off) ok = false;
      if (p.window > 0 && kj < qi + off - p.window) ok = false;
      sS[(sr0 + i) * F32_BK + sj] = ok ? s[i] * p.scale : -INFINITY;
    }
    __syncthreads();
    // online softmax: warp w owns rows 2w, 2w + 1
    for (int rr = 0; rr < 2; ++rr) {
      const int r = warp * 2 + rr;
      const float a0 = sS[r * F32_BK + lane], a1 = sS[r * F32_BK + 32 + lane];
      float mx = fmaxf(a0, a1);
      for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_old = sM[r], m_new = fmaxf(m_old, mx);
      const float p0 = (m_new == -INFINITY) ? 0.f : expf(a0 - m_new), p1 = (m_new == -INFINITY) ? 0.f : expf(a1 - m_new);
      float sum = p0 + p1;
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sS[r * F32_BK + lane] = p0;
      sS[r * F32_BK + 32 + lane] = p1;
      if (lane == 0) {
        const float alpha = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
        sAlpha[r] = alpha;
        sL[r] = sL[r] * alpha + sum;
        sM[r] = m_new;
      }
    }
    __syncthreads();
    if (c < D) {
#pragma unroll
      for (int i = 0; i < F32_BQ; ++i)
        if (i < nacc && rg + i * nrg < F32_BQ) acc[i] *= sAlpha[rg + i * nrg];
      for (int j = 0; j < F32_BK; ++j) {
        const float vv = sV[j * D + c];
#pragma unroll
        for (int i = 0; i < F32_BQ; ++i)
          if (i < nacc && rg + i * nrg < F32_BQ) acc[i] = fmaf(sS[(rg + i * nrg) * F32_BK + j], vv, acc[i]);
      }
    }
  }
  __syncthreads();
  if (c < D) {
#pragma unroll
    for (int i = 0; i < F32_BQ; ++i) {
      const int r = rg + i * nrg;
      if (i >= nacc || r >= F32_BQ || q0 + r >= p.Nq) continue;
      if (p.nsplit > 1) {
        float* wp = p.ws + ((((static_cast<long long>(split) * p.B + b) * p.H + h) * p.Nq) + q0 + r) * (D + 2);
        if (c == 0) {
          wp[0] = sM[r];
          wp[1] = sL[r];
        }
        wp[2 + c] = acc[i];
      } else {
        const float l = sL[r];
        const long long orow = tok_row_f(b, q0 + r, p.q_win, p.q_gridw, p.q_nwx, p.q_nwy);
        p.o[(p.q_win ? 0 : b * p.o_bs) + h * p.o_hs + orow * p.o_rs + c] = l > 0.f ? acc[i] / l : 0.f;
      }
    }
  }
}

__global__ void attn_f32_merge_kernel(const AttnF32Params p) {
  pdl_wait();
  pdl_launch_dependents();
  const long long gw = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long rows = static_cast<long long>(p.B) * p.H * p.Nq;
  if (gw >= rows) return;
  const int row = static_cast<int>(gw % p.Nq);
  const int h = static_cast<int>((gw / p.Nq) % p.H);
  const int b = static_cast<int>(gw / (static_cast<long long>(p.Nq) * p.H));
  const int W = p.D + 2;
  float m_all = -INFINITY;
  for (int s = 0; s < p.nsplit; ++s) m_all = fmaxf(m_all, p.ws[(static_cast<long long>(s) * rows + gw) * W]);
  float l_all = 0.f;
  for (int s = 0; s < p.nsplit; ++s) {
    const float* wp = p.ws + (static_cast<long long>(s) * rows + gw) * W;
    l_all += wp[1] * ((wp[0] == -INFINITY) ? 0.f : expf(wp[0] - m_all));
  }
  const float inv = l_all > 0.f ? 1.0f / l_all : 0.f;
  float* orow = p.o + b * p.o_bs + h * p.o_hs + static_cast<long long>(row) * p.o_rs;
  for (int d = lane; d < p.D; d += 32) {
    float acc = 0.f;
    for (int s = 0; s < p.nsplit; ++s) {
      const float* wp = p.ws + (static_cast<long long>(s) * rows + gw) * W;
      acc += wp[2 + d] * ((wp[0] == -INFINITY) ? 0.f : expf(wp[0] - m_all));
    }
    orow[d] = acc * inv;
  }
}

}  // namespace vg

using namespace vg;

extern "C" int vg_split3(const float* x, long long ldx, void* out, long long ldo, long long rows, int cols, int kp, int b_side, void* stream) {
  if (!x || !out || rows <= 0 || cols <= 0 || kp < cols || (kp % 8) || ldo < 3LL * kp) return set_error(VG_EINVAL, "vg_split3: bad argument (kp % 8 == 0, ldo >= 3 kp)");
  const long long n = rows * kp;
  launch_k(split3_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, ldx, static_cast<__nv_bfloat16*>(out), ldo,
           rows, cols, kp, b_side);
  count_launch();
  return check_launch("split3_kernel");
}

extern "C" int vg_attention_f32(const VgAttention* a, void* stream) {
  if (!a || !a->q || !a->k || !a->v || !a->o) return set_error(VG_EINVAL, "vg_attention_f32: null pointer");
  if (a->B <= 0 || a->H <= 0 || a->Nq <= 0 || a->Nk <= 0 || a->D <= 0 || a->D > 256) return set_error(VG_EINVAL, "vg_attention_f32: bad dims (D <= 256)");
  if (a->nk_dev != nullptr || a->fused_rope_theta != 0.f) return set_error(VG_ENOSUP, "vg_attention_f32: decode-graph features are bf16-path only");
  AttnF32Params p;
  p.q = static_cast<const float*>(a->q); p.k = static_cast<const float*>(a->k); p.v = static_cast<const float*>(a->v); p.o = static_cast<float*>(a->o);
  p.q_bs = a->q_bs; p.q_hs = a->q_hs; p.q_rs = a->q_rs; p.k_bs = a->k_bs; p.k_hs = a->k_hs; p.k_rs = a->k_rs;
  p.v_bs = a->v_bs; p.v_hs = a->v_hs; p.v_rs = a->v_rs; p.o_bs = a->o_bs; p.o_hs = a->o_hs; p.o_rs = a->o_rs;
  p.B = a->B; p.H = a->H; p.Nq = a->Nq; p.Nk = a->Nk; p.D = a->D;
  p.scale = a->scale; p.causal = a->causal; p.window = a->window;
  p.q_win = a->q_win; p.q_gridw = a->q_gridw; p.q_nwx = a->q_nwx; p.q_nwy = a->q_nwy;
  p.k_win = a->k_win; p.k_gridw = a->k_gridw; p.k_nwx = a->k_nwx; p.k_nwy = a->k_nwy;
  p.ws = nullptr; p.nsplit = 1;
  const int q_tiles = (a->Nq + F32_BQ - 1) / F32_BQ;
  const long long ctas = static_cast<long long>(q_tiles) * a->H * a->B;
  if (a->workspace != nullptr && !a->causal && a->window == 0 && a->q_win == 0 && a->k_win == 0 && ctas < 148 && a->Nk >= 4 * F32_BK) {
    int ns = static_cast<int>((2 * 148 + ctas - 1) / ctas);
    if (ns > 16) ns = 16;
    if (ns > a->Nk / (2 * F32_BK)) ns = a->Nk / (2 * F32_BK);
    const long long need = static_cast<long long>(ns) * a->B * a->H * a->Nq * (a->D + 2) * 4;
    if (ns > 1 && need <= a->workspace_bytes) {
      p.nsplit = ns;
      p.ws = static_cast<float*>(a->workspace);
    }
  }
  const int D = a->D;
  const int smem = (F32_BQ * (D + 1) + F32_BK * (D + 1) + F32_BK * D + F32_BQ * F32_BK + 3 * F32_BQ) * 4;
  static int attr_smem = 0;
  if (smem > 48 * 1024 && smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_error(VG_ECUDA, "attn_f32 smem attr: %s", cudaGetErrorString(e));
    attr_smem = 200 * 1024;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid(q_tiles * p.nsplit, a->H, a->B);
  launch_k(attn_f32_kernel, dim3(grid), dim3(256), smem, s, p);
  count_launch();
  if (p.nsplit > 1) {
    const long long rows = static_cast<long long>(p.B) * p.H * p.Nq;
    launch_k(attn_f32_merge_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, s, p);
    count_launch();
  }
  return check_launch("attn_f32_kernel");
}
