// Fused SAM2 mask decoder: PromptEncoder dense part + TwoWayTransformer (2 layers + final token->image attention) + both transposed convolutions +
// hypernetwork / IoU / object-score MLPs + the hyper_in x upscaled product, ONE launch per frame for all objects.
// Reference: sam2/modeling/sam/mask_decoder.py:167-245 (predict_masks), sam/transformer.py:63-115 (TwoWayTransformer), :160-193 (block), :233-260.
//
// Schedule (proven exact on the CPU by oracle/mask_decoder_rowblock.py): the HW image tokens of one object are split into row blocks of 32;
// thread block g of object b keeps its 32 rows (the "keys" stream of the two-way transformer) resident in shared memory for the whole decoder.
//   row-local work (no communication): K/V projections for the token->image attentions, the whole image->token attention (q projection, softmax
//     over the <= 12 tokens, out projection, residual, norm4), both ConvTranspose2d(k2,s2) (+ high-res features, LayerNorm2d, GELU) and the final
//     hyper_in x upscaled product;
//   token->image attention: every block writes flash-style partials (max, sum, weighted V) of its 32 rows, merged after a grid barrier;
//   the <= 12-token chain (self-attention, MLP 2048, projections, hypernetwork MLPs): each dependent GEMV is distributed over the blocks of the
//     object by output column and separated by grid barriers (an atomic counter per object in L2; all blocks of an object are co-resident).
// Arithmetic: fp32 FMA throughout (the mask logits need fp32-grade contractions, scripts/precision_study.py); exact erf GELU, expf.
// Row-local contractions read TRANSPOSED weights [K][N] (coalesced over n); token-side GEMVs read nn.Linear-layout rows [N][K] (coalesced over k).
#include <cuda_bf16.h>

#include "common.h"
#include "pdl.cuh"
#include "ptx.cuh"

namespace vg {

constexpr int MD_C = 256, MD_CI = 128, MD_H = 8, MD_RB = 32, MD_NT = 12, MD_MLP = 2048, MD_LDK = 132, MD_THREADS = 256;
constexpr int MD_TB = MD_NT * 768;  // token buffer (floats)

// per-object scratch layout (floats)
struct MdOff {
  int Q, QPRE, QKV, TQ, ATT, HID, TK, TV, HS, H1, H2, HYP, PART, total;
  __host__ __device__ explicit MdOff(int G) {
    Q = 0;
    QPRE = Q + MD_NT * MD_C;
    QKV = QPRE + MD_NT * MD_C;
    TQ = QKV + MD_NT * 3 * MD_C;
    ATT = TQ + MD_NT * MD_CI;
    HID = ATT + MD_NT * MD_C;          // self-attention output uses 256 columns, token->image 128
    TK = HID + MD_NT * MD_MLP;
    TV = TK + MD_NT * MD_CI;
    HS = TV + MD_NT * MD_CI;
    H1 = HS + MD_NT * MD_C;
    H2 = H1 + 6 * MD_C;
    HYP = H2 + 6 * MD_C;
    PART = HYP + 4 * 32;
    total = PART + G * MD_NT * MD_H * 18;
  }
};

struct MdParams {
  const float* pix; long long pix_obj_stride;
  const float *pe, *no_mask, *s0, *s1, *tokens;
  int B, Nt, fs, G;
  float *masks, *iou, *mask_tokens, *obj;
  float* ws; long long ws_obj_stride;
  unsigned* bar;
  VgMaskDecoderWeights w;
};

__device__ __forceinline__ float md_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// grid barrier over the G blocks of one object: counter in global memory, monotonically increasing (target = G * barrier index)
__device__ __forceinline__ void md_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned spins = 0, v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v < target) {
        __nanosleep(32);
        if (++spins > (1u << 25)) __trap();  // ~2 s: a block of this object never became resident (misconfigured launch) -> fail loudly, do not hang
      }
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// ---- token-side helpers -------------------------------------------------------------------------------------------------------------------
// rows [0, NT) of a global [NT][ld] matrix -> shared [NT][K] (L2 reads: the producer was another block)
__device__ __forceinline__ void md_load_rows(float* dst, const float* src, int ld, int K, int NT) {
  for (int i = threadIdx.x; i < NT * K; i += MD_THREADS) {
    const int t = i / K, k = i - t * K;
    dst[i] = __ldcg(src + t * ld + k);
  }
}

// dst[t][:] = LayerNorm(src[t][:]) over C = 256 channels (one warp per token, fp32 statistics); src global (L2), dst shared
__device__ __forceinline__ void md_load_rows_ln(float* dst, const float* src, int NT, const float* w, const float* b, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < NT; t += MD_THREADS / 32) {
    float v[8], s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = __ldcg(src + t * MD_C + lane + 32 * i);
      s += v[i];
    }
    const float mean = warp_sum(s) * (1.0f / MD_C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) q += (v[i] - mean) * (v[i] - mean);
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / MD_C) + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      dst[t * MD_C + c] = (v[i] - mean) * rstd * w[c] + b[c];
    }
  }
}

// Distributed linear: this block computes output columns n = blk + G * j of y[t][n] = act(bias[n] + x[t][:] . W[n][:]) (+ resid[t][n]); x in
// shared memory [NT][ldx], W global row-major [N][K] (nn.Linear layout), one warp per column, lanes stride over k. accumulate: y += (K chunks).
__device__ __forceinline__ void md_tok_linear(const float* xs, int ldx, int K, const float* W, int ldw, const float* bias, int N, float* y, int ldy, int NT,
                                              bool relu, const float* resid, int ldr, bool accumulate, int blk, int G) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = blk + G * warp; n < N; n += G * (MD_THREADS / 32)) {
    float acc[MD_NT];
#pragma unroll
    for (int t = 0; t < MD_NT; ++t) acc[t] = 0.f;
    const float* wr = W + static_cast<long long>(n) * ldw;
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int t = 0; t < MD_NT; ++t)
        if (t < NT) acc[t] = fmaf(xs[t * ldx + k], wv, acc[t]);
    }
#pragma unroll
    for (int t = 0; t < MD_NT; ++t)
      if (t < NT) acc[t] = warp_sum(acc[t]);
    if (lane < NT) {
      float v = 0.f;
#pragma unroll
      for (int t = 0; t < MD_NT; ++t)
        if (t == lane) v = acc[t];
      if (accumulate) {
        y[lane * ldy + n] += v;
      } else {
        v += bias ? bias[n] : 0.f;
        if (relu) v = fmaxf(v, 0.f);
        if (resid) v += __ldcg(resid + lane * ldr + n);
        y[lane * ldy + n] = v;
      }
    }
  }
}

// ---- image-side helpers -------------------------------------------------------------------------------------------------------------------
// out[r][n] = bias[n] + sum_k A[r][k] Wt[k][n] (+ add[r * ld_add + n]) for the block's 32 rows; A shared [32][lda], Wt global TRANSPOSED [K][N],
// out shared [32][ldo]. Thread tile: RPT rows x 4 columns, k unrolled by 4 (float4 loads of A along k and of Wt along n).
template <int N>
__device__ __forceinline__ void md_img_gemm(const float* A, int lda, int K, const float* __restrict__ Wt, const float* __restrict__ bias, const float* add,
                                            long long ld_add, float* out, int ldo) {
  constexpr int CG = N / 4, RG = MD_THREADS / CG, RPT = MD_RB / RG;
  const int cg = threadIdx.x % CG, rg = threadIdx.x / CG;
  float acc[RPT][4];
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float* wp = Wt + 4 * cg;
  for (int k = 0; k < K; k += 4) {
    float4 w[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) w[kk] = __ldg(reinterpret_cast<const float4*>(wp + static_cast<long long>(k + kk) * N));
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const float4 a = *reinterpret_cast<const float4*>(A + (rg * RPT + i) * lda + k);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        acc[i][0] = fmaf(av[kk], w[kk].x, acc[i][0]);
        acc[i][1] = fmaf(av[kk], w[kk].y, acc[i][1]);
        acc[i][2] = fmaf(av[kk], w[kk].z, acc[i][2]);
        acc[i][3] = fmaf(av[kk], w[kk].w, acc[i][3]);
      }
    }
  }
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + 4 * cg));
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int r = rg * RPT + i;
    float4 o = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
    if (add) {
      const float4 ad = __ldg(reinterpret_cast<const float4*>(add + r * ld_add + 4 * cg));
      o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
    }
    *reinterpret_cast<float4*>(out + r * ldo + 4 * cg) = o;
  }
}

// flash-style partials of the token->image attention over this block's 32 rows: one warp per (token, head) pair, lane = row.
// tq shared [NT][128]; Ks, Vs shared [32][MD_LDK]; part global [pairs][18] = (max, sum, weighted V[16])
__device__ __forceinline__ void md_t2i_partials(const float* tq, const float* Ks, const float* Vs, float* part, int NT) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale = 0.25f;  // head dim 16
  for (int pair = warp; pair < NT * MD_H; pair += MD_THREADS / 32) {
    const int t = pair / MD_H, h = pair - t * MD_H;
    const float* q = tq + t * MD_CI + h * 16;
    const float* k = Ks + lane * MD_LDK + h * 16;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 16; d += 4) {
      const float4 kv = *reinterpret_cast<const float4*>(k + d);
      s = fmaf(q[d], kv.x, s); s = fmaf(q[d + 1], kv.y, s); s = fmaf(q[d + 2], kv.z, s); s = fmaf(q[d + 3], kv.w, s);
    }
    s *= scale;
    const float m = warp_max(s);
    const float p = expf(s - m);
    const float l = warp_sum(p);
    // o[d] = sum_r p_r V[r][h*16 + d]: lane (d = lane & 15, half = lane >> 4) sums the rows of its half, then the halves are added
    const int d = lane & 15, half = lane >> 4;
    float o = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = half * 16 + i;
      const float pr = __shfl_sync(0xffffffffu, p, r);
      o = fmaf(pr, Vs[r * MD_LDK + h * 16 + d], o);
    }
    o += __shfl_xor_sync(0xffffffffu, o, 16);
    float* pp = part + pair * 18;
    if (lane == 0) {
      pp[0] = m;
      pp[1] = l;
    }
    if (lane < 16) pp[2 + lane] = o;
  }
}

// merge the G partials of the pairs assigned to this block -> att[t][h*16 + d] (global)
__device__ __forceinline__ void md_t2i_merge(const float* part_base, float* att, int NT, int blk, int G) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pairs = NT * MD_H;
  for (int pair = blk + G * warp; pair < pairs; pair += G * (MD_THREADS / 32)) {
    float m = -INFINITY;
    for (int g = lane; g < G; g += 32) m = fmaxf(m, __ldcg(part_base + (static_cast<long long>(g) * MD_NT * MD_H + pair) * 18));
    m = warp_max(m);
    float l = 0.f, o[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d] = 0.f;
    for (int g = lane; g < G; g += 32) {
      const float* pp = part_base + (static_cast<long long>(g) * MD_NT * MD_H + pair) * 18;
      const float wgt = expf(__ldcg(pp) - m);
      l = fmaf(__ldcg(pp + 1), wgt, l);
#pragma unroll
      for (int d = 0; d < 16; ++d) o[d] = fmaf(__ldcg(pp + 2 + d), wgt, o[d]);
    }
    l = warp_sum(l);
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d] = warp_sum(o[d]);
    const int t = pair / MD_H, h = pair - t * MD_H;
    if (lane < 16) {
      float v = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d)
        if (d == lane) v = o[d];
      att[t * MD_CI + h * 16 + lane] = v / l;
    }
  }
}

// X[r][:] = LayerNorm(X[r][:] + D[r][:]) for the 32 rows (warp per row), X shared [32][256], D shared [32][ldd]
__device__ __forceinline__ void md_rows_add_ln(float* X, const float* D, int ldd, const float* w, const float* b, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < MD_RB; r += MD_THREADS / 32) {
    float v[8], s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      v[i] = X[r * MD_C + c] + D[r * ldd + c];
      s += v[i];
    }
    const float mean = warp_sum(s) * (1.0f / MD_C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) q += (v[i] - mean) * (v[i] - mean);
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / MD_C) + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      X[r * MD_C + c] = (v[i] - mean) * rstd * __ldg(w + c) + __ldg(b + c);
    }
  }
}

__global__ void __launch_bounds__(MD_THREADS, 1) mask_decoder_kernel(const MdParams p) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float md_smem[];
  float* Xs = md_smem;                       // [32][256]  image rows ("keys" stream)
  float* Ks = Xs + MD_RB * MD_C;             // [32][132]  K projection / image-side q / up1 pixels (aliased per stage)
  float* Vs = Ks + MD_RB * MD_LDK;           // [32][132]  V projection / image-side attention output
  float* TB = Vs + MD_RB * MD_LDK;           // token buffer [12][768] / out-projection tile [32][256] / conv tiles
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int blk = blockIdx.x, b = blockIdx.y, G = p.G, NT = p.Nt, fs = p.fs;
  const MdOff off(G);
  float* ws = p.ws + b * p.ws_obj_stride;
  unsigned* bar = p.bar + b;
  unsigned nbar = 0;
  const VgMaskDecoderWeights& W = p.w;
  const float* tok_pe = p.tokens + static_cast<long long>(b) * NT * MD_C;   // query_pe = the initial tokens (transformer.py:88-99)
  const int row0 = blk * MD_RB;                                              // first image row of this block
  const float* pe_rows = p.pe;                                               // unused directly: the pe projections are folded (W.pek / W.peq)
  (void)pe_rows;
#define MD_BAR() md_barrier(bar, (++nbar) * static_cast<unsigned>(G))

  // keys = image embedding + dense prompt embedding (no_mask_embed); mask_decoder.py:199-205
  {
    const float* px = p.pix + b * p.pix_obj_stride + static_cast<long long>(row0) * MD_C;
    for (int i = tid; i < MD_RB * MD_C; i += MD_THREADS) Xs[i] = px[i] + __ldg(p.no_mask + (i & (MD_C - 1)));
  }
  __syncthreads();

  for (int layer = 0; layer < 2; ++layer) {
    const VgMdLayer& L = W.layer[layer];
    // ---- T1: token self-attention q/k/v projections (q, k inputs carry the positional tokens except in layer 0; transformer.py:160-168)
    float* Ta = TB;                    // [NT][256] un-augmented
    float* Tb = TB + MD_NT * MD_C;     // [NT][256] + query_pe
    if (layer == 0) {
      for (int i = tid; i < NT * MD_C; i += MD_THREADS) Ta[i] = tok_pe[i];
    } else {
      md_load_rows(Ta, ws + off.Q, MD_C, MD_C, NT);
    }
    __syncthreads();
    for (int i = tid; i < NT * MD_C; i += MD_THREADS) Tb[i] = layer == 0 ? Ta[i] : Ta[i] + tok_pe[i];
    __syncthreads();
    md_tok_linear(Tb, MD_C, MD_C, L.sa.q.w, MD_C, L.sa.q.b, MD_C, ws + off.QKV, 3 * MD_C, NT, false, nullptr, 0, false, blk, G);
    md_tok_linear(Tb, MD_C, MD_C, L.sa.k.w, MD_C, L.sa.k.b, MD_C, ws + off.QKV + MD_C, 3 * MD_C, NT, false, nullptr, 0, false, blk, G);
    md_tok_linear(Ta, MD_C, MD_C, L.sa.v.w, MD_C, L.sa.v.b, MD_C, ws + off.QKV + 2 * MD_C, 3 * MD_C, NT, false, nullptr, 0, false, blk, G);
    MD_BAR();
    // ---- T2: self-attention core (8 heads x 32, every block redundantly) + distributed out projection (+ residual except in layer 0)
    md_load_rows(TB, ws + off.QKV, 3 * MD_C, 3 * MD_C, NT);
    __syncthreads();
    float* Asa = Ks;                   // [NT][256] attention output (Ks is free until the K projection below)
    for (int job = tid; job < NT * MD_H; job += MD_THREADS) {
      const int t = job / MD_H, h = job - t * MD_H;
      const float* q = TB + t * 3 * MD_C + h * 32;
      float s[MD_NT], m = -INFINITY;
#pragma unroll
      for (int j = 0; j < MD_NT; ++j) {
        s[j] = -INFINITY;
        if (j < NT) {
          const float* k = TB + j * 3 * MD_C + MD_C + h * 32;
          float a = 0.f;
          for (int d = 0; d < 32; ++d) a = fmaf(q[d], k[d], a);
          s[j] = a * 0.17677669529663687f;  // 32^-0.5
          m = fmaxf(m, s[j]);
        }
      }
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < MD_NT; ++j) {
        s[j] = j < NT ? expf(s[j] - m) : 0.f;
        l += s[j];
      }
      const float inv = 1.0f / l;
      for (int d = 0; d < 32; ++d) {
        float o = 0.f;
#pragma unroll
        for (int j = 0; j < MD_NT; ++j)
          if (j < NT) o = fmaf(s[j], TB[j * 3 * MD_C + 2 * MD_C + h * 32 + d], o);
        Asa[t * MD_C + h * 32 + d] = o * inv;
      }
    }
    __syncthreads();
    md_tok_linear(Asa, MD_C, MD_C, L.sa.out.w, MD_C, L.sa.out.b, MD_C, ws + off.QPRE, MD_C, NT, false, layer == 0 ? nullptr : ws + off.Q, MD_C, false, blk, G);
    MD_BAR();
    // ---- T3: norm1 (redundant per block) -> queries; token->image q projection of (queries + query_pe)
    md_load_rows_ln(Ta, ws + off.QPRE, NT, L.n1.w, L.n1.b, 1e-5f);
    __syncthreads();
    for (int i = tid; i < NT * MD_C; i += MD_THREADS) {
      if (blk == 0) ws[off.Q + i] = Ta[i];
      Tb[i] = Ta[i] + tok_pe[i];
    }
    __syncthreads();
    md_tok_linear(Tb, MD_C, MD_C, L.t2i.q.w, MD_C, L.t2i.q.b, MD_CI, ws + off.TQ, MD_CI, NT, false, nullptr, 0, false, blk, G);
    // image side, row-local: K = (keys + pe) Wk + b (pe Wk folded: W.pek), V = keys Wv + b
    md_img_gemm<MD_CI>(Xs, MD_C, MD_C, L.t2i.k.w, L.t2i.k.b, W.pek[layer] + static_cast<long long>(row0) * MD_CI, MD_CI, Ks, MD_LDK);
    md_img_gemm<MD_CI>(Xs, MD_C, MD_C, L.t2i.v.w, L.t2i.v.b, nullptr, 0, Vs, MD_LDK);
    MD_BAR();
    // ---- I1: partial token->image attention over this block's rows
    md_load_rows(TB, ws + off.TQ, MD_CI, MD_CI, NT);
    __syncthreads();
    md_t2i_partials(TB, Ks, Vs, ws + off.PART + static_cast<long long>(blk) * MD_NT * MD_H * 18, NT);
    MD_BAR();
    // ---- T4a: merge the partials (pairs distributed over the blocks)
    md_t2i_merge(ws + off.PART, ws + off.ATT, NT, blk, G);
    MD_BAR();
    // ---- T4b: out projection + residual
    md_load_rows(TB, ws + off.ATT, MD_CI, MD_CI, NT);
    __syncthreads();
    md_tok_linear(TB, MD_CI, MD_CI, L.t2i.out.w, MD_CI, L.t2i.out.b, MD_C, ws + off.QPRE, MD_C, NT, false, ws + off.Q, MD_C, false, blk, G);
    MD_BAR();
    // ---- T5: norm2 -> queries; MLP fc1 + ReLU
    md_load_rows_ln(Ta, ws + off.QPRE, NT, L.n2.w, L.n2.b, 1e-5f);
    __syncthreads();
    if (blk == 0)
      for (int i = tid; i < NT * MD_C; i += MD_THREADS) ws[off.Q + i] = Ta[i];
    md_tok_linear(Ta, MD_C, MD_C, L.fc1.w, MD_C, L.fc1.b, MD_MLP, ws + off.HID, MD_MLP, NT, true, nullptr, 0, false, blk, G);
    MD_BAR();
    // ---- T6: MLP fc2 + residual (K = 2048 in chunks of 512 through the token buffer)
    for (int kc = 0; kc < MD_MLP; kc += 512) {
      __syncthreads();
      md_load_rows(TB, ws + off.HID + kc, MD_MLP, 512, NT);
      __syncthreads();
      md_tok_linear(TB, 512, 512, L.fc2.w + kc, MD_MLP, L.fc2.b, MD_C, ws + off.QPRE, MD_C, NT, false, ws + off.Q, MD_C, kc != 0, blk, G);
    }
    MD_BAR();
    // ---- T7: norm3 -> queries; image->token k = (queries + pe) Wk, v = queries Wv
    md_load_rows_ln(Ta, ws + off.QPRE, NT, L.n3.w, L.n3.b, 1e-5f);
    __syncthreads();
    for (int i = tid; i < NT * MD_C; i += MD_THREADS) {
      if (blk == 0) ws[off.Q + i] = Ta[i];
      Tb[i] = Ta[i] + tok_pe[i];
    }
    __syncthreads();
    md_tok_linear(Tb, MD_C, MD_C, L.i2t.k.w, MD_C, L.i2t.k.b, MD_CI, ws + off.TK, MD_CI, NT, false, nullptr, 0, false, blk, G);
    md_tok_linear(Ta, MD_C, MD_C, L.i2t.v.w, MD_C, L.i2t.v.b, MD_CI, ws + off.TV, MD_CI, NT, false, nullptr, 0, false, blk, G);
    // image side, row-local: q = (keys + pe) Wq + b
    md_img_gemm<MD_CI>(Xs, MD_C, MD_C, L.i2t.q.w, L.i2t.q.b, W.peq[layer] + static_cast<long long>(row0) * MD_CI, MD_CI, Ks, MD_LDK);
    MD_BAR();
    // ---- I2: image->token attention (softmax over the NT tokens per (row, head)), out projection, residual, norm4 — all row-local
    md_load_rows(TB, ws + off.TK, MD_CI, MD_CI, NT);
    md_load_rows(TB + MD_NT * MD_CI, ws + off.TV, MD_CI, MD_CI, NT);
    __syncthreads();
    {
      const int r = tid >> 3, h = tid & 7;
      const float* q = Ks + r * MD_LDK + h * 16;
      float qv[16];
#pragma unroll
      for (int d = 0; d < 16; d += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(q + d);
        qv[d] = t4.x; qv[d + 1] = t4.y; qv[d + 2] = t4.z; qv[d + 3] = t4.w;
      }
      float s[MD_NT], m = -INFINITY;
#pragma unroll
      for (int j = 0; j < MD_NT; ++j) {
        s[j] = -INFINITY;
        if (j < NT) {
          const float* k = TB + j * MD_CI + h * 16;
          float a = 0.f;
#pragma unroll
          for (int d = 0; d < 16; ++d) a = fmaf(qv[d], k[d], a);
          s[j] = a * 0.25f;
          m = fmaxf(m, s[j]);
        }
      }
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < MD_NT; ++j) {
        s[j] = j < NT ? expf(s[j] - m) : 0.f;
        l += s[j];
      }
      const float inv = 1.0f / l;
      float o[16];
#pragma unroll
      for (int d = 0; d < 16; ++d) o[d] = 0.f;
#pragma unroll
      for (int j = 0; j < MD_NT; ++j)
        if (j < NT) {
          const float* v = TB + MD_NT * MD_CI + j * MD_CI + h * 16;
#pragma unroll
          for (int d = 0; d < 16; ++d) o[d] = fmaf(s[j], v[d], o[d]);
        }
#pragma unroll
      for (int d = 0; d < 16; d += 4)
        *reinterpret_cast<float4*>(Vs + r * MD_LDK + h * 16 + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
    }
    __syncthreads();
    md_img_gemm<MD_C>(Vs, MD_LDK, MD_CI, L.i2t.out.w, L.i2t.out.b, nullptr, 0, TB, MD_C);   // [32][256] into the token buffer
    __syncthreads();
    md_rows_add_ln(Xs, TB, MD_C, L.n4.w, L.n4.b, 1e-5f);
    __syncthreads();
  }

  // ---- final token->image attention (transformer.py:99-115)
  {
    float* Ta = TB;
    md_load_rows(Ta, ws + off.Q, MD_C, MD_C, NT);
    __syncthreads();
    for (int i = tid; i < NT * MD_C; i += MD_THREADS) Ta[i] += tok_pe[i];
    __syncthreads();
    md_tok_linear(Ta, MD_C, MD_C, W.final_t2i.q.w, MD_C, W.final_t2i.q.b, MD_CI, ws + off.TQ, MD_CI, NT, false, nullptr, 0, false, blk, G);
    md_img_gemm<MD_CI>(Xs, MD_C, MD_C, W.final_t2i.k.w, W.final_t2i.k.b, W.pek[2] + static_cast<long long>(row0) * MD_CI, MD_CI, Ks, MD_LDK);
    md_img_gemm<MD_CI>(Xs, MD_C, MD_C, W.final_t2i.v.w, W.final_t2i.v.b, nullptr, 0, Vs, MD_LDK);
    MD_BAR();
    md_load_rows(TB, ws + off.TQ, MD_CI, MD_CI, NT);
    __syncthreads();
    md_t2i_partials(TB, Ks, Vs, ws + off.PART + static_cast<long long>(blk) * MD_NT * MD_H * 18, NT);
    MD_BAR();
    md_t2i_merge(ws + off.PART, ws + off.ATT, NT, blk, G);
    MD_BAR();
    md_load_rows(TB, ws + off.ATT, MD_CI, MD_CI, NT);
    __syncthreads();
    md_tok_linear(TB, MD_CI, MD_CI, W.final_t2i.out.w, MD_CI, W.final_t2i.out.b, MD_C, ws + off.QPRE, MD_C, NT, false, ws + off.Q, MD_C, false, blk, G);
    MD_BAR();
  }
  // ---- hs = norm_final(queries); hypernetwork / IoU / object-score MLPs: 6 heads x 256 columns per hidden layer (mask_decoder.py:227-240)
  float* HSs = TB;  // [NT][256]
  md_load_rows_ln(HSs, ws + off.QPRE, NT, W.norm_final.w, W.norm_final.b, 1e-5f);
  __syncthreads();
  if (blk == 0) {  // mask tokens out (hs[:, 2:6]) — sam_tokens_out / object pointers
    for (int i = tid; i < 4 * MD_C; i += MD_THREADS) p.mask_tokens[static_cast<long long>(b) * 4 * MD_C + i] = HSs[2 * MD_C + i];
  }
  {
    // head j: 0..3 hypernetwork MLP of mask token j (input row 2 + j), 4 IoU head (row 1), 5 object-score head (row 0)
    float* H1s = TB + MD_NT * MD_C;  // staging for the hidden activations [6][256]
    for (int stage = 0; stage < 2; ++stage) {
      const float* in = stage == 0 ? HSs : H1s;
      float* outg = ws + (stage == 0 ? off.H1 : off.H2);
      for (int o = blk + G * warp; o < 6 * MD_C; o += G * (MD_THREADS / 32)) {
        const int head = o >> 8, n = o & 255;
        const VgLinear lin = head < 4 ? W.hyper[head][stage] : (head == 4 ? W.iou[stage] : W.obj[stage]);
        const float* x = stage == 0 ? in + (head < 4 ? 2 + head : (head == 4 ? 1 : 0)) * MD_C : in + head * MD_C;
        float a = 0.f;
        for (int k = lane; k < MD_C; k += 32) a = fmaf(x[k], __ldg(lin.w + n * MD_C + k), a);
        a = warp_sum(a);
        if (lane == 0) outg[head * MD_C + n] = fmaxf(a + lin.b[n], 0.f);
      }
      MD_BAR();
      md_load_rows(H1s, outg, MD_C, MD_C, 6);
      __syncthreads();
    }
    // last layers: hyper -> [4][32]; IoU -> 4 (sigmoid); object score -> 1. 133 outputs distributed over the blocks.
    for (int o = blk + G * warp; o < 4 * 32 + 4 + 1; o += G * (MD_THREADS / 32)) {
      int head, n;
      if (o < 128) { head = o >> 5; n = o & 31; } else if (o < 132) { head = 4; n = o - 128; } else { head = 5; n = 0; }
      const VgLinear lin = head < 4 ? W.hyper[head][2] : (head == 4 ? W.iou[2] : W.obj[2]);
      const float* x = H1s + head * MD_C;
      float a = 0.f;
      for (int k = lane; k < MD_C; k += 32) a = fmaf(x[k], __ldg(lin.w + n * MD_C + k), a);
      a = warp_sum(a) + lin.b[n];
      if (lane == 0) {
        if (head < 4) ws[off.HYP + head * 32 + n] = a;
        else if (head == 4) p.iou[b * 4 + n] = 1.0f / (1.0f + expf(-a));
        else p.obj[b] = a;
      }
    }
    MD_BAR();
  }
  // ---- upscaling + masks, row-local (mask_decoder.py:221-234). up1 = GELU(LN2d(ConvT(keys) + feat_s1)); up2 = GELU(ConvT(up1) + feat_s0);
  // masks[m] = hyper_in[m] . up2
  __shared__ float hyp[4 * 32];
  if (tid < 128) hyp[tid] = __ldcg(ws + off.HYP + tid);
  md_img_gemm<MD_C>(Xs, MD_C, MD_C, W.up1.w, nullptr, nullptr, 0, TB, MD_C);   // [32][(dy,dx,co) = 4*64]
  __syncthreads();
  float* U1 = Ks;  // [128 up1 pixels][64] = 8192 floats over the Ks + Vs regions (32*132*2 = 8448)
  for (int px = warp; px < 4 * MD_RB; px += MD_THREADS / 32) {
    const int r = px >> 2, sub = px & 3;
    const int grow = row0 + r, y = grow / fs, x = grow - y * fs;
    const long long pix1 = static_cast<long long>(2 * y + (sub >> 1)) * (2 * fs) + 2 * x + (sub & 1);
    float v[2], s = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + 32 * i;
      v[i] = TB[r * MD_C + sub * 64 + c] + __ldg(W.up1.b + c) + __ldg(p.s1 + pix1 * 64 + c);
      s += v[i];
    }
    const float mean = warp_sum(s) * (1.0f / 64);
    float q = (v[0] - mean) * (v[0] - mean) + (v[1] - mean) * (v[1] - mean);
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / 64) + 1e-6f);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + 32 * i;
      U1[px * 64 + c] = md_gelu((v[i] - mean) * rstd * __ldg(W.up_ln.w + c) + __ldg(W.up_ln.b + c));
    }
  }
  __syncthreads();
  const int side = 4 * fs;
  float* mout = p.masks + static_cast<long long>(b) * 4 * side * side;
  for (int pass = 0; pass < 4; ++pass) {  // 32 up1 pixels per pass -> [32][(dy,dx,co) = 4*32] -> 128 up2 pixels
    md_img_gemm<MD_CI>(U1 + pass * 32 * 64, 64, 64, W.up2.w, nullptr, nullptr, 0, TB, MD_LDK);
    __syncthreads();
    if (tid < 128) {
      const int pl = tid >> 2, sub2 = tid & 3;           // local up1 pixel, position inside its 2x2 up2 patch
      const int px = pass * 32 + pl, r = px >> 2, sub = px & 3;
      const int grow = row0 + r, y = grow / fs, x = grow - y * fs;
      const int Y2 = 2 * (2 * y + (sub >> 1)) + (sub2 >> 1), X2 = 2 * (2 * x + (sub & 1)) + (sub2 & 1);
      const long long pix2 = static_cast<long long>(Y2) * side + X2;
      float m[4] = {0.f, 0.f, 0.f, 0.f};
      const float* g2 = TB + pl * MD_LDK + sub2 * 32;
      const float* f0 = p.s0 + pix2 * 32;
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        const float4 gv = *reinterpret_cast<const float4*>(g2 + c);
        const float4 fv = __ldg(reinterpret_cast<const float4*>(f0 + c));
        const float4 bv = __ldg(reinterpret_cast<const float4*>(W.up2.b + c));
        const float u[4] = {md_gelu(gv.x + fv.x + bv.x), md_gelu(gv.y + fv.y + bv.y), md_gelu(gv.z + fv.z + bv.z), md_gelu(gv.w + fv.w + bv.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int mm = 0; mm < 4; ++mm) m[mm] = fmaf(hyp[mm * 32 + c + j], u[j], m[mm]);
      }
#pragma unroll
      for (int mm = 0; mm < 4; ++mm) mout[static_cast<long long>(mm) * side * side + pix2] = m[mm];
    }
    __syncthreads();
  }
#undef MD_BAR
}

}  // namespace vg

using namespace vg;

extern "C" long long vg_mask_decoder_workspace_bytes(int B, int fs) {
  const int G = fs * fs / MD_RB;
  const MdOff off(G > 0 ? G : 1);
  return static_cast<long long>(B) * off.total * 4 + 256 * 4;
}

extern "C" int vg_mask_decoder(const VgMaskDecoder* d, void* stream) {
  if (!d || !d->pix || !d->dense_pe || !d->no_mask || !d->s0 || !d->s1 || !d->tokens || !d->weights || !d->masks || !d->iou || !d->mask_tokens || !d->obj ||
      !d->workspace)
    return set_error(VG_EINVAL, "vg_mask_decoder: null pointer");
  if (d->C != MD_C || d->heads != MD_H || d->mlp != MD_MLP || d->depth != 2)
    return set_error(VG_ENOSUP, "vg_mask_decoder: built for the SAM2 decoder geometry (C 256, 8 heads, MLP 2048, depth 2)");
  if (d->B <= 0 || d->B > 256 || d->Nt < 1 || d->Nt > MD_NT || d->fs <= 0 || (d->fs * d->fs) % MD_RB)
    return set_error(VG_EINVAL, "vg_mask_decoder: need 1 <= Nt <= %d, feat_size^2 %% %d == 0, B <= 256", MD_NT, MD_RB);
  const int G = d->fs * d->fs / MD_RB;
  if (d->workspace_bytes < vg_mask_decoder_workspace_bytes(d->B, d->fs)) return set_error(VG_EINVAL, "vg_mask_decoder: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const MdOff off(G);
  MdParams p;
  p.pix = d->pix; p.pix_obj_stride = d->pix_shared ? 0 : static_cast<long long>(d->fs) * d->fs * MD_C;
  p.pe = d->dense_pe; p.no_mask = d->no_mask; p.s0 = d->s0; p.s1 = d->s1; p.tokens = d->tokens;
  p.B = d->B; p.Nt = d->Nt; p.fs = d->fs; p.G = G;
  p.masks = d->masks; p.iou = d->iou; p.mask_tokens = d->mask_tokens; p.obj = d->obj;
  p.bar = static_cast<unsigned*>(d->workspace);          // first 256 words: one barrier counter per object
  p.ws = reinterpret_cast<float*>(static_cast<char*>(d->workspace) + 256 * 4);
  p.ws_obj_stride = off.total;
  p.w = *d->weights;
  cudaError_t e = cudaMemsetAsync(p.bar, 0, 256 * 4, s);
  if (e != cudaSuccess) return set_error(VG_ECUDA, "vg_mask_decoder memset: %s", cudaGetErrorString(e));
  const int smem = (MD_RB * MD_C + 2 * MD_RB * MD_LDK + MD_TB) * 4;
  static bool attr = false;
  if (!attr) {
    e = cudaFuncSetAttribute(mask_decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(VG_ECUDA, "vg_mask_decoder smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  // all G blocks of an object must be co-resident (grid barrier); objects are independent groups dispatched in order
  static int max_resident = 0;
  if (max_resident == 0) {
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mask_decoder_kernel, MD_THREADS, smem);
    max_resident = per_sm * sms;
  }
  if (G > max_resident) return set_error(VG_ENOSUP, "vg_mask_decoder: %d row blocks per object exceed the %d co-resident blocks of this device", G, max_resident);
  launch_k(mask_decoder_kernel, dim3(G, d->B), dim3(MD_THREADS), smem, s, p);
  count_launch();
  return check_launch("mask_decoder_kernel");
}
