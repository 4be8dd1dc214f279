// Flash-attention forward (online softmax, no score materialisation) for every attention on the path:
// CLIP (hd 64), Hiera windowed/global (hd 72), InternVideo2 (hd 88), Phi-3 causal / sliding-window (hd 96),
// SAM2 mask decoder (hd 16/32) and memory attention (1 head x 256).
// Q/K/V/O are addressed through (batch, head, row) element strides so packed qkv GEMM outputs are read in place;
// Hiera windows are addressed by index arithmetic on the un-partitioned [H,W] token grid (no window_partition copy).
// Bring-up tensor path: mma.sync m16n8k16 bf16 with fp32 accumulate, cp.async double-buffered K/V tiles.
#include <cstdlib>

#include "common.h"
#include "pdl.cuh"
#include "ptx.cuh"

namespace vg {

struct AttnParams {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* o;
  long long q_bs, q_hs, q_rs, k_bs, k_hs, k_rs, v_bs, v_hs, v_rs, o_bs, o_hs, o_rs;
  int B, H, Nq, Nk, D;
  float scale;
  int causal, window;
  int q_win, q_gridw, q_nwx, q_nwy;  // 2D window addressing for q/o rows (0 = linear)
  int k_win, k_gridw, k_nwx, k_nwy;  // 2D window addressing for k/v rows
  float* ws;                         // split-K partials for single-query attention
  int nsplit;
  const int* nk_dev;                 // optional: number of keys read from device memory (graph-replayable decode)
  int nk_per_batch;                  // nk_dev holds one key count per batch entry (batched decode: every sequence has its own length)
  int* counters;                     // optional [B*H] zeroed ints: the last split block of a (b,h) merges in-kernel (no merge launch)
  float rope_theta;                  // > 0: decode step with fused RoPE + KV append (q = packed un-rotated [q|k|v] row of position Nk-1)
};

// row index inside the (batch-strided) tensor for token t of batch entry b
__device__ __forceinline__ long long tok_row(int b, int t, int win, int gridw, int nwx, int nwy) {
  if (win == 0) return t;
  const int per_img = nwx * nwy;
  const int img = b / per_img, w = b - img * per_img;
  const int wy = w / nwx, wx = w - wy * nwx;
  const int ty = t / win, tx = t - ty * win;
  return static_cast<long long>(img) * (static_cast<long long>(nwy) * win * gridw) + static_cast<long long>(wy * win + ty) * gridw + wx * win + tx;
}

template <int DP>
__global__ void __launch_bounds__(128) flash_attn_kernel(const AttnParams p) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  constexpr int BM = 64, BN = 64, LD = DP + 8, CH = DP / 8;  // CH: 16-byte chunks per row
  constexpr bool Q_IN_REGS = DP <= 128;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sK = sQ + BM * LD;
  __nv_bfloat16* sV = sK + 2 * BN * LD;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q_tiles = (p.Nq + BM - 1) / BM;
  const int split = blockIdx.x / q_tiles;  // split-K over keys (nsplit > 1 only for non-causal, low-CTA-count problems)
  const int q0 = (blockIdx.x - split * q_tiles) * BM, h = blockIdx.y, b = blockIdx.z;
  const int off = p.Nk - p.Nq;  // causal alignment: query i may see keys j <= i + off
  int k_end = p.Nk;
  if (p.causal) k_end = min(p.Nk, q0 + BM + off);
  int k_begin = 0;
  if (p.window > 0) k_begin = max(0, (q0 + off - p.window) / BN * BN);
  if (p.nsplit > 1) {
    const int per = ((p.Nk + p.nsplit - 1) / p.nsplit + BN - 1) / BN * BN;
    k_begin = split * per;
    k_end = min(p.Nk, k_begin + per);
  }
  const int ntiles = (k_end - k_begin + BN - 1) / BN;

  const bool lin_q = p.q_win == 0, lin_k = p.k_win == 0;
  const __nv_bfloat16* qb = p.q + (lin_q ? b * p.q_bs : 0) + h * p.q_hs;
  const __nv_bfloat16* kb = p.k + (lin_k ? b * p.k_bs : 0) + h * p.k_hs;
  const __nv_bfloat16* vb = p.v + (lin_k ? b * p.v_bs : 0) + h * p.v_hs;

  for (int i = tid; i < BM * CH; i += 128) {
    const int r = i / CH, c = i - r * CH;
    const bool ok = (q0 + r < p.Nq) && (c * 8 < p.D);
    const long long row = ok ? tok_row(b, q0 + r, p.q_win, p.q_gridw, p.q_nwx, p.q_nwy) : 0;
    cp_async16(sQ + r * LD + c * 8, qb + row * p.q_rs + c * 8, ok);
  }
  auto load_kv = [&](int tile, int stage) {
    const int kbase = k_begin + tile * BN;
    for (int i = tid; i < BN * CH; i += 128) {
      const int r = i / CH, c = i - r * CH;
      const bool ok = (kbase + r < p.Nk) && (c * 8 < p.D);
      const long long row = ok ? tok_row(b, kbase + r, p.k_win, p.k_gridw, p.k_nwx, p.k_nwy) : 0;
      cp_async16(sK + (stage * BN + r) * LD + c * 8, kb + row * p.k_rs + c * 8, ok);
      cp_async16(sV + (stage * BN + r) * LD + c * 8, vb + row * p.v_rs + c * 8, ok);
    }
  };
  if (ntiles > 0) load_kv(0, 0);
  cp_async_commit();

  const float sl2 = p.scale * 1.4426950408889634f;
  float o_acc[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[Q_IN_REGS ? DP / 16 : 1][4];

  const int g = lane >> 2, t4 = lane & 3;
  const int a_row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;       // A-operand ldmatrix address
  const int b_row = (lane & 7) + (lane >> 4) * 8, b_col = ((lane >> 3) & 1) * 8;                  // B (K tile, non-trans)
  const int v_row = (lane & 7) + ((lane >> 3) & 1) * 8, v_col = (lane >> 4) * 8;                  // B (V tile, trans)

  for (int it = 0; it < ntiles; ++it) {
    if (it + 1 < ntiles) load_kv(it + 1, (it + 1) & 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (it == 0 && Q_IN_REGS) {
#pragma unroll
      for (int kt = 0; kt < DP / 16; ++kt) ldmatrix_x4(qf[Q_IN_REGS ? kt : 0], sQ + a_row * LD + kt * 16 + a_col);
    }
    const __nv_bfloat16* tK = sK + (it & 1) * BN * LD;
    const __nv_bfloat16* tV = sV + (it & 1) * BN * LD;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < DP / 16; ++kt) {
      uint32_t a[4];
      if (Q_IN_REGS) {
        a[0] = qf[Q_IN_REGS ? kt : 0][0]; a[1] = qf[Q_IN_REGS ? kt : 0][1]; a[2] = qf[Q_IN_REGS ? kt : 0][2]; a[3] = qf[Q_IN_REGS ? kt : 0][3];
      } else {
        ldmatrix_x4(a, sQ + a_row * LD + kt * 16 + a_col);
      }
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        uint32_t bb[4];
        ldmatrix_x4(bb, tK + (n2 * 16 + b_row) * LD + kt * 16 + b_col);
        mma_bf16_16816(s[2 * n2], a, bb[0], bb[1]);
        mma_bf16_16816(s[2 * n2 + 1], a, bb[2], bb[3]);
      }
    }
    // masking + online softmax (rows g and g+8 of this warp's 16)
    const int kbase = k_begin + it * BN;
    const int qrow0 = q0 + warp * 16 + g;
    float mx[2] = {-INFINITY, -INFINITY};
    // interior tiles (all 64 keys valid for all 64 rows) skip the mask arithmetic; block-uniform condition
    const bool need_mask = (kbase + BN > p.Nk) || (p.causal && (kbase + BN - 1 > q0 + off)) || (p.window > 0 && (kbase < q0 + BM - 1 + off - p.window));
    if (need_mask) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = kbase + nt * 8 + 2 * t4 + (e & 1);
          const int row = qrow0 + (e >> 1) * 8;
          bool ok = col < p.Nk;
          if (p.causal) ok = ok && (col <= row + off);
          if (p.window > 0) ok = ok && (col >= row + off - p.window);
          s[nt][e] = ok ? s[nt][e] : -INFINITY;
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
    }
    float alpha[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      msc[r] = (m_new == -INFINITY) ? 0.f : m_new * sl2;
      alpha[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f(m_run[r] * sl2 - msc[r]);
      m_run[r] = m_new;
      l_run[r] *= alpha[r];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f(s[nt][e] * sl2 - msc[e >> 1]);  // exp2(-inf) = 0 for masked entries
        s[nt][e] = pv;
        l_run[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      o_acc[i][0] *= alpha[0]; o_acc[i][1] *= alpha[0];
      o_acc[i][2] *= alpha[1]; o_acc[i][3] *= alpha[1];
    }
    // O += P V
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * k2][0], s[2 * k2][1]);
      a[1] = pack_bf16(s[2 * k2][2], s[2 * k2][3]);
      a[2] = pack_bf16(s[2 * k2 + 1][0], s[2 * k2 + 1][1]);
      a[3] = pack_bf16(s[2 * k2 + 1][2], s[2 * k2 + 1][3]);
#pragma unroll
      for (int d2 = 0; d2 < DP / 16; ++d2) {
        uint32_t bb[4];
        ldmatrix_x4_trans(bb, tV + (k2 * 16 + v_row) * LD + d2 * 16 + v_col);
        mma_bf16_16816(o_acc[2 * d2], a, bb[0], bb[1]);
        mma_bf16_16816(o_acc[2 * d2 + 1], a, bb[2], bb[3]);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  if (p.nsplit > 1) {  // un-normalised partial rows: [split][b][h][row][DP + 2] = {m, l, o[DP]}
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = q0 + warp * 16 + g + r * 8;
      if (row >= p.Nq) continue;
      float* wp = p.ws + ((((static_cast<long long>(split) * p.B + b) * p.H + h) * p.Nq) + row) * (DP + 2);
      if (t4 == 0) {
        wp[0] = m_run[r] * p.scale;  // scaled logit units
        wp[1] = l_run[r];
      }
#pragma unroll
      for (int nt = 0; nt < DP / 8; ++nt) {
        wp[2 + nt * 8 + 2 * t4] = o_acc[nt][2 * r];
        wp[2 + nt * 8 + 2 * t4 + 1] = o_acc[nt][2 * r + 1];
      }
    }
    return;
  }
  __nv_bfloat16* ob = p.o + (lin_q ? b * p.o_bs : 0) + h * p.o_hs;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + r * 8;
    if (row >= p.Nq) continue;
    const float inv = l_run[r] > 0.f ? 1.0f / l_run[r] : 0.f;
    __nv_bfloat16* orow = ob + tok_row(b, row, p.q_win, p.q_gridw, p.q_nwx, p.q_nwy) * p.o_rs;
#pragma unroll
    for (int nt = 0; nt < DP / 8; ++nt) {
      const int col = nt * 8 + 2 * t4;
      if (col < p.D) *reinterpret_cast<uint32_t*>(orow + col) = pack_bf16(o_acc[nt][2 * r] * inv, o_acc[nt][2 * r + 1] * inv);
    }
  }
}


// ---------------------------------------------------------------------------------------------- small Hiera windows (16 / 64 tokens)
// Hiera stages 1-2 attend inside 8x8 / 4x4-token windows with 72-wide heads. The generic 64x64-tile kernel spends a block per (window, head)
// on a tile that is 75-94 % padding (ncu: 385 us per stage-2 launch for 300 MB of traffic). Here ONE block owns one window and ALL its heads:
// the packed qkv rows of the window's tokens (contiguous 3*C bf16 each) are copied to shared memory once with 16-byte cp.async, warp
// (head, 16-row group) computes S = Q K^T, the softmax over the whole window and O = P V entirely in registers with mma.sync m16n8k16
// (+ an m16n8k8 step for the 8-feature tail of D = 72), writes O back over its own Q slice, and the block stores whole output rows.
template <int T>
__global__ void __launch_bounds__(512) window_attn_small_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld, __nv_bfloat16* __restrict__ out,
                                                                long long ldo, int H, int win, int gridw, int nwx, float scale) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  constexpr int D = 72, RG = T / 16;
  extern __shared__ __align__(16) uint8_t smem_win[];
  __nv_bfloat16* s = reinterpret_cast<__nv_bfloat16*>(smem_win);
  const int C = H * D, LDS = 3 * C + 8;  // padded row: consecutive rows start 16 B apart modulo 128 B -> conflict-free ldmatrix
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const long long wy0 = static_cast<long long>(b / nwx) * win;
  const int wx0 = (b % nwx) * win;
  auto grow = [&](int t) { return (wy0 + t / win) * gridw + wx0 + t % win; };
  const int cpr = (3 * C) >> 3;  // 16-byte chunks per token row
  for (int i = tid; i < T * cpr; i += blockDim.x) {
    const int t = i / cpr, c = i - t * cpr;
    cp_async16(s + t * LDS + c * 8, qkv + grow(t) * ld + c * 8, true);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int hh = warp / RG, rg = warp - hh * RG;
  if (hh < H) {
    const __nv_bfloat16* sQ = s + hh * D;
    const __nv_bfloat16* sK = s + C + hh * D;
    const __nv_bfloat16* sV = s + 2 * C + hh * D;
    const int g = lane >> 2, t4 = lane & 3;
    const int a_row = rg * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;  // A operand (x4): rows 0-7 / 8-15, cols 0-7 / 8-15
    const int b_row = (lane & 7) + (lane >> 4) * 8, b_col = ((lane >> 3) & 1) * 8;            // K tile (x4, non-trans)
    const int v_row = (lane & 7) + ((lane >> 3) & 1) * 8, v_col = (lane >> 4) * 8;            // V tile (x4, trans)
    const int x2_row = (lane & 7) + ((lane >> 3) & 1) * 8;                                    // x2 loads: matrices = rows 0-7, 8-15
    float sc[T / 8][4];
#pragma unroll
    for (int i = 0; i < T / 8; ++i) sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      uint32_t a[4];
      ldmatrix_x4(a, sQ + a_row * LDS + kt * 16 + a_col);
#pragma unroll
      for (int n2 = 0; n2 < T / 16; ++n2) {
        uint32_t bb[4];
        ldmatrix_x4(bb, sK + (n2 * 16 + b_row) * LDS + kt * 16 + b_col);
        mma_bf16_16816(sc[2 * n2], a, bb[0], bb[1]);
        mma_bf16_16816(sc[2 * n2 + 1], a, bb[2], bb[3]);
      }
    }
    {  // features 64..71
      uint32_t a2[2];
      ldmatrix_x2(a2, sQ + (rg * 16 + x2_row) * LDS + 64);
#pragma unroll
      for (int n2 = 0; n2 < T / 16; ++n2) {
        uint32_t b2[2];
        ldmatrix_x2(b2, sK + (n2 * 16 + x2_row) * LDS + 64);
        mma_bf16_1688(sc[2 * n2], a2, b2[0]);
        mma_bf16_1688(sc[2 * n2 + 1], a2, b2[1]);
      }
    }
    // softmax over the whole window (rows g and g + 8 of this warp's 16)
    const float sl2 = scale * 1.4426950408889634f;
    float mx[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < T / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) mx[e >> 1] = fmaxf(mx[e >> 1], sc[nt][e]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mx[r] *= sl2;
    }
#pragma unroll
    for (int nt = 0; nt < T / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        sc[nt][e] = ex2_approx(fmaf(sc[nt][e], sl2, -mx[e >> 1]));
        l[e >> 1] += sc[nt][e];
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
    }
    // O = P V
    float o[9][4];
#pragma unroll
    for (int i = 0; i < 9; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
    for (int k2 = 0; k2 < T / 16; ++k2) {
      uint32_t a[4];
      a[0] = pack_bf16(sc[2 * k2][0], sc[2 * k2][1]);
      a[1] = pack_bf16(sc[2 * k2][2], sc[2 * k2][3]);
      a[2] = pack_bf16(sc[2 * k2 + 1][0], sc[2 * k2 + 1][1]);
      a[3] = pack_bf16(sc[2 * k2 + 1][2], sc[2 * k2 + 1][3]);
#pragma unroll
      for (int d2 = 0; d2 < 4; ++d2) {
        uint32_t bb[4];
        ldmatrix_x4_trans(bb, sV + (k2 * 16 + v_row) * LDS + d2 * 16 + v_col);
        mma_bf16_16816(o[2 * d2], a, bb[0], bb[1]);
        mma_bf16_16816(o[2 * d2 + 1], a, bb[2], bb[3]);
      }
      uint32_t b2[2];
      ldmatrix_x2_trans(b2, sV + (k2 * 16 + x2_row) * LDS + 64);
      mma_bf16_16816(o[8], a, b2[0], b2[1]);
    }
    // normalised O over this warp's own Q slice (nobody else reads these rows x columns)
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float inv = 1.0f / l[r];
      __nv_bfloat16* orow = s + (rg * 16 + g + r * 8) * LDS + hh * D;
#pragma unroll
      for (int nt = 0; nt < 9; ++nt) *reinterpret_cast<uint32_t*>(orow + nt * 8 + 2 * t4) = pack_bf16(o[nt][2 * r] * inv, o[nt][2 * r + 1] * inv);
    }
  }
  __syncthreads();
  const int cpo = C >> 3;
  for (int i = tid; i < T * cpo; i += blockDim.x) {
    const int t = i / cpo, c = i - t * cpo;
    *reinterpret_cast<uint4*>(out + grow(t) * ldo + c * 8) = *reinterpret_cast<const uint4*>(s + t * LDS + c * 8);
  }
}

// Returns VG_OK when the small-window kernel handled the call, VG_ENOSUP otherwise.
static int window_attn_small_try(const VgAttention* a, cudaStream_t s) {
  static const int enabled = [] { const char* e = getenv("VG_WIN_SMALL"); return e ? atoi(e) : 1; }();
  if (!enabled || a->q_win == 0 || a->q_win != a->k_win || a->q_gridw != a->k_gridw || a->q_nwx != a->k_nwx || a->q_nwy != a->k_nwy) return VG_ENOSUP;
  const int T = a->q_win * a->q_win;
  if ((T != 16 && T != 64) || a->Nq != T || a->Nk != T || a->D != 72 || a->causal || a->window) return VG_ENOSUP;
  const int warps = a->H * T / 16;
  if (warps > 16) return VG_ENOSUP;
  const long long C = static_cast<long long>(a->H) * a->D;
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(a->q);
  // q, k, v must be the three column blocks of one packed [rows, 3C] matrix; o rows must hold whole [H, D] outputs
  if (static_cast<const __nv_bfloat16*>(a->k) != q + C || static_cast<const __nv_bfloat16*>(a->v) != q + 2 * C || a->q_hs != a->D || a->k_hs != a->D ||
      a->v_hs != a->D || a->o_hs != a->D || a->q_rs != a->k_rs || a->q_rs != a->v_rs || (a->q_rs % 8) || (a->o_rs % 8) ||
      (reinterpret_cast<uintptr_t>(a->o) & 15))
    return VG_ENOSUP;
  const int smem = T * (3 * static_cast<int>(C) + 8) * 2;
  if (smem > 200 * 1024) return VG_ENOSUP;
  auto kern = T == 16 ? window_attn_small_kernel<16> : window_attn_small_kernel<64>;
  static bool attr[2] = {false, false};
  if (!attr[T == 64]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_error(VG_ECUDA, "window attention smem attr: %s", cudaGetErrorString(e));
    attr[T == 64] = true;
  }
  launch_k(kern, dim3(a->B), dim3(warps * 32), smem, s, q, a->q_rs, static_cast<__nv_bfloat16*>(a->o), a->o_rs, a->H, a->q_win, a->q_gridw, a->q_nwx,
           a->scale);
  count_launch();
  return check_launch("window_attn_small_kernel");
}

// ---------------------------------------------------------------------------------------------- single-query attention (LLM decode)
// HBM/L2-bound: one block per (head, batch); each warp walks 32-key groups: lane-per-key QK^T (no shuffles in the dot product),
// warp-level online softmax, then lane-per-feature PV with the probabilities staged in shared memory; warps merged at the end.
template <int DMAX>
__global__ void __launch_bounds__(128) attn_decode_kernel(AttnParams p) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  if (p.nk_dev != nullptr) p.Nk = p.nk_dev[p.nk_per_batch ? blockIdx.y : 0];
  constexpr int NW = 4, PPL = DMAX / 64;  // warps per block; bf16x2 feature pairs per lane in the PV phase
  __shared__ float s_q[DMAX];
  __shared__ float s_p[NW][32];
  __shared__ float s_o[NW][DMAX];
  __shared__ float s_m[NW], s_l[NW];
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __nv_bfloat16* qp = p.q + b * p.q_bs + h * p.q_hs;
  const __nv_bfloat16* kb = p.k + b * p.k_bs + h * p.k_hs;
  const __nv_bfloat16* vb = p.v + b * p.v_bs + h * p.v_hs;
  for (int d = tid; d < DMAX; d += 128) s_q[d] = d < p.D ? __bfloat162float(qp[d]) * p.scale : 0.f;
  __syncthreads();
  int k_begin = 0;
  if (p.window > 0) k_begin = max(0, p.Nk - 1 - p.window);
  // this block's key range (split-K over blockIdx.z)
  const int span = ((p.Nk - k_begin + p.nsplit - 1) / p.nsplit + 31) / 32 * 32;
  const int k_lo = k_begin + blockIdx.z * span;
  const int k_hi = min(p.Nk, k_lo + span);
  float m_run = -INFINITY, l_run = 0.f, o[2 * PPL];
#pragma unroll
  for (int e = 0; e < 2 * PPL; ++e) o[e] = 0.f;
  const int chunks = p.D >> 3, npairs = p.D >> 1;
  for (int g0 = k_lo + warp * 32; g0 < k_hi; g0 += NW * 32) {
    const int j = g0 + lane;
    float sc = -INFINITY;
    if (j < k_hi) {  // lane-per-key dot product: 16-byte loads of this lane's key row
      const uint4* kr = reinterpret_cast<const uint4*>(kb + static_cast<long long>(j) * p.k_rs);
      float acc = 0.f;
#pragma unroll 4
      for (int c = 0; c < chunks; ++c) {
        const uint4 kv = __ldg(kr + c);
        const float* qq = s_q + c * 8;
        acc = fmaf(bf16_lo(kv.x), qq[0], acc); acc = fmaf(bf16_hi(kv.x), qq[1], acc);
        acc = fmaf(bf16_lo(kv.y), qq[2], acc); acc = fmaf(bf16_hi(kv.y), qq[3], acc);
        acc = fmaf(bf16_lo(kv.z), qq[4], acc); acc = fmaf(bf16_hi(kv.z), qq[5], acc);
        acc = fmaf(bf16_lo(kv.w), qq[6], acc); acc = fmaf(bf16_hi(kv.w), qq[7], acc);
      }
      sc = acc;
    }
    const float m_new = fmaxf(m_run, warp_max(sc));
    const float pj = (j < k_hi) ? __expf(sc - m_new) : 0.f;
    const float alpha = (m_run == -INFINITY) ? 0.f : __expf(m_run - m_new);
    l_run = l_run * alpha + warp_sum(pj);
    m_run = m_new;
    s_p[warp][lane] = pj;
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 2 * PPL; ++e) o[e] *= alpha;
    // lane-per-feature-pair PV: 4 keys per step, all loads issued before the FMAs (rows past k_hi are clamped; their p is 0)
    for (int jj = 0; jj < 32; jj += 4) {
      if (g0 + jj >= k_hi) break;
      uint32_t vv[4][PPL];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int key = min(g0 + jj + u, k_hi - 1);
        const uint32_t* vr = reinterpret_cast<const uint32_t*>(vb + static_cast<long long>(key) * p.v_rs);
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
          const int pr = lane + e * 32;
          vv[u][e] = pr < npairs ? __ldg(vr + pr) : 0u;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float pv = s_p[warp][jj + u];
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
          o[2 * e] = fmaf(pv, bf16_lo(vv[u][e]), o[2 * e]);
          o[2 * e + 1] = fmaf(pv, bf16_hi(vv[u][e]), o[2 * e + 1]);
        }
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    s_m[warp] = m_run;
    s_l[warp] = l_run;
  }
#pragma unroll
  for (int e = 0; e < PPL; ++e) {
    const int pr = lane + e * 32;
    s_o[warp][2 * pr] = o[2 * e];
    s_o[warp][2 * pr + 1] = o[2 * e + 1];
  }
  __syncthreads();
  float m_all = -INFINITY;
#pragma unroll
  for (int w = 0; w < NW; ++w) m_all = fmaxf(m_all, s_m[w]);
  float l_all = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) l_all += (s_m[w] == -INFINITY) ? 0.f : s_l[w] * __expf(s_m[w] - m_all);
  if (p.nsplit > 1) {  // un-normalised partial (m, l, o[D]) for the merge
    float* wp = p.ws + ((static_cast<long long>(b) * gridDim.x + h) * p.nsplit + blockIdx.z) * (DMAX + 2);
    if (tid == 0) {
      wp[0] = m_all;
      wp[1] = l_all;
    }
    for (int d = tid; d < p.D; d += 128) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) acc += (s_m[w] == -INFINITY) ? 0.f : s_o[w][d] * __expf(s_m[w] - m_all);
      wp[2 + d] = acc;
    }
    if (p.counters == nullptr) return;  // separate merge kernel follows
    // fused merge: the last block to finish for this (b, h) combines all partials (threadfence + atomic ticket)
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int prev = atomicAdd(p.counters + b * gridDim.x + h, 1);
      s_last = (prev == p.nsplit - 1);
      if (s_last) p.counters[b * gridDim.x + h] = 0;  // self-cleaning for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* base = p.ws + (static_cast<long long>(b) * gridDim.x + h) * p.nsplit * (DMAX + 2);
    float mm = -INFINITY;
    for (int sidx = 0; sidx < p.nsplit; ++sidx) mm = fmaxf(mm, __ldcg(base + sidx * (DMAX + 2)));
    float ll = 0.f;
    for (int sidx = 0; sidx < p.nsplit; ++sidx) {
      const float m = __ldcg(base + sidx * (DMAX + 2));
      ll += (m == -INFINITY) ? 0.f : __ldcg(base + sidx * (DMAX + 2) + 1) * __expf(m - mm);
    }
    __nv_bfloat16* opm = p.o + b * p.o_bs + h * p.o_hs;
    for (int d = tid; d < p.D; d += 128) {
      float acc = 0.f;
      for (int sidx = 0; sidx < p.nsplit; ++sidx) {
        const float m = __ldcg(base + sidx * (DMAX + 2));
        acc += (m == -INFINITY) ? 0.f : __ldcg(base + sidx * (DMAX + 2) + 2 + d) * __expf(m - mm);
      }
      opm[d] = __float2bfloat16(ll > 0.f ? acc / ll : 0.f);
    }
    return;
  }
  __nv_bfloat16* op = p.o + b * p.o_bs + h * p.o_hs;
  for (int d = tid; d < p.D; d += 128) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) acc += (s_m[w] == -INFINITY) ? 0.f : s_o[w][d] * __expf(s_m[w] - m_all);
    op[d] = __float2bfloat16(l_all > 0.f ? acc / l_all : 0.f);
  }
}

// Version 2 (head dim <= 128): ncu showed version 1 latency-bound (256 CTAs, 10 % warps active, 25 us for 21 MB: every warp walks its keys
// through ~20 dependent round trips of small loads). Here each block owns <= 64 keys of one head and issues ALL of its K and V bytes as
// 16-byte cp.async copies up front (24 KB in flight per block, ~6 blocks per SM), then does QK^T, softmax and PV out of shared memory.
// Grid (H, B, nsplit) with nsplit ~ keys / 64; partials are merged by the last block of each head (ticket counter) or a merge launch.
constexpr int DEC2_KEYS = 64;
__global__ void __launch_bounds__(128) attn_decode2_kernel(AttnParams p) {
  constexpr int DMAX = 128, LDS = DMAX + 8;  // padded row (bf16 elements)
  pdl_launch_dependents();
  pdl_wait();  // q / the newest cache row come from the preceding kernels; nk_dev is advanced by the previous decode step
  if (p.nk_dev != nullptr) p.Nk = p.nk_dev[p.nk_per_batch ? blockIdx.y : 0];
  __shared__ __align__(16) __nv_bfloat16 s_k[DEC2_KEYS * LDS];
  __shared__ __align__(16) __nv_bfloat16 s_v[DEC2_KEYS * LDS];
  __shared__ __align__(16) float s_q[DMAX];
  __shared__ float s_p[DEC2_KEYS];
  __shared__ float s_o[2][DMAX];
  __shared__ float s_red[4];
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __nv_bfloat16* qp = p.q + b * p.q_bs + h * p.q_hs;
  const __nv_bfloat16* kb = p.k + b * p.k_bs + h * p.k_hs;
  const __nv_bfloat16* vb = p.v + b * p.v_bs + h * p.v_hs;
  int k_begin = 0;
  if (p.window > 0) k_begin = max(0, p.Nk - 1 - p.window);
  const int span = ((p.Nk - k_begin + p.nsplit - 1) / p.nsplit + 31) / 32 * 32;
  const int k_lo = k_begin + blockIdx.z * span;
  const int k_hi = min(p.Nk, k_lo + span);
  const int chunks = p.D >> 3;  // 16-byte chunks per row
  float m_run = -INFINITY, l_run = 0.f, o0 = 0.f, o1 = 0.f;  // thread (kh = tid / 64, pr = tid % 64) accumulates feature pair pr over its key half
  const int half = p.D >> 1;
  if (p.rope_theta > 0.f) {
    // fused RoPE: rotate-half at position Nk-1, same arithmetic as rope_kv_append_kernel / rope_neox_kernel
    for (int d = tid; d < DMAX; d += 128) {
      float v = 0.f;
      if (d < p.D) {
        const int j = d < half ? d : d - half;
        float sn, cs;
        sincosf((p.Nk - 1) * powf(p.rope_theta, -2.0f * j / p.D), &sn, &cs);
        const float a = __bfloat162float(qp[j]), bq = __bfloat162float(qp[j + half]);
        // the un-fused path rounds the rotated q to bf16 before the dot product: keep that rounding so both paths give identical logits
        v = __bfloat162float(__float2bfloat16(d < half ? a * cs - bq * sn : bq * cs + a * sn)) * p.scale;
      }
      s_q[d] = v;
    }
  } else {
    for (int d = tid; d < DMAX; d += 128) s_q[d] = d < p.D ? __bfloat162float(qp[d]) * p.scale : 0.f;
  }
  for (int k0 = k_lo; k0 < k_hi; k0 += DEC2_KEYS) {
    const int n = min(DEC2_KEYS, k_hi - k0);
    __syncthreads();  // previous tile fully consumed (and s_q written)
    const int ops = n * chunks;
    for (int i = tid; i < 2 * ops; i += 128) {
      const int which = i >= ops, j = which ? i - ops : i;
      const int key = j / chunks, c = j - key * chunks;
      const __nv_bfloat16* src = (which ? vb + static_cast<long long>(k0 + key) * p.v_rs : kb + static_cast<long long>(k0 + key) * p.k_rs) + c * 8;
      cp_async16((which ? s_v : s_k) + key * LDS + c * 8, src, true);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (p.rope_theta > 0.f && p.Nk - 1 >= k0 && p.Nk - 1 < k0 + n) {
      // this block owns the newest position: its K/V row is not in the cache yet. Rotate k_new, take v_new from the packed row, use them
      // from shared memory and append them to the cache for the following steps (exactly one block per head gets here).
      const int rowi = p.Nk - 1 - k0;
      const __nv_bfloat16* kn = qp + static_cast<long long>(p.H) * p.D;       // k_new of this head
      const __nv_bfloat16* vn = qp + 2LL * p.H * p.D;                          // v_new of this head
      __nv_bfloat16* kc = const_cast<__nv_bfloat16*>(kb) + static_cast<long long>(p.Nk - 1) * p.k_rs;
      __nv_bfloat16* vc = const_cast<__nv_bfloat16*>(vb) + static_cast<long long>(p.Nk - 1) * p.v_rs;
      for (int d = tid; d < p.D; d += 128) {
        const int j = d < half ? d : d - half;
        float sn, cs;
        sincosf((p.Nk - 1) * powf(p.rope_theta, -2.0f * j / p.D), &sn, &cs);
        const float a = __bfloat162float(kn[j]), bk = __bfloat162float(kn[j + half]);
        const __nv_bfloat16 kr = __float2bfloat16(d < half ? a * cs - bk * sn : bk * cs + a * sn);
        const __nv_bfloat16 vr = vn[d];
        s_k[rowi * LDS + d] = kr;
        s_v[rowi * LDS + d] = vr;
        kc[d] = kr;
        vc[d] = vr;
      }
      __syncthreads();
    }
    // QK^T: two threads per key (feature halves), combined with one shuffle
    {
      const int key = tid >> 1, half = tid & 1;
      float acc = 0.f;
      if (key < n) {
        const int c_lo = half ? (chunks + 1) / 2 : 0, c_hi = half ? chunks : (chunks + 1) / 2;
        for (int c = c_lo; c < c_hi; ++c) {
          const uint4 kv = *reinterpret_cast<const uint4*>(s_k + key * LDS + c * 8);
          const float4 qa = *reinterpret_cast<const float4*>(s_q + c * 8), qb = *reinterpret_cast<const float4*>(s_q + c * 8 + 4);
          acc = fmaf(bf16_lo(kv.x), qa.x, acc); acc = fmaf(bf16_hi(kv.x), qa.y, acc);
          acc = fmaf(bf16_lo(kv.y), qa.z, acc); acc = fmaf(bf16_hi(kv.y), qa.w, acc);
          acc = fmaf(bf16_lo(kv.z), qb.x, acc); acc = fmaf(bf16_hi(kv.z), qb.y, acc);
          acc = fmaf(bf16_lo(kv.w), qb.z, acc); acc = fmaf(bf16_hi(kv.w), qb.w, acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      const float sc = key < n ? acc : -INFINITY;
      const float wm = warp_max(sc);
      if (lane == 0) s_red[warp] = wm;
      __syncthreads();
      const float m_new = fmaxf(m_run, fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3])));
      const float pj = key < n ? __expf(sc - m_new) : 0.f;
      if (half == 0) s_p[key] = pj;
      const float ws = warp_sum(half == 0 ? pj : 0.f);
      __syncthreads();  // s_red reads done, s_p written
      if (lane == 0) s_red[warp] = ws;
      __syncthreads();
      const float alpha = (m_run == -INFINITY) ? 0.f : __expf(m_run - m_new);
      l_run = l_run * alpha + (s_red[0] + s_red[1] + s_red[2] + s_red[3]);
      m_run = m_new;
      o0 *= alpha;
      o1 *= alpha;
    }
    // PV: thread = (key half, feature pair)
    {
      const int kh = tid >> 6, pr = tid & 63;
      if (2 * pr < p.D) {
        const int j_lo = kh * (DEC2_KEYS / 2), j_hi = min(n, j_lo + DEC2_KEYS / 2);
#pragma unroll 8
        for (int j = j_lo; j < j_hi; ++j) {
          const uint32_t vv = *reinterpret_cast<const uint32_t*>(s_v + j * LDS + 2 * pr);
          const float pj = s_p[j];
          o0 = fmaf(pj, bf16_lo(vv), o0);
          o1 = fmaf(pj, bf16_hi(vv), o1);
        }
      }
    }
  }
  __syncthreads();
  {
    const int kh = tid >> 6, pr = tid & 63;
    s_o[kh][2 * pr] = o0;
    s_o[kh][2 * pr + 1] = o1;
  }
  __syncthreads();
  if (p.nsplit > 1) {  // un-normalised partial (m, l, o[D]) for the merge
    float* wp = p.ws + ((static_cast<long long>(b) * gridDim.x + h) * p.nsplit + blockIdx.z) * (DMAX + 2);
    if (tid == 0) {
      wp[0] = m_run;
      wp[1] = l_run;
    }
    for (int d = tid; d < p.D; d += 128) wp[2 + d] = s_o[0][d] + s_o[1][d];
    if (p.counters == nullptr) return;  // separate merge kernel follows
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int prev = atomicAdd(p.counters + b * gridDim.x + h, 1);
      s_last = (prev == p.nsplit - 1);
      if (s_last) p.counters[b * gridDim.x + h] = 0;  // self-cleaning for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* base = p.ws + (static_cast<long long>(b) * gridDim.x + h) * p.nsplit * (DMAX + 2);
    // weights of the partials: lane s < nsplit of warp 0 owns partial s (nsplit <= 32), published through shared memory
    if (warp == 0) {
      const float m_s = lane < p.nsplit ? __ldcg(base + lane * (DMAX + 2)) : -INFINITY;
      const float mm = warp_max(m_s);
      const float w_s = (m_s == -INFINITY) ? 0.f : __expf(m_s - mm);
      const float ll = warp_sum(lane < p.nsplit ? __ldcg(base + lane * (DMAX + 2) + 1) * w_s : 0.f);
      s_p[lane] = w_s;
      if (lane == 0) s_red[0] = ll;
    }
    __syncthreads();
    const float ll = s_red[0];
    __nv_bfloat16* opm = p.o + b * p.o_bs + h * p.o_hs;
    for (int d = tid; d < p.D; d += 128) {
      float acc = 0.f;
      for (int sidx = 0; sidx < p.nsplit; ++sidx) acc = fmaf(__ldcg(base + sidx * (DMAX + 2) + 2 + d), s_p[sidx], acc);
      opm[d] = __float2bfloat16(ll > 0.f ? acc / ll : 0.f);
    }
    return;
  }
  __nv_bfloat16* op = p.o + b * p.o_bs + h * p.o_hs;
  for (int d = tid; d < p.D; d += 128) op[d] = __float2bfloat16(l_run > 0.f ? (s_o[0][d] + s_o[1][d]) / l_run : 0.f);
}

template <int DMAX>
__global__ void attn_decode_merge_kernel(const AttnParams p) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const int h = blockIdx.x, b = blockIdx.y;
  const float* wp = p.ws + (static_cast<long long>(b) * gridDim.x + h) * p.nsplit * (DMAX + 2);
  float m_all = -INFINITY;
  for (int s = 0; s < p.nsplit; ++s) m_all = fmaxf(m_all, wp[s * (DMAX + 2)]);
  float l_all = 0.f;
  for (int s = 0; s < p.nsplit; ++s) {
    const float m = wp[s * (DMAX + 2)];
    l_all += (m == -INFINITY) ? 0.f : wp[s * (DMAX + 2) + 1] * __expf(m - m_all);
  }
  __nv_bfloat16* op = p.o + b * p.o_bs + h * p.o_hs;
  for (int d = threadIdx.x; d < p.D; d += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < p.nsplit; ++s) {
      const float m = wp[s * (DMAX + 2)];
      acc += (m == -INFINITY) ? 0.f : wp[s * (DMAX + 2) + 2 + d] * __expf(m - m_all);
    }
    op[d] = __float2bfloat16(l_all > 0.f ? acc / l_all : 0.f);
  }
}

// one warp per (b, h, row): combine nsplit partials
template <int DP>
__global__ void flash_attn_merge_kernel(const AttnParams p) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const long long gw = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long rows = static_cast<long long>(p.B) * p.H * p.Nq;
  if (gw >= rows) return;
  const int row = static_cast<int>(gw % p.Nq);
  const int h = static_cast<int>((gw / p.Nq) % p.H);
  const int b = static_cast<int>(gw / (static_cast<long long>(p.Nq) * p.H));
  float m_all = -INFINITY;
  for (int s = 0; s < p.nsplit; ++s) m_all = fmaxf(m_all, p.ws[(static_cast<long long>(s) * rows + gw) * (DP + 2)]);
  float l_all = 0.f, wgt[16];
  for (int s = 0; s < p.nsplit; ++s) {
    const float* wp = p.ws + (static_cast<long long>(s) * rows + gw) * (DP + 2);
    wgt[s] = (wp[0] == -INFINITY) ? 0.f : __expf(wp[0] - m_all);
    l_all += wp[1] * wgt[s];
  }
  const float inv = l_all > 0.f ? 1.0f / l_all : 0.f;
  __nv_bfloat16* orow = p.o + b * p.o_bs + h * p.o_hs + static_cast<long long>(row) * p.o_rs;
  for (int d = lane; d < p.D; d += 32) {
    float acc = 0.f;
    for (int s = 0; s < p.nsplit; ++s) acc += p.ws[(static_cast<long long>(s) * rows + gw) * (DP + 2) + 2 + d] * wgt[s];
    orow[d] = __float2bfloat16(acc * inv);
  }
}

template <int DP>
static int launch_attn(AttnParams p, cudaStream_t s, void* workspace, long long workspace_bytes) {
  constexpr int LD = DP + 8;
  constexpr int smem = (64 * LD + 4 * 64 * LD) * 2;
  const int q_tiles = (p.Nq + 63) / 64;
  const long long ctas = static_cast<long long>(q_tiles) * p.H * p.B;
  p.nsplit = 1;
  if (workspace != nullptr && !p.causal && p.window == 0 && p.q_win == 0 && p.k_win == 0 && ctas < 148 && p.Nk >= 2048) {
    int ns = static_cast<int>((2 * 148 + ctas - 1) / ctas);
    if (ns > 8) ns = 8;
    if (ns > p.Nk / 1024) ns = p.Nk / 1024;
    const long long need = static_cast<long long>(ns) * p.B * p.H * p.Nq * (DP + 2) * 4;
    if (ns > 1 && need <= workspace_bytes) {
      p.nsplit = ns;
      p.ws = static_cast<float*>(workspace);
    }
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(VG_ECUDA, "attention smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid(q_tiles * p.nsplit, p.H, p.B);
  launch_k(flash_attn_kernel<DP>, dim3(grid), dim3(128), smem, s, p);
  count_launch();
  if (p.nsplit > 1) {
    const long long rows = static_cast<long long>(p.B) * p.H * p.Nq;
    launch_k(flash_attn_merge_kernel<DP>, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, s, p);
    count_launch();
  }
  return check_launch("flash_attn_kernel");
}

}  // namespace vg

namespace vg {
int attention_tc_try(const VgAttention* a, cudaStream_t s, int* nsplit_out, int* dpm_out);  // attention_tc.cu
}

using namespace vg;

extern "C" int vg_attention(const VgAttention* a, void* stream) {
  if (!a || !a->q || !a->k || !a->v || !a->o) return set_error(VG_EINVAL, "vg_attention: null pointer");
  if (a->B <= 0 || a->H <= 0 || a->Nq <= 0 || a->Nk <= 0 || a->D <= 0) return set_error(VG_EINVAL, "vg_attention: bad dims");
  if (a->D % 8) return set_error(VG_EINVAL, "vg_attention: head dim must be a multiple of 8 (got %d)", a->D);
  const long long strides[] = {a->q_bs, a->q_hs, a->q_rs, a->k_bs, a->k_hs, a->k_rs, a->v_bs, a->v_hs, a->v_rs};
  for (long long st : strides)
    if (st % 8) return set_error(VG_EINVAL, "vg_attention: q/k/v strides must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v)) & 15)
    return set_error(VG_EINVAL, "vg_attention: q/k/v must be 16-byte aligned");
  if ((a->o_rs | a->o_hs | a->o_bs) % 2 || (reinterpret_cast<uintptr_t>(a->o) & 3)) return set_error(VG_EINVAL, "vg_attention: o strides must be even");
  AttnParams p;
  p.q = static_cast<const __nv_bfloat16*>(a->q); p.k = static_cast<const __nv_bfloat16*>(a->k);
  p.v = static_cast<const __nv_bfloat16*>(a->v); p.o = static_cast<__nv_bfloat16*>(a->o);
  p.q_bs = a->q_bs; p.q_hs = a->q_hs; p.q_rs = a->q_rs; p.k_bs = a->k_bs; p.k_hs = a->k_hs; p.k_rs = a->k_rs;
  p.v_bs = a->v_bs; p.v_hs = a->v_hs; p.v_rs = a->v_rs; p.o_bs = a->o_bs; p.o_hs = a->o_hs; p.o_rs = a->o_rs;
  p.B = a->B; p.H = a->H; p.Nq = a->Nq; p.Nk = a->Nk; p.D = a->D;
  p.scale = a->scale; p.causal = a->causal; p.window = a->window;
  p.q_win = a->q_win; p.q_gridw = a->q_gridw; p.q_nwx = a->q_nwx; p.q_nwy = a->q_nwy;
  p.k_win = a->k_win; p.k_gridw = a->k_gridw; p.k_nwx = a->k_nwx; p.k_nwy = a->k_nwy;
  p.ws = nullptr; p.nsplit = 1; p.nk_dev = static_cast<const int*>(a->nk_dev); p.counters = nullptr; p.rope_theta = 0.f;
  p.nk_per_batch = a->nk_per_batch;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->nk_dev != nullptr && a->Nq != 1) return set_error(VG_EINVAL, "vg_attention: nk_dev is only supported for single-query attention");
  if (a->Nq == 1 && a->q_win == 0 && a->k_win == 0 && (a->Nk >= 256 || a->nk_dev != nullptr)) {  // decode: causal is a no-op for the newest query
    const int dmax = a->D <= 128 ? 128 : 256;
    if (a->fused_rope_theta > 0.f) {
      if (dmax != 128 || a->q_hs != a->D || (a->D % 2)) return set_error(VG_EINVAL, "vg_attention: fused RoPE/append needs D <= 128 and packed [q|k|v] rows");
      p.rope_theta = a->fused_rope_theta;
    }
    int nsplit = 1;
    if (a->workspace != nullptr) {
      nsplit = dmax == 128 ? (a->Nk + DEC2_KEYS - 1) / DEC2_KEYS : (a->Nk + 255) / 256;
      const int cap = dmax == 128 ? 32 : 16;
      if (nsplit > cap) nsplit = cap;
      const long long need = static_cast<long long>(a->B) * a->H * nsplit * (dmax + 2) * 4 + (a->ws_has_counters ? static_cast<long long>(a->B) * a->H * 4 : 0);
      if (need > a->workspace_bytes) nsplit = 1;
    }
    p.ws = static_cast<float*>(a->workspace);
    p.nsplit = nsplit;
    if (nsplit > 1 && a->ws_has_counters)  // zeroed ticket counters live in the last B*H ints of the workspace
      p.counters = reinterpret_cast<int*>(static_cast<char*>(a->workspace) + a->workspace_bytes - static_cast<long long>(a->B) * a->H * 4);
    dim3 grid(a->H, a->B, nsplit);
    if (dmax == 128)
    {
      const cudaError_t e = launch_pdl(attn_decode2_kernel, grid, dim3(128), 0, s, p);
      if (e != cudaSuccess) return set_error(VG_ECUDA, "attn_decode2 launch: %s", cudaGetErrorString(e));
    }
    else
      launch_k(attn_decode_kernel<256>, dim3(grid), dim3(128), 0, s, p);
    count_launch();
    if (nsplit > 1 && p.counters == nullptr) {
      dim3 mg(a->H, a->B);
      if (dmax == 128)
        launch_k(attn_decode_merge_kernel<128>, dim3(mg), dim3(128), 0, s, p);
      else
        launch_k(attn_decode_merge_kernel<256>, dim3(mg), dim3(128), 0, s, p);
      count_launch();
    }
    return check_launch("attn_decode_kernel");
  }
  if (!a->force_legacy) {
    const int wrc = window_attn_small_try(a, s);
    if (wrc != VG_ENOSUP) return wrc;
  }
  if (!a->force_legacy) {  // tcgen05 path (large dense problems); falls through to the mma.sync kernel when not applicable
    int ns = 1, dpm = 128;
    VgAttention at = *a;
    if (at.window > 0 && at.causal && at.Nk <= at.window + 1) at.window = 0;  // band j >= i - window never cuts anything: plain causal
    const int rc = attention_tc_try(&at, s, &ns, &dpm);
    if (rc == VG_OK) {
      if (ns > 1) {
        p.ws = static_cast<float*>(a->workspace);
        p.nsplit = ns;
        const long long rows = static_cast<long long>(p.B) * p.H * p.Nq;
        if (dpm == 128)
          launch_k(flash_attn_merge_kernel<128>, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, s, p);
        else
          launch_k(flash_attn_merge_kernel<256>, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, s, p);
        count_launch();
        return check_launch("flash_attn_merge_kernel");
      }
      return VG_OK;
    }
    if (rc != VG_ENOSUP) return rc;
  }
  if (a->D <= 32) return launch_attn<32>(p, s, a->workspace, a->workspace_bytes);
  if (a->D <= 64) return launch_attn<64>(p, s, a->workspace, a->workspace_bytes);
  if (a->D <= 80) return launch_attn<80>(p, s, a->workspace, a->workspace_bytes);
  if (a->D <= 96) return launch_attn<96>(p, s, a->workspace, a->workspace_bytes);
  if (a->D <= 128) return launch_attn<128>(p, s, a->workspace, a->workspace_bytes);
  if (a->D <= 256) return launch_attn<256>(p, s, a->workspace, a->workspace_bytes);
  return set_error(VG_ENOSUP, "vg_attention: head dim %d > 256", a->D);
}
