// tcgen05 flash attention (sm_100a): S = Q K^T and O += P V on the 5th-gen tensor cores with accumulators in TMEM.
//   warp 0 (1 lane)  TMA producer: Q tile once, then the K ring     warp 3 (1 lane)  TMA producer: V ring   (4D tensor maps over the strided
//                    q/k/v views — linear rows or Hiera 16x16 windows as {D, H, x, y} boxes —, head-dim tail zero-filled by TMA, SWIZZLE_128B)
//   warp 1 (1 lane)  MMA issuer, order QK(0), QK(1), PV(0), QK(2), PV(1), ...: S is double-buffered in TMEM, so the softmax of tile j never waits
//                    for a QK^T and PV(j) runs under the softmax of tile j+1. S[128 x BN] = Q K^T (both K-major); O[128 x Dp] += P V (P K-major
//                    from smem, V MN-major)
//   warp 2           TMEM allocator
//   warps 4..7       softmax: thread r owns query row r (TMEM lane r): tcgen05.ld S, masks (boundary tiles only), online max with LAZY rescaling
//                    of O (O is only rescaled in TMEM when a row max grows by more than 2^8), P -> bf16 -> swizzled smem, final O / l.
// 64-key tiles; when O needs <= 128 TMEM columns the block allocates 256 columns and ~112 KB of shared memory so that TWO blocks share an SM.
// Optional split-K over keys for problems with few query tiles (SAM2 memory attention: 4096 queries x <= 28.7k keys, 1 head).
// History and measurements of the variants: profiles/r01_attention.md.
#include <mutex>
#include <unordered_map>

#include "common.h"
#include "pdl.cuh"
#include "ptx.cuh"

namespace vg {

struct AttnTcParams {
  int B, H, Nq, Nk, D, Dp, nblk;
  float scale;
  int causal, nsplit;
  __nv_bfloat16* o;
  long long o_bs, o_hs, o_rs;
  float* ws;  // split-K partials [split][b][h][row][DPM + 2]
  int dpm;    // row width of the partials (128 or 256)
  int win, gridw, nwx;  // Hiera 2D windows (win > 0): batch entry b is window (b / nwx, b % nwx) of a [rows_y, gridw] token grid; tile rows are
                        // (ty, tx) = (r / win, r % win) and the tensor maps are {D, H, x, y} with boxes of win x (rows / win) tokens
  int window;           // > 0 (causal only): query i additionally sees only keys j >= i + (Nk - Nq) - window (Phi-3 sliding window, transformers 4.41)
  int s_col, tmem_cols; // TMEM column of S and the allocation size: 128/256 when O needs <= 128 columns (two blocks per SM), else 256/512
};

template <int BN, int MINB>
__global__ void __launch_bounds__(256, MINB)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023) != 0) __trap();
  constexpr int BM = 128;
  const int kv_blk_bytes = BN * 128;           // one 64-feature block of a K or V tile
  const int kv_stage_bytes = p.nblk * kv_blk_bytes;
  uint8_t* sQ = smem;                          // nblk x [128 rows x 128 B]
  uint8_t* sK = sQ + p.nblk * 16384;           // 2 stages
  uint8_t* sV = sK + 2 * kv_stage_bytes;       // 2 stages
  uint8_t* sP = sV + 2 * kv_stage_bytes;       // (BN / 64) x [128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (BN / 64) * 16384);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]  S = Q K^T of tile j landed in TMEM buffer j & 1
  uint64_t* s_empty = bars + 11;  // [2]  the softmax threads have read that buffer
  uint64_t* p_full = bars + 13;   //      P of tile j is in shared memory
  uint64_t* p_empty = bars + 14;  //      PV of tile j retired: P may be overwritten, O may be rescaled
  uint64_t* o_full = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tiles = (p.Nq + BM - 1) / BM;
  // causal: the query tiles with the most key tiles of EVERY head are issued first (blocks are scheduled in blockIdx order; with the light
  // tiles first the heaviest blocks formed the tail: Phi-3 prefill 66 -> 53 us)
  int bx = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  if (p.causal) {   // 1-D launch: block i = (query tile, descending) x (head, batch)
    const int hbn = p.H * p.B, xr = static_cast<int>(blockIdx.x) / hbn, r = static_cast<int>(blockIdx.x) - xr * hbn;
    bx = q_tiles * p.nsplit - 1 - xr;
    b = r / p.H;
    h = r - b * p.H;
  }
  const int split = bx / q_tiles;
  const int q0 = (bx - split * q_tiles) * BM;
  const int off = p.Nk - p.Nq;
  int k_begin = 0, k_end = p.Nk;
  if (p.causal) k_end = min(p.Nk, q0 + BM + off);
  if (p.window > 0) k_begin = max(0, q0 + off - p.window) / BN * BN;  // the first row of the tile reaches furthest back
  if (p.nsplit > 1) {
    const int per = ((p.Nk + p.nsplit - 1) / p.nsplit + BN - 1) / BN * BN;
    k_begin = split * per;
    k_end = min(k_end, k_begin + per);
  }
  const int ntiles = k_end > k_begin ? (k_end - k_begin + BN - 1) / BN : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 128);
    }
    mbar_init(p_full, 128);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();  // q/k/v are the predecessor's outputs; the prologue above overlapped its tail
  const uint32_t tO = tmem_base;             // columns [0, Dp)
  const uint32_t tS = tmem_base + p.s_col;   // two S buffers: columns [s_col + st * BN, s_col + (st + 1) * BN)
  // tensor-map coordinates of a tile whose first row is r0: linear (row, batch) or windowed (x, y)
  const int wx0 = p.win ? (b % p.nwx) * p.win : 0, wy0 = p.win ? (b / p.nwx) * p.win : 0;
  auto c2 = [&](int r0) { return p.win ? wx0 : r0; };
  auto c3 = [&](int r0) { return p.win ? wy0 + r0 / p.win : b; };

  if (warp == 0) {
    if (lane == 0 && ntiles > 0) {
      // ------------------------------------------------ TMA producer: Q once, then the K ring (slot free as soon as its QK^T retired)
      mbar_arrive_expect_tx(q_full, p.nblk * 16384);
      for (int blk = 0; blk < p.nblk; ++blk) tma_load_4d(&tmQ, q_full, sQ + blk * 16384, blk * 64, h, c2(q0), c3(q0));
      for (int j = 0; j < ntiles; ++j) {
        const int st = j & 1;
        mbar_wait(&k_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[st], kv_stage_bytes);
        const int k0 = k_begin + j * BN;
        for (int blk = 0; blk < p.nblk; ++blk)
          tma_load_4d(&tmK, &k_full[st], sK + st * kv_stage_bytes + blk * kv_blk_bytes, blk * 64, h, c2(k0), c3(k0));
      }
    }
  } else if (warp == 3) {
    if (lane == 0 && ntiles > 0) {
      // ------------------------------------------------ TMA producer for V (slot free when its PV retired)
      for (int j = 0; j < ntiles; ++j) {
        const int st = j & 1;
        mbar_wait(&v_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[st], kv_stage_bytes);
        const int k0 = k_begin + j * BN;
        for (int blk = 0; blk < p.nblk; ++blk)
          tma_load_4d(&tmV, &v_full[st], sV + st * kv_stage_bytes + blk * kv_blk_bytes, blk * 64, h, c2(k0), c3(k0));
      }
    }
  } else if (warp == 1) {
    if (ntiles > 0) {
      // ------------------------------------------------ MMA issuer. Issue order QK(0), QK(1), PV(0), QK(2), PV(1), ...: S is double-buffered
      // in TMEM so that the softmax of tile j never waits for a QK^T, and PV(j) runs under the softmax of tile j + 1. The whole warp runs the
      // loop converged and one elected lane issues (umma_bf16_elect): no per-instruction ELECT loop in the issuing thread.
      const uint32_t idesc_s = umma_idesc_bf16(BM, BN);
      const uint32_t idesc_o = umma_idesc_bf16_bmn(BM, p.Dp);
      mbar_wait(q_full, 0);
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        if (j >= 2) mbar_wait(&s_empty[st], ((j >> 1) - 1) & 1);
        tc_fence_after();
        // S = Q K^T over the (zero-padded) head dim, 16 features per MMA
        for (int blk = 0; blk < p.nblk; ++blk) {
          const int steps = min(4, (p.Dp - 64 * blk) >> 4);
          const uint64_t qd = umma_desc_sw128(smem_u32(sQ + blk * 16384));
          const uint64_t kd = umma_desc_sw128(smem_u32(sK + st * kv_stage_bytes + blk * kv_blk_bytes));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < steps) umma_bf16_elect(tS + st * BN, qd + 2 * k, kd + 2 * k, idesc_s, (blk | k) != 0);
        }
        umma_commit_elect(&k_empty[st]);
        umma_commit_elect(&s_full[st]);
      };
      issue_qk(0);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j & 1;
        if (j + 1 < ntiles) issue_qk(j + 1);
        mbar_wait(&v_full[st], (j >> 1) & 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        // O += P V: 16 keys per MMA; V is MN-major (features contiguous): 8-key groups every 1024 B, 64-feature blocks every BN*128 B
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          const uint64_t pd = umma_desc_sw128(smem_u32(sP + (kk >> 2) * 16384)) + 2 * (kk & 3);
          const uint64_t vd = umma_desc_sw128_mn(smem_u32(sV + st * kv_stage_bytes + kk * 2048), kv_blk_bytes, 1024);
          umma_bf16_elect(tO, pd, vd, idesc_o, (j | kk) != 0);
        }
        umma_commit_elect(&v_empty[st]);
        umma_commit_elect(p_empty);
      }
      umma_commit_elect(o_full);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ softmax / correction / epilogue: one query row per thread
    const int ew = warp & 3;
    const int r = ew * 32 + lane;  // row in the tile == TMEM lane
    const int row = q0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(ew * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const int st = j & 1;
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after();
      const int k0 = k_begin + j * BN;
      const int col_lim = p.causal ? min(p.Nk, row + off + 1) : p.Nk;  // valid columns are < col_lim
      const int col_lo = p.window > 0 ? row + off - p.window : 0;      // ... and >= col_lo (sliding window band)
      // masking is only needed on tiles that cross a key limit of some row of this warp (the last tile / the causal diagonal / the band edge)
      const bool need_mask = __any_sync(0xffffffffu, k0 + BN > col_lim || k0 < col_lo);
      float sv[BN];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t raw[32];
        tmem_ld32(tS + st * BN + lane_addr + c * 32, raw);
        tmem_ld_wait();
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = k0 + c * 32 + i;
            const float x = (col < col_lim && col >= col_lo) ? __uint_as_float(raw[i]) : -INFINITY;
            sv[c * 32 + i] = x;
            mx = fmaxf(mx, x);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            sv[c * 32 + i] = __uint_as_float(raw[i]);
            mx = fmaxf(mx, sv[c * 32 + i]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&s_empty[st]);  // this S buffer may be overwritten by QK^T of tile j + 2
      const float m_new = fmaxf(m_ref, mx);
      // lazy rescale (warp-uniform because tcgen05.ld/st are warp-collective): only when some row's max grew by more than 2^8
      const bool need = (m_ref != -INFINITY) && ((m_new - m_ref) * sl2 > 8.0f);
      if (m_ref == -INFINITY) m_ref = m_new;
      // PV of the previous tile must have retired before O is rescaled or P is overwritten. It was issued when that tile's P arrived, and the
      // issue -> execute -> commit -> mbarrier chain takes ~0.5 us, so the wait is deferred until the exponentials of THIS tile are done
      // (they only need registers); the rare rescale path waits first.
      bool pv_done = (j == 0);
      if (__any_sync(0xffffffffu, need)) {
        if (!pv_done) {
          mbar_wait(p_empty, (j - 1) & 1);
          tc_fence_after();
          pv_done = true;
        }
        const float alpha = need ? ex2_approx((m_ref - m_new) * sl2) : 1.0f;
        for (int c = 0; c < p.Dp; c += 32) {
          uint32_t raw[32];
          tmem_ld32(tO + lane_addr + c, raw);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * alpha);
          tmem_st32(tO + lane_addr + c, raw);
        }
        tmem_st_wait();
        l *= alpha;
        if (need) m_ref = m_new;
      }
      const float msc = (m_ref == -INFINITY) ? 0.f : m_ref * sl2;
      // P = exp2(S*scale*log2e - m) -> bf16 (registers) -> shared memory in the K-major SWIZZLE_128B operand layout
      uint32_t pk[BN / 2];
#pragma unroll
      for (int i = 0; i < BN / 2; ++i) {
        const float p0 = ex2_approx(fmaf(sv[2 * i], sl2, -msc)), p1 = ex2_approx(fmaf(sv[2 * i + 1], sl2, -msc));
        l += p0 + p1;
        pk[i] = pack_bf16(p0, p1);
      }
      if (!pv_done) {
        mbar_wait(p_empty, (j - 1) & 1);
        tc_fence_after();
      }
      uint8_t* prow = sP + r * 128;
      const int sw = r & 7;
#pragma unroll
      for (int c8 = 0; c8 < BN / 8; ++c8)
        *reinterpret_cast<uint4*>(prow + (c8 >> 3) * 16384 + (((c8 & 7) ^ sw) << 4)) = make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
      fence_async_shared();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    if (ntiles > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
    }
    if (p.nsplit > 1) {
      if (row < p.Nq) {
        float* wp = p.ws + ((((static_cast<long long>(split) * p.B + b) * p.H + h) * p.Nq) + row) * (p.dpm + 2);
        wp[0] = (m_ref == -INFINITY) ? -INFINITY : m_ref * p.scale;
        wp[1] = l;
      }
    }
    const float inv = (p.nsplit > 1) ? 1.0f : (l > 0.f ? 1.0f / l : 0.f);
    for (int c = 0; c < p.Dp; c += 32) {
      uint32_t raw[32];
      if (ntiles > 0) {
        tmem_ld32(tO + lane_addr + c, raw);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = 0u;
      }
      if (row >= p.Nq) continue;
      if (p.nsplit > 1) {
        float* wp = p.ws + ((((static_cast<long long>(split) * p.B + b) * p.H + h) * p.Nq) + row) * (p.dpm + 2) + 2 + c;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < p.D) wp[i] = __uint_as_float(raw[i]);
      } else {
        const long long orow = p.win ? static_cast<long long>(wy0 + row / p.win) * p.gridw + wx0 + row % p.win : row;
        __nv_bfloat16* op = p.o + (p.win ? 0 : b * p.o_bs) + h * p.o_hs + orow * p.o_rs + c;
        if (c + 32 <= p.D && (reinterpret_cast<uintptr_t>(op) & 15) == 0) {
#pragma unroll
          for (int i = 0; i < 32; i += 8)
            *reinterpret_cast<uint4*>(op + i) = make_uint4(pack_bf16(__uint_as_float(raw[i]) * inv, __uint_as_float(raw[i + 1]) * inv),
                                                           pack_bf16(__uint_as_float(raw[i + 2]) * inv, __uint_as_float(raw[i + 3]) * inv),
                                                           pack_bf16(__uint_as_float(raw[i + 4]) * inv, __uint_as_float(raw[i + 5]) * inv),
                                                           pack_bf16(__uint_as_float(raw[i + 6]) * inv, __uint_as_float(raw[i + 7]) * inv));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i < p.D) op[i] = __float2bfloat16(__uint_as_float(raw[i]) * inv);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------- two query tiles per block (D <= 128)
// The v2 kernel above leaves the softmax threads waiting ~40 % of the time: P(j+1) can only be written once PV(j) has retired (one P buffer), and
// the block's 112 KB of shared memory leave no room for a second one while two blocks share an SM. Here ONE block per SM owns TWO query tiles
// (256 rows) that share every K/V tile (loaded once, half the TMA / L2 traffic per query): the shared memory that frees pays for double-buffered
// P per tile, so a softmax group never waits for a PV except on the rare O rescale; each tile has its own softmax warpgroup (warps 4-7 / 8-11),
// its own O and double-buffered S in TMEM (O_A 0.., O_B 128.., S_A 256/320, S_B 384/448 = 512 columns), and the single MMA issuer interleaves
// QK_A(j+1) QK_B(j+1) PV_A(j) PV_B(j), so one tile's MMAs run under the other tile's exponentials. Causal / banded tiles of the pair have
// different key ranges: the K/V ring walks the union, each tile takes its sub-range. Blocks are issued heaviest first (causal: last pair first).
template <int BN>
__global__ void __launch_bounds__(384, 1)
flash_attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                      const AttnTcParams p, const int kst) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023) != 0) __trap();
  constexpr int BM = 128;
  static_assert(BN == 64, "P buffers are one 64-key swizzle block");
  const int kv_blk_bytes = BN * 128;
  const int kv_stage_bytes = p.nblk * kv_blk_bytes;
  uint8_t* sQ = smem;                                   // 2 tiles x nblk x [128 rows x 128 B]
  uint8_t* sK = sQ + 2 * p.nblk * 16384;                // kst stages
  uint8_t* sV = sK + kst * kv_stage_bytes;              // kst stages
  uint8_t* sP = sV + kst * kv_stage_bytes;              // [tile][buffer] x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * 16384);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // [4]
  uint64_t* k_empty = bars + 5;       // [4]
  uint64_t* v_full = bars + 9;        // [4]
  uint64_t* v_empty = bars + 13;      // [4]
  uint64_t* s_full = bars + 17;       // [tile][2]
  uint64_t* s_empty = bars + 21;      // [tile][2]
  uint64_t* p_full = bars + 25;       // [tile][2]
  uint64_t* p_empty = bars + 29;      // [tile][2]
  uint64_t* o_full = bars + 33;       // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 34);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_pairs = (p.Nq + 2 * BM - 1) / (2 * BM);
  int pair = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  if (p.causal) {   // 1-D launch: heaviest pairs of every (head, batch) first
    const int hbn = p.H * p.B, xr = static_cast<int>(blockIdx.x) / hbn, r = static_cast<int>(blockIdx.x) - xr * hbn;
    pair = q_pairs - 1 - xr;
    b = r / p.H;
    h = r - b * p.H;
  }
  const int off = p.Nk - p.Nq;
  // key range of each tile of the pair, and the union the K/V ring walks
  int q0x[2], jb[2], je[2];
  int kb = 0, nt = 0;
  {
    int kbx[2], kex[2];
    for (int x = 0; x < 2; ++x) {
      q0x[x] = (pair * 2 + x) * BM;
      kbx[x] = 0;
      kex[x] = p.Nk;
      if (p.causal) kex[x] = min(p.Nk, q0x[x] + BM + off);
      if (p.window > 0) kbx[x] = max(0, q0x[x] + off - p.window) / BN * BN;
      if (q0x[x] >= p.Nq || kex[x] <= kbx[x]) kex[x] = kbx[x] = -1;   // inactive tile
    }
    kb = kbx[0] >= 0 ? kbx[0] : 0;     // tile A reaches furthest back (and is active whenever B is)
    int ke = max(kex[0], kex[1]);
    nt = ke > kb ? (ke - kb + BN - 1) / BN : 0;
    for (int x = 0; x < 2; ++x) {
      jb[x] = kbx[x] >= 0 ? (kbx[x] - kb) / BN : 0;
      je[x] = kbx[x] >= 0 ? jb[x] + (kex[x] - kbx[x] + BN - 1) / BN : 0;
    }
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 128);
      mbar_init(&p_full[s], 128);
      mbar_init(&p_empty[s], 1);
    }
    mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  const int wx0 = p.win ? (b % p.nwx) * p.win : 0, wy0 = p.win ? (b / p.nwx) * p.win : 0;
  auto c2 = [&](int r0) { return p.win ? wx0 : r0; };
  auto c3 = [&](int r0) { return p.win ? wy0 + r0 / p.win : b; };

  if (warp == 0) {
    if (lane == 0 && nt > 0) {
      // ------------------------------------------------ TMA producer: both Q tiles once, then the K ring
      mbar_arrive_expect_tx(q_full, 2 * p.nblk * 16384);
      for (int x = 0; x < 2; ++x)
        for (int blk = 0; blk < p.nblk; ++blk) tma_load_4d(&tmQ, q_full, sQ + (x * p.nblk + blk) * 16384, blk * 64, h, c2(q0x[x]), c3(q0x[x]));
      int st = 0, ph = 0;
      for (int j = 0; j < nt; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], kv_stage_bytes);
        const int k0 = kb + j * BN;
        for (int blk = 0; blk < p.nblk; ++blk)
          tma_load_4d(&tmK, &k_full[st], sK + st * kv_stage_bytes + blk * kv_blk_bytes, blk * 64, h, c2(k0), c3(k0));
        if (++st == kst) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0 && nt > 0) {
      int st = 0, ph = 0;
      for (int j = 0; j < nt; ++j) {
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], kv_stage_bytes);
        const int k0 = kb + j * BN;
        for (int blk = 0; blk < p.nblk; ++blk)
          tma_load_4d(&tmV, &v_full[st], sV + st * kv_stage_bytes + blk * kv_blk_bytes, blk * 64, h, c2(k0), c3(k0));
        if (++st == kst) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (nt > 0) {
      // ------------------------------------------------ MMA issuer: the whole warp runs the loop converged, one elected lane issues (see
      // umma_bf16_elect); stages / phases are carried incrementally and the descriptors are precomputed: the issuing thread was the bottleneck
      // of the first version (ELECT loops, runtime divisions: ~1900 cycles per (query tile, key tile))
      const uint32_t idesc_s = umma_idesc_bf16(BM, BN);
      const uint32_t idesc_o = umma_idesc_bf16_bmn(BM, p.Dp);
      const int steps0 = min(4, p.Dp >> 4), steps1 = p.nblk > 1 ? (p.Dp - 64) >> 4 : 0;
      const uint64_t qd00 = umma_desc_sw128(smem_u32(sQ)), qd01 = umma_desc_sw128(smem_u32(sQ + 16384));
      const uint64_t qd10 = umma_desc_sw128(smem_u32(sQ + p.nblk * 16384)), qd11 = umma_desc_sw128(smem_u32(sQ + (p.nblk + 1) * 16384));
      const uint64_t kd0 = umma_desc_sw128(smem_u32(sK)), pd0 = umma_desc_sw128(smem_u32(sP));
      const uint64_t vd0 = umma_desc_sw128_mn(smem_u32(sV), kv_blk_bytes, 1024);
      const uint32_t stage16 = kv_stage_bytes >> 4, blk16 = kv_blk_bytes >> 4;
      const int jb0 = jb[0], je0 = je[0], jb1 = jb[1], je1 = je[1];
      mbar_wait(q_full, 0);
      int kst_i = 0, kph = 0;                            // K stage / phase of the next QK
      auto issue_qk = [&](int j) {                       // S_x(j) = Q_x K(j)^T for both tiles that reach key tile j
        mbar_wait(&k_full[kst_i], kph);
        const uint64_t kd = kd0 + kst_i * stage16;
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const int jbx = x ? jb1 : jb0, jex = x ? je1 : je0;
          if (j < jbx || j >= jex) continue;
          const int jx = j - jbx, sb = jx & 1;
          if (jx >= 2) mbar_wait(&s_empty[x * 2 + sb], ((jx >> 1) - 1) & 1);
          tc_fence_after();
          const uint32_t tS = tmem_base + 256 + x * 128 + sb * BN;
          const uint64_t q0d = x ? qd10 : qd00, q1d = x ? qd11 : qd01;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < steps0) umma_bf16_elect(tS, q0d + 2 * k, kd + 2 * k, idesc_s, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < steps1) umma_bf16_elect(tS, q1d + 2 * k, kd + blk16 + 2 * k, idesc_s, 1);
          umma_commit_elect(&s_full[x * 2 + sb]);
        }
        umma_commit_elect(&k_empty[kst_i]);
        if (++kst_i == kst) {
          kst_i = 0;
          kph ^= 1;
        }
      };
      issue_qk(0);
      int vst = 0, vph = 0;
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) issue_qk(j + 1);
        mbar_wait(&v_full[vst], vph);
        const uint64_t vd = vd0 + vst * stage16;
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const int jbx = x ? jb1 : jb0, jex = x ? je1 : je0;
          if (j < jbx || j >= jex) continue;
          const int jx = j - jbx, pb = jx & 1;
          mbar_wait(&p_full[x * 2 + pb], (jx >> 1) & 1);
          tc_fence_after();
          const uint32_t tO = tmem_base + x * 128;
          const uint64_t pd = pd0 + (x * 2 + pb) * 1024;          // 16 KB per P buffer
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk) umma_bf16_elect(tO, pd + 2 * kk, vd + kk * 128, idesc_o, (jx | kk) != 0);   // 16 keys = 2048 B of V
          umma_commit_elect(&p_empty[x * 2 + pb]);
        }
        umma_commit_elect(&v_empty[vst]);
        if (++vst == kst) {
          vst = 0;
          vph ^= 1;
        }
      }
      umma_commit_elect(o_full);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ softmax / correction / epilogue of tile x: one query row per thread
    const int x = (warp - 4) >> 2;
    const int ew = warp & 3;
    const int r = ew * 32 + lane;
    const int q0 = q0x[x];
    const int row = q0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t tO = tmem_base + x * 128;
    const float sl2 = p.scale * 1.4426950408889634f;
    const int ntx = je[x] - jb[x];
    float m_ref = -INFINITY, l = 0.f;
    for (int jx = 0; jx < ntx; ++jx) {
      const int sb = jx & 1;
      mbar_wait(&s_full[x * 2 + sb], (jx >> 1) & 1);
      tc_fence_after();
      const uint32_t tS = tmem_base + 256 + x * 128 + sb * BN;
      const int k0 = kb + (jb[x] + jx) * BN;
      const int col_lim = p.causal ? min(p.Nk, row + off + 1) : p.Nk;
      const int col_lo = p.window > 0 ? row + off - p.window : 0;
      const bool need_mask = __any_sync(0xffffffffu, k0 + BN > col_lim || k0 < col_lo);
      // both halves of the S row in flight, ONE wait; the values stay in the registers the loads wrote (no copies)
      uint32_t ra[32], rb[32];
      tmem_ld32(tS + lane_addr, ra);
      tmem_ld32(tS + lane_addr + 32, rb);
      tmem_ld_wait();
      if (need_mask) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int ca = k0 + i, cb = k0 + 32 + i;
          if (!(ca < col_lim && ca >= col_lo)) ra[i] = 0xff800000u;   // -inf
          if (!(cb < col_lim && cb >= col_lo)) rb[i] = 0xff800000u;
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(ra[i]), __uint_as_float(rb[i])));
      tc_fence_before();
      mbar_arrive(&s_empty[x * 2 + sb]);
      const float m_new = fmaxf(m_ref, mx);
      const bool need = (m_ref != -INFINITY) && ((m_new - m_ref) * sl2 > 8.0f);
      if (m_ref == -INFINITY) m_ref = m_new;
      if (__any_sync(0xffffffffu, need)) {
        // rare: O may only be rescaled once every PV issued so far has retired (the commit of PV(jx-1) covers the earlier ones)
        if (jx > 0) {
          mbar_wait(&p_empty[x * 2 + ((jx - 1) & 1)], ((jx - 1) >> 1) & 1);
          tc_fence_after();
        }
        const float alpha = need ? ex2_approx((m_ref - m_new) * sl2) : 1.0f;
        for (int c = 0; c < p.Dp; c += 32) {
          uint32_t raw[32];
          tmem_ld32(tO + lane_addr + c, raw);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * alpha);
          tmem_st32(tO + lane_addr + c, raw);
        }
        tmem_st_wait();
        l *= alpha;
        if (need) m_ref = m_new;
      }
      const float msc = (m_ref == -INFINITY) ? 0.f : m_ref * sl2;
      // P = exp2(S * scale * log2e - m): packed fp32x2 FMA / add (half the issue slots), bf16 packing on the ALU pipe (the XU pipe is the bound)
      uint32_t pk[BN / 2];
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a0, a1, b0, b1;
        ffma2_bc(a0, a1, __uint_as_float(ra[2 * i]), __uint_as_float(ra[2 * i + 1]), sl2, -msc);
        ffma2_bc(b0, b1, __uint_as_float(rb[2 * i]), __uint_as_float(rb[2 * i + 1]), sl2, -msc);
        a0 = ex2_approx(a0); a1 = ex2_approx(a1); b0 = ex2_approx(b0); b1 = ex2_approx(b1);
        fadd2(l0, l1, a0, a1);
        fadd2(l0, l1, b0, b1);
        pk[i] = pack_bf16_alu(a0, a1);
        pk[16 + i] = pack_bf16_alu(b0, b1);
      }
      l += l0 + l1;
      const int pb = jx & 1;
      if (jx >= 2) {                                   // this P buffer was read by PV(jx - 2)
        mbar_wait(&p_empty[x * 2 + pb], ((jx >> 1) - 1) & 1);
        tc_fence_after();
      }
      uint8_t* prow = sP + (x * 2 + pb) * 16384 + r * 128;
      const int sw = r & 7;
#pragma unroll
      for (int c8 = 0; c8 < BN / 8; ++c8)
        *reinterpret_cast<uint4*>(prow + ((c8 ^ sw) << 4)) = make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
      fence_async_shared();
      tc_fence_before();
      mbar_arrive(&p_full[x * 2 + pb]);
    }
    if (nt > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
    }
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    for (int c = 0; c < p.Dp; c += 32) {
      uint32_t raw[32];
      if (ntx > 0) {
        tmem_ld32(tO + lane_addr + c, raw);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = 0u;
      }
      if (row >= p.Nq) continue;
      const long long orow = p.win ? static_cast<long long>(wy0 + row / p.win) * p.gridw + wx0 + row % p.win : row;
      __nv_bfloat16* op = p.o + (p.win ? 0 : b * p.o_bs) + h * p.o_hs + orow * p.o_rs + c;
      if (c + 32 <= p.D && (reinterpret_cast<uintptr_t>(op) & 15) == 0) {
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          *reinterpret_cast<uint4*>(op + i) = make_uint4(pack_bf16(__uint_as_float(raw[i]) * inv, __uint_as_float(raw[i + 1]) * inv),
                                                         pack_bf16(__uint_as_float(raw[i + 2]) * inv, __uint_as_float(raw[i + 3]) * inv),
                                                         pack_bf16(__uint_as_float(raw[i + 4]) * inv, __uint_as_float(raw[i + 5]) * inv),
                                                         pack_bf16(__uint_as_float(raw[i + 6]) * inv, __uint_as_float(raw[i + 7]) * inv));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < p.D) op[i] = __float2bfloat16(__uint_as_float(raw[i]) * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

struct Key4 {
  const void* ptr;
  long long hs, rs, bs;
  int D, H, N, B, rows, rows3;
  bool operator==(const Key4& o) const {
    return ptr == o.ptr && hs == o.hs && rs == o.rs && bs == o.bs && D == o.D && H == o.H && N == o.N && B == o.B && rows == o.rows && rows3 == o.rows3;
  }
};
struct Key4Hash {
  size_t operator()(const Key4& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    for (size_t v : {static_cast<size_t>(k.hs), static_cast<size_t>(k.rs), static_cast<size_t>(k.bs), static_cast<size_t>(k.D), static_cast<size_t>(k.H),
                     static_cast<size_t>(k.N), static_cast<size_t>(k.B), static_cast<size_t>(k.rows), static_cast<size_t>(k.rows3)})
      h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    return h;
  }
};

// bf16 view x[b][n][h][d] with element strides (bs, rs, hs, 1) -> 4D map {D, H, N, B}, box {64, 1, rows, 1}, SWIZZLE_128B, zero OOB fill
// (windowed use: N = grid width, B = stacked grid rows, box {64, 1, win, rows3})
static int make_map4(CUtensorMap* out, const void* ptr, int D, int H, int N, int B, long long hs, long long rs, long long bs, int rows, int rows3 = 1) {
  static std::mutex mu;
  static std::unordered_map<Key4, CUtensorMap, Key4Hash> cache;
  if (H == 1) hs = D;  // strides of size-1 dimensions are irrelevant but must be valid
  if (B == 1) bs = static_cast<long long>(N) * rs;
  Key4 key{ptr, hs, rs, bs, D, H, N, B, rows, rows3};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return VG_OK;
    }
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return set_error(VG_ECUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(hs) * 2, static_cast<cuuint64_t>(rs) * 2, static_cast<cuuint64_t>(bs) * 2};
  cuuint32_t box[4] = {64, 1, static_cast<cuuint32_t>(rows), static_cast<cuuint32_t>(rows3)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(VG_ENOSUP, "attention tensor map rejected (%d)", (int)r);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, *out);
  }
  return VG_OK;
}

template <int BN, int MINB>
static int launch_tc(const VgAttention* a, AttnTcParams& p, cudaStream_t s) {
  CUtensorMap tq, tk, tv;
  int rc;
  if (p.win) {
    // token grid [rows_y = windows_y * win (all images stacked), gridw] -> map {D, H, x, y}, box {64, 1, win, rows / win}
    const int rows_y = a->B / p.nwx * p.win;
    rc = make_map4(&tq, a->q, a->D, a->H, p.gridw, rows_y, a->q_hs, a->q_rs, a->q_rs * p.gridw, p.win, 128 / p.win);
    if (rc) return rc;
    rc = make_map4(&tk, a->k, a->D, a->H, p.gridw, rows_y, a->k_hs, a->k_rs, a->k_rs * p.gridw, p.win, BN / p.win);
    if (rc) return rc;
    rc = make_map4(&tv, a->v, a->D, a->H, p.gridw, rows_y, a->v_hs, a->v_rs, a->v_rs * p.gridw, p.win, BN / p.win);
    if (rc) return rc;
  } else {
    rc = make_map4(&tq, a->q, a->D, a->H, a->Nq, a->B, a->q_hs, a->q_rs, a->q_bs, 128);
    if (rc) return rc;
    rc = make_map4(&tk, a->k, a->D, a->H, a->Nk, a->B, a->k_hs, a->k_rs, a->k_bs, BN);
    if (rc) return rc;
    rc = make_map4(&tv, a->v, a->D, a->H, a->Nk, a->B, a->v_hs, a->v_rs, a->v_bs, BN);
    if (rc) return rc;
  }
  const int smem = p.nblk * 16384 + 4 * p.nblk * BN * 128 + (BN / 64) * 16384 + 256;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_tc_kernel<BN, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error(VG_ECUDA, "attention_tc smem attr: %s", cudaGetErrorString(e));
    attr_smem = 227 * 1024;
  }
  const int q_tiles = (a->Nq + 127) / 128;
  dim3 grid(q_tiles * p.nsplit, a->H, a->B);
  if (p.causal) grid = dim3(q_tiles * p.nsplit * a->H * a->B, 1, 1);
  launch_k(flash_attn_tc_kernel<BN, MINB>, dim3(grid), dim3(256), smem, s, tq, tk, tv, p);
  count_launch();
  return check_launch("flash_attn_tc_kernel");
}

static int launch_tc2(const VgAttention* a, AttnTcParams& p, cudaStream_t s) {
  constexpr int BN = 64;
  CUtensorMap tq, tk, tv;
  int rc;
  if (p.win) {
    const int rows_y = a->B / p.nwx * p.win;
    rc = make_map4(&tq, a->q, a->D, a->H, p.gridw, rows_y, a->q_hs, a->q_rs, a->q_rs * p.gridw, p.win, 128 / p.win);
    if (rc) return rc;
    rc = make_map4(&tk, a->k, a->D, a->H, p.gridw, rows_y, a->k_hs, a->k_rs, a->k_rs * p.gridw, p.win, BN / p.win);
    if (rc) return rc;
    rc = make_map4(&tv, a->v, a->D, a->H, p.gridw, rows_y, a->v_hs, a->v_rs, a->v_rs * p.gridw, p.win, BN / p.win);
    if (rc) return rc;
  } else {
    rc = make_map4(&tq, a->q, a->D, a->H, a->Nq, a->B, a->q_hs, a->q_rs, a->q_bs, 128);
    if (rc) return rc;
    rc = make_map4(&tk, a->k, a->D, a->H, a->Nk, a->B, a->k_hs, a->k_rs, a->k_bs, BN);
    if (rc) return rc;
    rc = make_map4(&tv, a->v, a->D, a->H, a->Nk, a->B, a->v_hs, a->v_rs, a->v_bs, BN);
    if (rc) return rc;
  }
  // K / V ring depth: Q 2 x nblk x 16 KB + P 64 KB + 2 x kst x nblk x 8 KB <= 227 KB. Two stages starve the issuer (a stage is re-requested one
  // tile time before it is needed, less than the L2 -> smem latency: ncu showed the softmax warps 37 % in the s_full wait)
  static const int kst_env = [] { const char* e = getenv("VG_ATTN_KST"); return e ? atoi(e) : 0; }();
  const int kst = kst_env ? kst_env : (p.nblk == 1 ? 4 : 3);
  const int smem = 2 * p.nblk * 16384 + 2 * kst * p.nblk * BN * 128 + 4 * 16384 + 512;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error(VG_ECUDA, "attention_tc2 smem attr: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const int q_pairs = (a->Nq + 255) / 256;
  dim3 grid(q_pairs, a->H, a->B);
  if (p.causal) grid = dim3(q_pairs * a->H * a->B, 1, 1);
  launch_k(flash_attn_tc2_kernel<BN>, dim3(grid), dim3(384), smem, s, tq, tk, tv, p, kst);
  count_launch();
  return check_launch("flash_attn_tc2_kernel");
}

// Returns VG_OK if the tcgen05 kernel handled the call, VG_ENOSUP if the caller should use the generic kernel.
int attention_tc_try(const VgAttention* a, cudaStream_t s, int* nsplit_out, int* dpm_out) {
  *nsplit_out = 1;
  if ((a->window != 0 && !a->causal) || a->window < 0 || a->nk_dev != nullptr) return VG_ENOSUP;
  const bool windowed = a->q_win != 0 || a->k_win != 0;
  if (windowed && !(a->q_win == a->k_win && a->q_gridw == a->k_gridw && a->q_nwx == a->k_nwx && a->q_nwy == a->k_nwy && a->q_win > 0 &&
                    128 % a->q_win == 0 && 64 % a->q_win == 0 && a->Nq == a->q_win * a->q_win && a->Nq % 128 == 0 && !a->causal && a->D <= 128))
    return VG_ENOSUP;  // pooled-query windows and windows smaller than one tile stay on the generic kernel
  if (a->Nq < 128 || a->Nk < 128 || a->D > 256 || (a->D % 8)) return VG_ENOSUP;
  if ((reinterpret_cast<uintptr_t>(a->o) & 3) || (a->o_rs % 2) || (a->o_hs % 2) || (a->o_bs % 2)) return VG_ENOSUP;
  AttnTcParams p;
  p.B = a->B; p.H = a->H; p.Nq = a->Nq; p.Nk = a->Nk; p.D = a->D;
  p.Dp = (a->D + 15) / 16 * 16;
  p.nblk = (p.Dp + 63) / 64;
  p.scale = a->scale; p.causal = a->causal;
  p.o = static_cast<__nv_bfloat16*>(a->o);
  p.o_bs = a->o_bs; p.o_hs = a->o_hs; p.o_rs = a->o_rs;
  p.ws = nullptr; p.nsplit = 1;
  p.dpm = a->D <= 128 ? 128 : 256;
  p.window = a->window;
  p.win = windowed ? a->q_win : 0;
  p.gridw = a->q_gridw;
  p.nwx = a->q_nwx;
  // 64-key tiles: O (<= 128 columns) + S (64) fit a 256-column TMEM allocation and ~112 KB of shared memory, so TWO blocks share an SM and one
  // block's softmax overlaps the other's MMAs (ncu: tensor pipe 16-26 % with one serial block per SM). VG_ATTN_BN=128 restores the wide tile.
  static const int bn_env = [] { const char* e = getenv("VG_ATTN_BN"); return e ? atoi(e) : 64; }();
  const int BN = a->D > 128 ? 64 : (bn_env == 128 ? 128 : 64);
  const bool small_tmem = p.Dp <= 128 && BN == 64;
  p.s_col = small_tmem ? 128 : 256;
  p.tmem_cols = small_tmem ? 256 : 512;
  const long long ctas = static_cast<long long>((a->Nq + 127) / 128) * a->H * a->B;
  if (a->workspace != nullptr && !a->causal && ctas < 148 && a->Nk >= 2048) {
    // key splits: minimise waves x key tiles per split (32 query tiles: 5 splits = 160 blocks = TWO waves of 90 tiles on 148 SMs, 4 splits =
    // one wave of 113)
    const int ns_max = a->Nk / 1024 < 8 ? a->Nk / 1024 : 8;
    const long long key_tiles = (a->Nk + 63) / 64;
    int ns = 1;
    long long best = key_tiles * ((ctas + 147) / 148);
    for (int c = 2; c <= ns_max; ++c) {
      const long long cost = ((ctas * c + 147) / 148) * ((key_tiles + c - 1) / c) + 2;   // + the merge pass
      if (cost < best) {
        best = cost;
        ns = c;
      }
    }
    const long long need = static_cast<long long>(ns) * a->B * a->H * a->Nq * (p.dpm + 2) * 4;
    if (ns > 1 && need <= a->workspace_bytes) {
      p.nsplit = ns;
      p.ws = static_cast<float*>(a->workspace);
    }
  }
  *nsplit_out = p.nsplit;
  *dpm_out = p.dpm;
  // two query tiles per block (double-buffered P, K/V shared by the pair): O of both tiles must fit TMEM next to the S buffers (D <= 128), and
  // the problem must be several waves deep, because one block per SM with half as many blocks loses to two co-resident blocks on the tail
  // (measured, us, pair vs single: Hiera global 8x8x4096^2 566 vs 650; Phi-3 prefill 66 vs 67; IV2 2x16x1025^2 59 vs 47; 16x16 windows 105 vs
  // 106). VG_ATTN_TC2=0 never, 2 whenever possible.
  static const int tc2_env = [] { const char* e = getenv("VG_ATTN_TC2"); return e ? atoi(e) : 1; }();
  if (p.Dp <= 128 && p.nsplit == 1 && BN == 64 && a->Nq > 128 && (tc2_env == 2 || (tc2_env == 1 && !windowed && ctas >= 1024)))
    return launch_tc2(a, p, s);
  if (BN == 128) return launch_tc<128, 1>(a, p, s);
  return small_tmem ? launch_tc<64, 2>(a, p, s) : launch_tc<64, 1>(a, p, s);
}

}  // namespace vg
