// EXPERIMENTAL — NOT part of libvglamm.so (videoglamm_b200/build.py only compiles csrc/*.cu) and NEVER RUN on a GPU yet: written at the end of
// round 1 after the GPU budget was spent, so that round 2 can start from a compiled draft. Compile check:
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v -I videoglamm_b200/csrc -c videoglamm_b200/csrc/experimental/attention_tc2.cu -o /tmp/a.o
//
// Two query tiles per block (the design profiles/r01_attention.md arrives at): the shipped kernel (attention_tc.cu) reaches 24-33 % tensor pipe
// because a block's softmax -> P -> PV -> commit -> wake-up chain (~0.5 us per key tile) can only hide behind the same tile's next step. Here a
// block owns 256 query rows = tiles a and b, each with its own softmax warpgroup, S double buffer and P box, and shares every K/V tile between
// them (half the operand traffic per FLOP). The single MMA thread issues, per key tile j:
//     QK_a(j+1)  QK_b(j+1)  PV_a(j)  PV_b(j)
// so tile a's softmax runs under tile b's MMAs and vice versa. One block per SM (512 TMEM columns, ~200 KB of shared memory).
//   warp 0 (1 lane)  TMA: Q_a, Q_b once, then the K ring          warp 3 (1 lane)  TMA: V ring            (KV_STAGES deep each)
//   warp 1 (1 lane)  MMA issuer                                    warp 2           TMEM allocator
//   warps 4..7       softmax / correction / epilogue of tile a     warps 8..11      same for tile b
// TMEM columns: O_a [0,128)  O_b [128,256)  S_a[0..1] [256,384)  S_b[0..1] [384,512).   Head dim <= 128 only (O_a + O_b must fit 256 columns).
#include "common.h"
#include "pdl.cuh"
#include "ptx.cuh"

namespace vg {

struct AttnTc2Params {
  int B, H, Nq, Nk, D, Dp, nblk;
  float scale;
  int causal, nsplit;
  __nv_bfloat16* o;
  long long o_bs, o_hs, o_rs;
  float* ws;            // split-K partials [split][b][h][row][dpm + 2]
  int dpm;
  int win, gridw, nwx;  // Hiera 2D windows (see attention_tc.cu)
};

constexpr int TC2_BN = 64;         // keys per tile
constexpr int TC2_STAGES = 3;      // K and V ring depth: Q 2 x 32 KB + K/V 2 x 3 x 16 KB + P 2 x 16 KB = 192 KB (D <= 128)
constexpr int TC2_THREADS = 384;

__global__ void __launch_bounds__(TC2_THREADS, 1)
flash_attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                      const AttnTc2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023) != 0) __trap();
  constexpr int BM = 128, BN = TC2_BN, NS = TC2_STAGES;
  const int kv_blk_bytes = BN * 128;                 // one 64-feature block of a K or V tile
  const int kv_stage_bytes = p.nblk * kv_blk_bytes;
  const int q_bytes = p.nblk * 16384;                // one query tile
  uint8_t* sQ = smem;                                // [2 tiles][nblk][128 rows x 128 B]
  uint8_t* sK = sQ + 2 * q_bytes;                    // NS stages
  uint8_t* sV = sK + NS * kv_stage_bytes;            // NS stages
  uint8_t* sP = sV + NS * kv_stage_bytes;            // [2 tiles][128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * 16384);
  uint64_t* q_full = bars;                           // both query tiles landed
  uint64_t* k_full = bars + 1;                       // [NS]
  uint64_t* k_empty = k_full + NS;                   // [NS]  both tiles' QK^T of this key tile retired
  uint64_t* v_full = k_empty + NS;                   // [NS]
  uint64_t* v_empty = v_full + NS;                   // [NS]  both tiles' PV of this key tile retired
  uint64_t* s_full = v_empty + NS;                   // [2 tiles][2 buffers]
  uint64_t* s_empty = s_full + 4;                    // [2][2]
  uint64_t* p_full = s_empty + 4;                    // [2]
  uint64_t* p_empty = p_full + 2;                    // [2]
  uint64_t* o_full = p_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_blocks = (p.Nq + 2 * BM - 1) / (2 * BM);
  const int split = blockIdx.x / q_blocks;
  const int q0 = (blockIdx.x - split * q_blocks) * 2 * BM, h = blockIdx.y, b = blockIdx.z;   // first row of tile a; tile b starts at q0 + 128
  const int off = p.Nk - p.Nq;
  int k_begin = 0, k_end = p.Nk;
  if (p.causal) k_end = min(p.Nk, q0 + 2 * BM + off);   // tile b reaches furthest; tile a masks the surplus
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
    for (int s = 0; s < NS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&p_full[t], 128);
      mbar_init(&p_empty[t], 1);
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
  // tensor-map coordinates of a tile whose first row is r0: linear (row, batch) or windowed (x, y)
  const int wx0 = p.win ? (b % p.nwx) * p.win : 0, wy0 = p.win ? (b / p.nwx) * p.win : 0;
  auto c2 = [&](int r0) { return p.win ? wx0 : r0; };
  auto c3 = [&](int r0) { return p.win ? wy0 + r0 / p.win : b; };

  if (warp == 0) {
    if (lane == 0 && ntiles > 0) {
      // ------------------------------------------------ TMA: both query tiles, then the K ring
      mbar_arrive_expect_tx(q_full, 2 * q_bytes);
      for (int t = 0; t < 2; ++t)
        for (int blk = 0; blk < p.nblk; ++blk) tma_load_4d(&tmQ, q_full, sQ + t * q_bytes + blk * 16384, blk * 64, h, c2(q0 + t * BM), c3(q0 + t * BM));
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % NS;
        mbar_wait(&k_empty[st], ((j / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[st], kv_stage_bytes);
        const int k0 = k_begin + j * BN;
        for (int blk = 0; blk < p.nblk; ++blk)
          tma_load_4d(&tmK, &k_full[st], sK + st * kv_stage_bytes + blk * kv_blk_bytes, blk * 64, h, c2(k0), c3(k0));
      }
    }
  } else if (warp == 3) {
    if (lane == 0 && ntiles > 0) {
      // ------------------------------------------------ TMA: V ring
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % NS;
        mbar_wait(&v_empty[st], ((j / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[st], kv_stage_bytes);
        const int k0 = k_begin + j * BN;
        for (int blk = 0; blk < p.nblk; ++blk)
          tma_load_4d(&tmV, &v_full[st], sV + st * kv_stage_bytes + blk * kv_blk_bytes, blk * 64, h, c2(k0), c3(k0));
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && ntiles > 0) {
      // ------------------------------------------------ MMA issuer
      const uint32_t idesc_s = umma_idesc_bf16(BM, BN);
      const uint32_t idesc_o = umma_idesc_bf16_bmn(BM, p.Dp);
      mbar_wait(q_full, 0);
      auto issue_qk = [&](int j) {   // S_t[j & 1] = Q_t K_j^T for both tiles
        const int st = j % NS, sb = j & 1;
        mbar_wait(&k_full[st], (j / NS) & 1);
        for (int t = 0; t < 2; ++t) {
          if (j >= 2) mbar_wait(&s_empty[t * 2 + sb], ((j >> 1) - 1) & 1);
          tc_fence_after();
          const uint32_t tS = tmem_base + 256 + t * 128 + sb * BN;
          for (int blk = 0; blk < p.nblk; ++blk) {
            const int steps = min(4, (p.Dp - 64 * blk) >> 4);
            const uint64_t qd = umma_desc_sw128(smem_u32(sQ + t * q_bytes + blk * 16384));
            const uint64_t kd = umma_desc_sw128(smem_u32(sK + st * kv_stage_bytes + blk * kv_blk_bytes));
            for (int k = 0; k < steps; ++k) umma_bf16(tS, qd + 2 * k, kd + 2 * k, idesc_s, (blk | k) != 0);
          }
          umma_commit(&s_full[t * 2 + sb]);
        }
        umma_commit(&k_empty[st]);
      };
      issue_qk(0);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % NS;
        if (j + 1 < ntiles) issue_qk(j + 1);
        mbar_wait(&v_full[st], (j / NS) & 1);
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t], j & 1);
          tc_fence_after();
          const uint32_t tO = tmem_base + t * 128;
          for (int kk = 0; kk < BN / 16; ++kk) {
            const uint64_t pd = umma_desc_sw128(smem_u32(sP + t * 16384)) + 2 * kk;
            const uint64_t vd = umma_desc_sw128_mn(smem_u32(sV + st * kv_stage_bytes + kk * 2048), kv_blk_bytes, 1024);
            umma_bf16(tO, pd, vd, idesc_o, (j | kk) != 0);
          }
          umma_commit(&p_empty[t]);
        }
        umma_commit(&v_empty[st]);
      }
      umma_commit(o_full);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ softmax / correction / epilogue: warpgroup t owns query tile t, one row per thread
    const int t = (warp - 4) >> 2;
    const int ew = warp & 3;
    const int r = ew * 32 + lane;                  // row in the tile == TMEM lane
    const int row = q0 + t * BM + r;
    const uint32_t lane_addr = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t tO = tmem_base + t * 128;
    const float sl2 = p.scale * 1.4426950408889634f;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[t * 2 + sb], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t tS = tmem_base + 256 + t * 128 + sb * BN;
      const int k0 = k_begin + j * BN;
      const int col_lim = p.causal ? min(p.Nk, row + off + 1) : p.Nk;  // valid columns are < col_lim (<= 0: the whole row is masked in this tile)
      const bool need_mask = __any_sync(0xffffffffu, k0 + BN > col_lim);
      float sv[BN];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t raw[32];
        tmem_ld32(tS + lane_addr + c * 32, raw);
        tmem_ld_wait();
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = (k0 + c * 32 + i < col_lim) ? __uint_as_float(raw[i]) : -INFINITY;
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
      mbar_arrive(&s_empty[t * 2 + sb]);
      const float m_new = fmaxf(m_ref, mx);
      const bool need = (m_ref != -INFINITY) && ((m_new - m_ref) * sl2 > 8.0f);
      if (m_ref == -INFINITY) m_ref = m_new;
      bool pv_done = (j == 0);
      if (__any_sync(0xffffffffu, need)) {
        if (!pv_done) {
          mbar_wait(&p_empty[t], (j - 1) & 1);
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
      uint32_t pk[BN / 2];
#pragma unroll
      for (int i = 0; i < BN / 2; ++i) {
        const float p0 = ex2_approx(fmaf(sv[2 * i], sl2, -msc)), p1 = ex2_approx(fmaf(sv[2 * i + 1], sl2, -msc));
        l += p0 + p1;
        pk[i] = pack_bf16(p0, p1);
      }
      if (!pv_done) {
        mbar_wait(&p_empty[t], (j - 1) & 1);
        tc_fence_after();
      }
      uint8_t* prow = sP + t * 16384 + r * 128;
      const int sw = r & 7;
#pragma unroll
      for (int c8 = 0; c8 < BN / 8; ++c8)
        *reinterpret_cast<uint4*>(prow + (((c8 & 7) ^ sw) << 4)) = make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
      fence_async_shared();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
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
        const long long orow = p.win ? static_cast<long long>(wy0 + row / p.win) * p.gridw + wx0 + row % p.win : row;  // row is relative to the batch entry
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
    tmem_dealloc(tmem_base, 512);
  }
}

// Host side for round 2: same tensor maps as attention_tc.cu (make_map4: Q box 128 rows, K/V box 64 rows), grid
// ((Nq + 255) / 256 * nsplit, H, B), TC2_THREADS threads, dynamic shared memory
//   2 * nblk * 16384 + 2 * TC2_STAGES * nblk * 64 * 128 + 2 * 16384 + 256   (<= 227 KB for D <= 128),
// eligibility: D <= 128, Nq >= 256 (for Nq in (128, 256] the second tile is zero-filled by TMA and simply discarded), no sliding window.

}  // namespace vg
