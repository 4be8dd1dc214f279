// Persistent warp-specialised bf16 GEMM for sm_100a:
//   warp 0 (1 lane)  TMA producer      global --cp.async.bulk.tensor (SWIZZLE_128B)--> smem ring
//   warp 1 (1 lane)  MMA issuer        tcgen05.mma.cta_group::1.kind::f16, fp32 accumulators in TMEM (double-buffered)
//   warp 2           TMEM allocator
//   warps 4..11      epilogue          tcgen05.ld -> bias/act/GLU/scale/residual -> global (lane quadrant x column half)
// C[M,N] = A[M,K] * B[N,K]^T, both operands K-major (nn.Linear layout). See include/vglamm.h for the contract.
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.h"
#include "pdl.cuh"
#include "ptx.cuh"

namespace vg {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;  // 4 control warps (TMA, MMA, TMEM alloc, spare) + 8 epilogue warps

// WIDE: 16 epilogue warps instead of 8 (4 per scheduler; 112 registers each through setmaxnreg), one staging box per warp (64 KB), paid for with
// one pipeline stage where shared memory is short. For the small-K GEMMs, whose epilogue (tcgen05.ld -> bias -> activation -> pack -> st.shared ->
// TMA store) is a latency/ILP-bound instruction stream at two warps per scheduler (profiles/r01_gemm_epilogue.md).
template <int BN, bool WIDE = false>
struct GemmCfg {
  static constexpr int EPI_WARPS = WIDE ? 16 : 8;
  static constexpr int THREADS = (4 + EPI_WARPS) * 32;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ACC_STRIDE = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);  // TMEM columns between the two accumulator stages
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;  // two accumulator stages; allocations are powers of two (128/256/512)
  static constexpr int STG_PER_WARP = 4096;  // per epilogue warp: one 32-row TMA-store box (<= 128 B per row)
  static constexpr int STAGING_BYTES = EPI_WARPS * STG_PER_WARP;
  static constexpr int VEC_BYTES = 2 /*acc stages*/ * BN * 4;  // bias of the tile being drained
  static constexpr int STAGES_WANT = BN == 256 ? 4 : (BN == 192 ? 4 : (BN >= 96 ? 6 : 8));
  static constexpr int STAGES_FIT = (227 * 1024 - STAGING_BYTES - VEC_BYTES - 256) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_WANT < STAGES_FIT ? STAGES_WANT : STAGES_FIT;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + VEC_BYTES + 256 /*barriers*/;  // base is 1024-aligned (checked)
  static_assert(STAGES >= 3, "pipeline too shallow");
};

struct GemmParams {
  int M, N, K, batch;
  int tiles_m, tiles_n;
  int n_fastest;  // tile order: N tiles of one M stripe adjacent (A is the big operand) or M tiles of one N stripe adjacent (W is)
  void* C;
  long long ldc, strideC;
  int out_f32;
  const float* bias;
  long long strideBias;
  int act, glu;
  const float* col_scale;
  const void* resid;
  long long ldr, strideR;
  int resid_f32;
  float alpha;
  int b_shared;  // B (weights) shared by every batch entry
  int use_tma_store;  // C leaves through per-warp smem boxes + bulk tensor stores
  int tma_reduce;     // residual aliases C: TMA reduce-add instead of load + add + store
  int backoff_ns;     // > 0: the single-lane control warps sleep between mbarrier polls (their spin loops compete with the epilogue warps for issue slots)
};

// Activations for the epilogues. ncu (profiles/r01_ncu_targets_d.md) showed the small-K GEMMs issue-bound in the epilogue warps at ~35
// instructions per output element, ~25 of them the activation (libdevice-style __fdividef/__expf range handling). These forms are branch-free
// and use the raw MUFU approximations (rcp.approx / ex2.approx, ~1-2 ulp): GELU = 11 ALU + 2 MUFU, SiLU/QuickGELU/sigmoid = 3 ALU + 2 MUFU.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exact-erf GELU x*Phi(x) with erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7): for z = |x|/sqrt(2), t = 1/(1 + p z),
// Phi(|x|) = 1 - 0.5 t (a1 + t(a2 + t(a3 + t(a4 + t a5)))) exp(-x^2/2) =: 1 - h, and x*Phi(x) = max(x,0) - |x| h for either sign.
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(ax, 0.2316418882f, 1.0f));
  float q = fmaf(t, 0.5307027145f, -0.7265760135f);
  q = fmaf(q, t, 0.7107068705f);
  q = fmaf(q, t, -0.142248368f);
  q = fmaf(q, t, 0.127414796f);
  const float e = ex2_approx((x * -0.7213475204f) * x);
  const float h = q * (t * e);
  return fmaf(-ax, h, fmaxf(x, 0.0f));
}
__device__ __forceinline__ float sigmoid_scaled(float x, float neg_scale_log2e) {  // 1 / (1 + exp(-scale * x))
  return rcp_approx(1.0f + ex2_approx(x * neg_scale_log2e));
}

template <int N>
__device__ __forceinline__ void apply_act_vec(float (&v)[32], int act) {
  // one switch per 32-column chunk (a per-element switch compiled to ~300 cycles/element and made every activated GEMM 4x slower)
  switch (act) {
    case VG_ACT_GELU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = gelu_erf(v[j]);
      break;
    case VG_ACT_QUICKGELU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] *= sigmoid_scaled(v[j], -2.4554669596f);
      break;
    case VG_ACT_RELU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = fmaxf(v[j], 0.0f);
      break;
    case VG_ACT_SILU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] *= sigmoid_scaled(v[j], -1.4426950409f);
      break;
    case VG_ACT_SIGMOID:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = sigmoid_scaled(v[j], -1.4426950409f);
      break;
    default: break;
  }
}

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case VG_ACT_GELU: return gelu_erf(x);
    case VG_ACT_QUICKGELU: return x * sigmoid_scaled(x, -2.4554669596f);
    case VG_ACT_RELU: return fmaxf(x, 0.0f);
    case VG_ACT_SILU: return x * sigmoid_scaled(x, -1.4426950409f);
    case VG_ACT_SIGMOID: return sigmoid_scaled(x, -1.4426950409f);
    default: return x;
  }
}

template <int BN, bool WIDE>
__global__ void __launch_bounds__(GemmCfg<BN, WIDE>::THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                    const GemmParams p) {
  using Cfg = GemmCfg<BN, WIDE>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023) != 0) __trap();  // SWIZZLE_128B tiles need a 1024-byte aligned base
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  float* s_vec = reinterpret_cast<float*>(staging + Cfg::STAGING_BYTES);  // [acc][BN] bias
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES + Cfg::VEC_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.use_tma_store) tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], Cfg::EPI_WARPS * 32);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the predecessor's tail;
  // operands, residual and output belong to it from here on.
  pdl_launch_dependents();
  pdl_wait();

  const int num_k = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int tiles_per_batch = p.tiles_m * p.tiles_n;
  const int total_tiles = tiles_per_batch * p.batch;

  // Register rebalancing (the three control warps + spare form warpgroup 0 and need few registers; the two epilogue warpgroups hold a
  // 32-column accumulator chunk plus a prefetched residual chunk): 128*56 + 256*224 = 64512 <= 65536.
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = t / tiles_per_batch;
        const int r = t - b * tiles_per_batch;
        const int n_blk = p.n_fastest ? r % p.tiles_n : r / p.tiles_m;
        const int m_blk = p.n_fastest ? r / p.tiles_n : r - n_blk * p.tiles_m;
        for (int kb = 0; kb < num_k; ++kb) {
          while (!mbar_try_wait(&empty_bar[stage], phase ^ 1)) {
            if (p.backoff_ns) __nanosleep(p.backoff_ns);
          }
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_3d(&tmA, &full_bar[stage], sA + stage * Cfg::A_BYTES, kb * BLOCK_K, m_blk * BLOCK_M, b);
          tma_load_3d(&tmB, &full_bar[stage], sB + stage * Cfg::B_BYTES, kb * BLOCK_K, n_blk * BN, p.b_shared ? 0 : b);
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ------------------------------------------------ MMA issuer: the whole warp runs the loop converged and one elected lane issues
      // (umma_bf16_elect). Under `if (lane == 0)` the compiler wraps every tcgen05 instruction in an ELECT / BRA.U.ANY loop, ~50 cycles of the
      // single issuing thread each: more than the 32-64 cycles a BLOCK_N <= 128 MMA keeps the tensor pipe busy.
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        while (!mbar_try_wait(&tempty_bar[acc], acc_phase ^ 1)) {
          if (p.backoff_ns) __nanosleep(p.backoff_ns);
        }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::ACC_STRIDE;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_bf16_elect(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit_elect(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_elect(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  }
  } else {
    if (WIDE)
      asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");  // 128*56 + 512*104 = 60416 <= 640*96 (the block's own allocation; 112 deadlocked)
    else
      asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ------------------------------------------------ epilogue: TMEM -> regs -> (smem -> TMA store | global) ; 8 warps
    // Each warp owns a TMEM lane quadrant (warp % 4) and every other 32-column chunk. Results leave through a per-warp swizzled
    // shared-memory box and ONE bulk tensor store (or reduce-add, when the residual aliases the output) issued by lane 0:
    // thread-per-row st.global touched 32 lines per instruction and capped the epilogue at ~18 GB/s per SM (profiles/r01_gemm_epilogue.md).
    const int ew = warp & 3;
    const int half = (warp - 4) >> 2;                 // column group of this warp: chunks half, half + CG, ...
    constexpr int CG = Cfg::EPI_WARPS / 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int n_out = p.glu ? p.N / 2 : p.N;
    const bool unit_alpha = p.alpha == 1.0f;
    uint8_t* stg_base = staging + (warp - 4) * Cfg::STG_PER_WARP;
    const int esz = p.out_f32 ? 4 : 2;
    const int nvals_c = p.glu ? 16 : 32;
    const int row_bytes = nvals_c * esz;                         // 128 / 64 / 32: equals the TMA swizzle span
    const int sw_mask = (row_bytes >> 4) - 1;
    const int sw = (lane >> (row_bytes == 128 ? 0 : (row_bytes == 64 ? 1 : 2))) & sw_mask;  // XOR applied to the 16-byte chunk index
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int b = t / tiles_per_batch;
      const int r = t - b * tiles_per_batch;
      const int n_blk = p.n_fastest ? r % p.tiles_n : r / p.tiles_m;
      const int m_blk = p.n_fastest ? r / p.tiles_n : r - n_blk * p.tiles_m;
      // stage the tile's bias in shared memory BEFORE waiting for the accumulator (latency overlaps the MMA main loop)
      float* s_bias = s_vec + acc * BN;
      if (p.bias != nullptr) {
        const int et = threadIdx.x - 128;
        for (int j = et; j < BN; j += Cfg::EPI_WARPS * 32) {
          const int col = n_blk * BN + j;
          s_bias[j] = col < p.N ? __ldg(p.bias + b * p.strideBias + col) : 0.f;
        }
        if (WIDE)
          asm volatile("bar.sync 1, 512;" ::: "memory");  // epilogue warps only
        else
          asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BLOCK_M + ew * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * Cfg::ACC_STRIDE;
#pragma unroll 1
      for (int c = half; c < BN / 32; c += CG) {
        const int col0 = n_blk * BN + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t raw[32];
        tmem_ld32(taddr + c * 32, raw);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        if (!unit_alpha) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
        }
        if (p.bias != nullptr) {
          const float4* bp = reinterpret_cast<const float4*>(s_bias + c * 32);  // broadcast shared loads
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = bp[j];
            v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
          }
        }
        int ocol0, nvals;
        if (p.glu) {
          apply_act_vec<16>(v, p.act);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= v[16 + j];
          ocol0 = col0 >> 1;
          nvals = 16;
        } else {
          if (p.act != VG_ACT_NONE) apply_act_vec<32>(v, p.act);
          ocol0 = col0;
          nvals = 32;
        }
        const int nvalid = min(nvals, n_out - ocol0);
        if (p.col_scale != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) v[j] *= __ldg(p.col_scale + ocol0 + j);
        }
        // residual that does NOT alias the output: thread-per-row loads (rare; aliasing residuals use the TMA reduce-add below)
        if (p.resid != nullptr && !p.tma_reduce && row_ok) {
          const long long roff = b * p.strideR + static_cast<long long>(row) * p.ldr + ocol0;
          if (p.resid_f32) {
            const float* rp = reinterpret_cast<const float*>(p.resid) + roff;
            if (nvalid == nvals && (reinterpret_cast<uintptr_t>(rp) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (4 * j < nvals) {
                  const float4 q = *reinterpret_cast<const float4*>(rp + 4 * j);
                  v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) v[j] += rp[j];
            }
          } else {
            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.resid) + roff;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) v[j] += __bfloat162float(rp[j]);
          }
        }
        if (p.use_tma_store) {
          // the previous bulk store of this warp must have finished READING the staging box before it is overwritten
          uint8_t* stg = stg_base;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
          uint8_t* srow = stg + lane * row_bytes;
          if (p.out_f32) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (4 * j < nvals) *reinterpret_cast<float4*>(srow + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (8 * j < nvals)
                *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                                             pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            const int r0 = m_blk * BLOCK_M + ew * 32;
            if (p.tma_reduce) {
              asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                               reinterpret_cast<uint64_t>(&tmC)),
                           "r"(smem_u32(stg)), "r"(ocol0), "r"(r0), "r"(b)
                           : "memory");
            } else {
              asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tmC)),
                           "r"(smem_u32(stg)), "r"(ocol0), "r"(r0), "r"(b)
                           : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else if (row_ok) {
          // fallback (unaligned C / ldc): thread-per-row scalar stores, aliasing residual added here
          const long long coff = b * p.strideC + static_cast<long long>(row) * p.ldc + ocol0;
          if (p.resid != nullptr && p.tma_reduce) {
            const long long roff = b * p.strideR + static_cast<long long>(row) * p.ldr + ocol0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid)
                v[j] += p.resid_f32 ? reinterpret_cast<const float*>(p.resid)[roff + j] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.resid)[roff + j]);
          }
          if (p.out_f32) {
            float* cp = reinterpret_cast<float*>(p.C) + coff;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) cp[j] = v[j];
          } else {
            __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + coff;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) cp[j] = __float2bfloat16(v[j]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all bulk stores of this warp complete before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------- skinny GEMM (M <= 8): LLM decode
// HBM-bound weight streaming: one warp per output column (two weight rows in GLU mode), 128-bit non-allocating loads of the
// weight row, A rows re-read through L1, fp32 accumulate, warp-shuffle reduction, same epilogue contract as the tcgen05 path.
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float dot8(const uint4& w, const uint4& a) {
  float s = bf16_lo(w.x) * bf16_lo(a.x);
  s = fmaf(bf16_hi(w.x), bf16_hi(a.x), s);
  s = fmaf(bf16_lo(w.y), bf16_lo(a.y), s);
  s = fmaf(bf16_hi(w.y), bf16_hi(a.y), s);
  s = fmaf(bf16_lo(w.z), bf16_lo(a.z), s);
  s = fmaf(bf16_hi(w.z), bf16_hi(a.z), s);
  s = fmaf(bf16_lo(w.w), bf16_lo(a.w), s);
  s = fmaf(bf16_hi(w.w), bf16_hi(a.w), s);
  return s;
}

template <int MR, int WPC, int U, bool AF32>
// MR: rows of A; WPC: warps cooperating on one output column (split along K); U: independent 16-byte weight loads in flight per lane;
// AF32: A is the fp32 residual stream and the row's RMS normalisation (x * rsqrt(mean(x^2) + eps), norm weight folded into B at load
// time) is fused: the sum of squares is accumulated from the same A loads the dot product needs.
__global__ void __launch_bounds__(256) gemv_kernel(const void* __restrict__ Av, long long lda, const __nv_bfloat16* __restrict__ B, long long ldb,
                                                   const GemmParams p, float rms_eps) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  __shared__ float s_part[8][3 * MR];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_out = p.glu ? p.N / 2 : p.N;
  const int oc = (blockIdx.x * 8 + wib) / WPC;
  const int ks = wib % WPC;  // K slice of this warp
  const bool active = oc < n_out;
  const int n0 = p.glu ? (oc / 16) * 32 + (oc % 16) : oc;
  const __nv_bfloat16* w0 = B + static_cast<long long>(active ? n0 : 0) * ldb;
  const __nv_bfloat16* w1 = w0 + 16 * ldb;  // "up" row (GLU only)
  float acc0[MR], acc1[MR], ssq[MR];
#pragma unroll
  for (int m = 0; m < MR; ++m) acc0[m] = acc1[m] = ssq[m] = 0.f;
  const int chunks = p.K >> 3;
  if (active) {
    for (int c0 = ks * 32 + lane; c0 < chunks; c0 += 32 * WPC * U) {
      uint4 wv0[U], wv1[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + u * 32 * WPC;
        wv0[u] = make_uint4(0, 0, 0, 0);
        wv1[u] = make_uint4(0, 0, 0, 0);
        if (c < chunks) {
          wv0[u] = ld_stream16(w0 + c * 8);
          if (p.glu) wv1[u] = ld_stream16(w1 + c * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + u * 32 * WPC;
        if (c < chunks) {
#pragma unroll
          for (int m = 0; m < MR; ++m) {
            if (m < p.M) {
              if (AF32) {
                const float4* ap = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(Av) + m * lda + c * 8);
                const float4 a0 = __ldg(ap), a1 = __ldg(ap + 1);
                ssq[m] += a0.x * a0.x + a0.y * a0.y + a0.z * a0.z + a0.w * a0.w + a1.x * a1.x + a1.y * a1.y + a1.z * a1.z + a1.w * a1.w;
                const uint4 w = wv0[u];
                acc0[m] += bf16_lo(w.x) * a0.x + bf16_hi(w.x) * a0.y + bf16_lo(w.y) * a0.z + bf16_hi(w.y) * a0.w + bf16_lo(w.z) * a1.x +
                           bf16_hi(w.z) * a1.y + bf16_lo(w.w) * a1.z + bf16_hi(w.w) * a1.w;
                if (p.glu) {
                  const uint4 g = wv1[u];
                  acc1[m] += bf16_lo(g.x) * a0.x + bf16_hi(g.x) * a0.y + bf16_lo(g.y) * a0.z + bf16_hi(g.y) * a0.w + bf16_lo(g.z) * a1.x +
                             bf16_hi(g.z) * a1.y + bf16_lo(g.w) * a1.z + bf16_hi(g.w) * a1.w;
                }
              } else {
                const uint4 av = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(Av) + m * lda + c * 8));
                acc0[m] += dot8(wv0[u], av);
                if (p.glu) acc1[m] += dot8(wv1[u], av);
              }
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MR; ++m) {
    acc0[m] = warp_sum(acc0[m]);
    if (p.glu) acc1[m] = warp_sum(acc1[m]);
    if (AF32) ssq[m] = warp_sum(ssq[m]);
  }
  if (WPC > 1) {
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        s_part[wib][3 * m] = acc0[m];
        s_part[wib][3 * m + 1] = acc1[m];
        s_part[wib][3 * m + 2] = ssq[m];
      }
    }
    __syncthreads();
    if (ks != 0) return;
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      acc0[m] = acc1[m] = ssq[m] = 0.f;
#pragma unroll
      for (int j = 0; j < WPC; ++j) {
        acc0[m] += s_part[wib + j][3 * m];
        acc1[m] += s_part[wib + j][3 * m + 1];
        ssq[m] += s_part[wib + j][3 * m + 2];
      }
    }
  }
  if (lane != 0 || !active) return;
#pragma unroll
  for (int m = 0; m < MR; ++m) {
    if (m >= p.M) break;
    const float nrm = AF32 ? rsqrtf(ssq[m] / p.K + rms_eps) * p.alpha : p.alpha;
    float v = acc0[m] * nrm;
    if (p.bias) v += p.bias[n0];
    if (p.glu) {
      float u = acc1[m] * nrm;
      if (p.bias) u += p.bias[n0 + 16];
      v = apply_act(v, p.act) * u;
    } else {
      v = apply_act(v, p.act);
    }
    if (p.col_scale) v *= p.col_scale[oc];
    if (p.resid) {
      const long long off = static_cast<long long>(m) * p.ldr + oc;
      v += p.resid_f32 ? reinterpret_cast<const float*>(p.resid)[off] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.resid)[off]);
    }
    const long long off = static_cast<long long>(m) * p.ldc + oc;
    if (p.out_f32)
      reinterpret_cast<float*>(p.C)[off] = v;
    else
      reinterpret_cast<__nv_bfloat16*>(p.C)[off] = __float2bfloat16(v);
  }
}

static int gemv_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VG_GEMV_U");
    v = e ? atoi(e) : 4;
  }
  return v;
}

template <int MR, bool AF32>
static void launch_gemv(const VgGemm* g, const GemmParams& p, cudaStream_t s) {
  const int n_out = g->glu ? g->N / 2 : g->N;
  const __nv_bfloat16* B = reinterpret_cast<const __nv_bfloat16*>(g->B);
  // few columns and long rows: split K over 4 warps so enough bytes are in flight to cover HBM latency
  const bool split = n_out < 6144 && g->K >= 2048;
  const unsigned blocks = split ? (static_cast<unsigned>(n_out) * 4 + 7) / 8 : (static_cast<unsigned>(n_out) + 7) / 8;
  const int u = MR > 4 ? 2 : gemv_variant();   // 16-byte weight loads in flight per lane (registers: 4 U, x2 with GLU, next to 3 MR accumulators)
#define VG_GEMV_LAUNCH(WPC, UU) launch_k(gemv_kernel<MR, WPC, UU, AF32>, dim3(blocks), dim3(256), 0, s, g->A, g->lda, B, g->ldb, p, g->a_rms_eps)
  if (split) {
    if (u >= 8) VG_GEMV_LAUNCH(4, 8); else if (u >= 4) VG_GEMV_LAUNCH(4, 4); else VG_GEMV_LAUNCH(4, 2);
  } else {
    if (u >= 8) VG_GEMV_LAUNCH(1, 8); else if (u >= 4) VG_GEMV_LAUNCH(1, 4); else VG_GEMV_LAUNCH(1, 2);
  }
#undef VG_GEMV_LAUNCH
}

// ---------------------------------------------------------------------------------------------- naive reference
__global__ void gemm_naive_kernel(const __nv_bfloat16* A, long long lda, long long sA, const __nv_bfloat16* B, long long ldb,
                                  long long sB, GemmParams p) {
  pdl_wait();  // programmatic dependent launch: inputs belong to the predecessor kernel
  pdl_launch_dependents();
  const int n_out = p.glu ? p.N / 2 : p.N;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long per_batch = static_cast<long long>(p.M) * n_out;
  if (idx >= per_batch * p.batch) return;
  const int b = static_cast<int>(idx / per_batch);
  const long long r = idx - b * per_batch;
  const int m = static_cast<int>(r / n_out);
  const int oc = static_cast<int>(r - static_cast<long long>(m) * n_out);
  auto dot = [&](int n) {
    const __nv_bfloat16* a = A + b * sA + static_cast<long long>(m) * lda;
    const __nv_bfloat16* w = B + b * sB + static_cast<long long>(n) * ldb;
    float acc = 0.f;
    for (int k = 0; k < p.K; ++k) acc = fmaf(__bfloat162float(a[k]), __bfloat162float(w[k]), acc);
    acc *= p.alpha;
    if (p.bias) acc += p.bias[b * p.strideBias + n];
    return acc;
  };
  float v;
  if (p.glu) {
    const int blk = oc / 16, j = oc % 16;
    v = apply_act(dot(blk * 32 + j), p.act) * dot(blk * 32 + 16 + j);
  } else {
    v = apply_act(dot(oc), p.act);
  }
  if (p.col_scale) v *= p.col_scale[oc];
  if (p.resid) {
    const long long off = b * p.strideR + static_cast<long long>(m) * p.ldr + oc;
    v += p.resid_f32 ? reinterpret_cast<const float*>(p.resid)[off] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.resid)[off]);
  }
  const long long off = b * p.strideC + static_cast<long long>(m) * p.ldc + oc;
  if (p.out_f32)
    reinterpret_cast<float*>(p.C)[off] = v;
  else
    reinterpret_cast<__nv_bfloat16*>(p.C)[off] = __float2bfloat16(v);
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  long long ld, stride;
  int rows, cols, batch, box_rows;
  int kind = 0;  // 0: bf16 operand (box 64 x box_rows, SWIZZLE_128B); else (elem_bytes << 8 | box_cols) for output boxes
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && ld == o.ld && stride == o.stride && rows == o.rows && cols == o.cols && batch == o.batch &&
           box_rows == o.box_rows && kind == o.kind;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&](size_t v) { h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
    mix(static_cast<size_t>(k.ld));
    mix(static_cast<size_t>(k.stride));
    mix(static_cast<size_t>(k.rows));
    mix(static_cast<size_t>(k.cols));
    mix(static_cast<size_t>(k.batch));
    mix(static_cast<size_t>(k.box_rows));
    mix(static_cast<size_t>(k.kind));
    return h;
  }
};

// bf16 [batch][rows][ld] (cols valid per row) -> 3D tiled map with box {64, box_rows, 1}, 128B swizzle, zero OOB fill.
static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

// Output map: [batch][rows][ld] of `esz`-byte elements, box {box_cols, 32, 1}; the shared-memory side is swizzled with a span equal to
// the box row (128/64/32 bytes) so that thread-per-row 16-byte writes into the box are bank-conflict-free.
static int make_out_map(CUtensorMap* out, const void* ptr, int rows, int cols, int batch, long long ld, long long stride, int esz, int box_cols) {
  MapKey key{ptr, ld, stride, rows, cols, batch, 32, (esz << 8) | box_cols};
  {
    std::lock_guard<std::mutex> g(g_map_mu);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
      *out = it->second;
      return VG_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(VG_ECUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batch)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(ld) * esz, static_cast<cuuint64_t>(batch > 1 ? stride : static_cast<long long>(rows) * ld) * esz};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const int row_bytes = box_cols * esz;
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(out, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(VG_ECUDA, "cuTensorMapEncodeTiled (output) failed (%d): rows=%d cols=%d ld=%lld", (int)r, rows, cols, ld);
  {
    std::lock_guard<std::mutex> g(g_map_mu);
    if (g_map_cache.size() > 65536) g_map_cache.clear();
    g_map_cache.emplace(key, *out);
  }
  return VG_OK;
}

static int make_map(CUtensorMap* out, const void* ptr, int rows, int cols, int batch, long long ld, long long stride, int box_rows) {
  std::mutex& mu = g_map_mu;
  auto& cache = g_map_cache;
  MapKey key{ptr, ld, stride, rows, cols, batch, box_rows, 0};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return VG_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(VG_ECUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batch)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(batch > 1 ? stride : static_cast<long long>(rows) * ld) * 2};
  cuuint32_t box[3] = {BLOCK_K, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(VG_ECUDA, "cuTensorMapEncodeTiled failed (%d): rows=%d cols=%d ld=%lld", (int)r, rows, cols, ld);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, *out);
  }
  return VG_OK;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int check_gemm(const VgGemm* g) {
  if (!g || !g->A || !g->B || !g->C) return set_error(VG_EINVAL, "vg_gemm: null pointer");
  if (g->M <= 0 || g->N <= 0 || g->K <= 0 || g->batch <= 0) return set_error(VG_EINVAL, "vg_gemm: non-positive dimension");
  if (g->glu && (g->N % 32) != 0) return set_error(VG_EINVAL, "vg_gemm: glu needs N %% 32 == 0");
  if (g->resid && g->ldr <= 0) return set_error(VG_EINVAL, "vg_gemm: residual needs ldr");
  return VG_OK;
}

static GemmParams to_params(const VgGemm* g) {
  GemmParams p;
  p.M = g->M; p.N = g->N; p.K = g->K; p.batch = g->batch;
  p.tiles_m = 0; p.tiles_n = 0; p.n_fastest = 0; p.backoff_ns = 0;
  p.C = g->C; p.ldc = g->ldc; p.strideC = g->strideC;
  p.out_f32 = g->out_dtype == VG_F32;
  p.bias = g->bias; p.strideBias = g->strideBias; p.act = g->act; p.glu = g->glu; p.col_scale = g->col_scale;
  p.resid = g->resid; p.ldr = g->ldr; p.strideR = g->strideR; p.resid_f32 = g->resid_dtype == VG_F32;
  p.alpha = g->alpha;
  p.b_shared = (g->batch > 1 && g->strideB == 0) ? 1 : 0;
  p.use_tma_store = 0;
  p.tma_reduce = 0;
  return p;
}

template <int BN, bool WIDE>
static int launch_gemm_cfg(const VgGemm* g, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, WIDE>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(VG_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, g->A, g->M, g->K, g->batch, g->lda, g->strideA, BLOCK_M);
  if (rc) return rc;
  rc = make_map(&tmB, g->B, g->N, g->K, (g->batch > 1 && g->strideB == 0) ? 1 : g->batch, g->ldb, g->strideB, BN);
  if (rc) return rc;
  GemmParams p = to_params(g);
  p.tiles_m = (g->M + BLOCK_M - 1) / BLOCK_M;
  p.tiles_n = (g->N + BN - 1) / BN;
  // Tile order: the operand that does NOT fit the concurrent working set should be swept once. A (M x K) bigger than W (N x K): the N
  // tiles of one M stripe are adjacent, so A is read from HBM once and W stays in L2 (ncu: fc2 32768x576x2304 read 501 MB instead of 226).
  p.n_fastest = (static_cast<long long>(g->M) * g->batch > g->N && p.tiles_n > 1) ? 1 : 0;
  static const int backoff = [] { const char* e = getenv("VG_GEMM_BACKOFF"); return e ? atoi(e) : 0; }();
  p.backoff_ns = backoff;
  // Output path: bulk tensor stores need a 16-byte aligned base and 16-byte multiple strides. A residual that aliases the output
  // (the in-place residual-stream update every transformer block does) becomes a TMA reduce-add; any other residual is added in registers.
  const int esz = g->out_dtype == VG_F32 ? 4 : 2;
  const int n_out = g->glu ? g->N / 2 : g->N;
  const bool aliased = g->resid != nullptr && g->resid == g->C && g->ldr == g->ldc && (g->batch == 1 || g->strideR == g->strideC) &&
                       g->resid_dtype == g->out_dtype;
  p.tma_reduce = aliased ? 1 : 0;
  p.use_tma_store = ((reinterpret_cast<uintptr_t>(g->C) & 15) == 0) && ((g->ldc * esz) % 16 == 0) && (g->batch == 1 || (g->strideC * esz) % 16 == 0);
  CUtensorMap tmC = tmA;  // placeholder when the fallback path is taken
  if (p.use_tma_store) {
    rc = make_out_map(&tmC, g->C, g->M, n_out, g->batch, g->ldc, g->strideC, esz, g->glu ? 16 : 32);
    if (rc) return rc;
  }
  const long long total = static_cast<long long>(p.tiles_m) * p.tiles_n * g->batch;
  const int grid = static_cast<int>(total < num_sms() ? total : num_sms());
  launch_k(gemm_tcgen05_kernel<BN, WIDE>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, tmC, p);
  count_launch();
  return check_launch("gemm_tcgen05_kernel");
}

// Small-K problems are bound by the epilogue, large-K ones by the main loop: pick the epilogue width / pipeline depth accordingly.
// VG_GEMM_WIDE=0 forces the 8-warp epilogue everywhere (A/B timing).
template <int BN>
static int launch_gemm(const VgGemm* g, cudaStream_t stream) {
  // measured (profiles/r02_gemm_epilogue.md): 16 epilogue warps gain 10-15 % on the small-K GEMMs whose epilogue evaluates an activation
  // (Hiera fc1: 111 -> 96 us, stage-2 fc1: 199 -> 181 us) and nothing (-3 %) on bias-only epilogues -> default on for the former only
  static const int wide = [] { const char* e = getenv("VG_GEMM_WIDE"); return e ? atoi(e) : 2; }();  // 0 never, 1 every small-K GEMM, 2 (default) activated epilogues
  if (wide && g->K <= 1024 && static_cast<long long>(g->M) * g->N >= (1LL << 22) && (wide == 1 || g->act != VG_ACT_NONE))
    return launch_gemm_cfg<BN, true>(g, stream);
  return launch_gemm_cfg<BN, false>(g, stream);
}

}  // namespace vg

namespace vg {
int gemv_ring_try(const VgGemm* g, cudaStream_t s);
int gemv_ring_mr_try(const VgGemm* g, cudaStream_t s);  // gemv_ring.cu
}
using namespace vg;

extern "C" int vg_gemm(const VgGemm* g, void* stream) {
  int rc = check_gemm(g);
  if (rc) return rc;
  if (g->a_dtype == VG_F32 && !(g->M <= 8 && g->batch == 1 && g->force_block_n == 0))
    return set_error(VG_ENOSUP, "vg_gemm: fp32 A (fused RMSNorm) is only supported on the skinny path (M <= 8, batch 1)");
  if (((g->lda % 8) && g->a_dtype != VG_F32) || (g->ldb % 8) || (reinterpret_cast<uintptr_t>(g->A) & 15) || (reinterpret_cast<uintptr_t>(g->B) & 15) ||
      (g->batch > 1 && ((g->strideA % 8) || (g->strideB % 8))))
    return set_error(VG_EINVAL, "vg_gemm: A/B must be 16B aligned with lda/ldb/strides multiples of 8 (lda=%lld ldb=%lld)", g->lda, g->ldb);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (g->M <= 8 && g->batch == 1 && g->force_block_n == 0 && (g->K % 8) == 0) {
    GemmParams p = to_params(g);
    const bool af32 = g->a_dtype == VG_F32;
    if (af32 && ((g->lda % 4) || (reinterpret_cast<uintptr_t>(g->A) & 15))) return set_error(VG_EINVAL, "vg_gemm: fp32 A must be 16B aligned");
    if (g->M == 1) {
      const int rrc = gemv_ring_try(g, s);  // persistent TMA-ring kernel (decode-sized problems); VG_ENOSUP -> generic skinny kernel
      if (rrc != VG_ENOSUP) return rrc;
    } else {
      const int rrc = gemv_ring_mr_try(g, s);  // 2..8 rows (batched decode): the same streaming ring feeding mma.sync
      if (rrc != VG_ENOSUP) return rrc;
    }
    if (g->M == 1) { if (af32) launch_gemv<1, true>(g, p, s); else launch_gemv<1, false>(g, p, s); }
    else if (g->M <= 2) { if (af32) launch_gemv<2, true>(g, p, s); else launch_gemv<2, false>(g, p, s); }
    else if (g->M <= 4) { if (af32) launch_gemv<4, true>(g, p, s); else launch_gemv<4, false>(g, p, s); }
    else { if (af32) launch_gemv<8, true>(g, p, s); else launch_gemv<8, false>(g, p, s); }
    count_launch();
    return check_launch("gemv_kernel");
  }
  int bn = g->force_block_n;
  if (bn == 0) {
    const long long tm = (g->M + BLOCK_M - 1) / BLOCK_M;
    const long long t256 = tm * ((g->N + 255) / 256) * g->batch;
    const long long t128 = tm * ((g->N + 127) / 128) * g->batch;
    if (g->N > 128 && t256 >= num_sms())
      bn = 256;
    else if (g->N > 64 && t128 >= num_sms() / 2)
      bn = 128;
    else
      bn = (g->N > 64 && g->N % 128 == 0 && t128 * 2 > num_sms()) ? 128 : 64;
  }
  if (bn == 256 && g->force_block_n == 0) {
    // N not a multiple of 256 (Hiera widths are multiples of 144): 192- or 96-column tiles waste fewer MMA columns. Cost model: waves x
    // (tile columns + a fixed per-tile overhead of ~32 columns); ties keep the wider tile (fewer A re-reads from shared memory).
    // Measured exceptions (profiles/r02_gemm_bn_sweep.md): in the epilogue-bound small-K regime a 256-wide tile beats 192 whenever it wastes
    // <= 12 % of its columns and the problem is several waves deep (131072x1152x288 GELU: 165 vs 183 us; 32768x1728x576: 61 vs 66 us).
    const long long tm = (g->M + BLOCK_M - 1) / BLOCK_M;
    const long long n256 = (g->N + 255) / 256;
    const bool keep256 = g->K <= 576 && n256 * 256 * 100 <= static_cast<long long>(g->N) * 112 && tm * n256 * g->batch >= 4LL * num_sms();
    double best = 1e30;
    for (int cand : {256, 192, 96}) {
      if (keep256 && cand != 256) continue;
      if (cand == 96 && (g->N % 96 != 0 || g->N % 128 == 0)) continue;
      const long long tiles = tm * ((g->N + cand - 1) / cand) * g->batch;
      const double cost = static_cast<double>((tiles + num_sms() - 1) / num_sms()) * (cand + 32);
      if (cost < best) {   // strictly cheaper (131072x864x288: 192 at 7840 vs 256 at 8064 measures 101 vs 109 us)
        best = cost;
        bn = cand;
      }
    }
  }
  switch (bn) {
    case 256: return launch_gemm<256>(g, s);
    case 192: return launch_gemm<192>(g, s);
    case 128: return launch_gemm<128>(g, s);
    case 96: return launch_gemm<96>(g, s);
    case 64: return launch_gemm<64>(g, s);
    default: return set_error(VG_EINVAL, "vg_gemm: force_block_n must be 0/64/96/128/192/256");
  }
}

extern "C" int vg_gemm_naive(const VgGemm* g, void* stream) {
  int rc = check_gemm(g);
  if (rc) return rc;
  GemmParams p = to_params(g);
  const int n_out = g->glu ? g->N / 2 : g->N;
  const long long total = static_cast<long long>(g->M) * n_out * g->batch;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  gemm_naive_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(g->A), g->lda, g->strideA, reinterpret_cast<const __nv_bfloat16*>(g->B), g->ldb,
      g->strideB, p);
  count_launch();
  return check_launch("gemm_naive_kernel");
}
