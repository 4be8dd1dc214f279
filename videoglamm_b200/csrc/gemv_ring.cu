// Weight-streaming GEMV for LLM decode (M = 1) on sm_100a: y = epilogue(W x), W bf16 [N, K] row-major, HBM-bound.
//
// ncu on the warp-per-column kernel it replaces (profiles/r01_ncu_targets_d.md rows 4-5): 51 % / 33 % of DRAM peak when a launch stands
// alone — 8-25 us kernels pay launch ramp, a partial last wave and a drain. This version
//   * is persistent: min(N, 2 x SMs) blocks, each walking its rows through a shared-memory ring;
//   * moves weights with the TMA engine: one elected thread issues 1D bulk copies (cp.async.bulk, whole 6-16 KB rows) that complete on
//     mbarriers, so ~100 KB per block (~200 KB per SM) is in flight regardless of what the math warps are doing;
//   * is launched with programmatic dependent launch: the producer starts streaming weights (constants) as soon as the block is resident,
//     i.e. during the predecessor kernel's tail; only the activation vector waits for `griddepcontrol.wait`.
//   warps 0..7  consumers: one ring slot per warp and work item; 16-byte conflict-free shared loads, fp32 FMAs, warp-shuffle reduction,
//               lane 0 applies the epilogue (fused RMSNorm scale, bias, SwiGLU, column scale, residual) and stores one value
//   warp 8      producer (1 lane)
#include <cstdlib>

#include "common.h"
#include "pdl.cuh"
#include "ptx.cuh"

namespace vg {

struct GemvRingParams {
  const void* A;             // [K] fp32 (AF32) or bf16
  const __nv_bfloat16* B;    // [N, K]
  long long ldb;
  void* C;
  int N, K, n_out, glu, act, out_f32, resid_f32;
  const float* bias;
  const float* col_scale;
  const void* resid;
  float alpha, rms_eps;
  int depth, slot_bytes, seg_elems, a_bytes;  // per-warp ring depth; bytes per slot (one K-segment of one row, x2 for GLU)
};

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float act_ring(float x, int act) {
  switch (act) {
    case VG_ACT_RELU: return fmaxf(x, 0.0f);
    case VG_ACT_SILU: return x / (1.0f + __expf(-x));
    case VG_ACT_SIGMOID: return 1.0f / (1.0f + __expf(-x));
    case VG_ACT_QUICKGELU: return x / (1.0f + __expf(-1.702f * x));
    case VG_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    default: return x;
  }
}

constexpr int RING_CONSUMERS = 8;
constexpr int RING_THREADS = (RING_CONSUMERS + 1) * 32;
constexpr int RING_MAX_DEPTH = 4;

// Each consumer warp owns a private mini-ring of `depth` slots and consumes its items strictly in order (so an mbarrier is never waited on
// more than one phase ahead). An item is one K-segment of one output row (of the gate/up row pair for GLU); the segments of a row go to the
// same warp back to back, the dot product is carried in registers across them.
template <bool AF32, bool GLU>
__global__ void __launch_bounds__(RING_THREADS) gemv_ring_kernel(const GemvRingParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  // layout: [A: a_bytes][ring: 8 warps x depth x slot_bytes][full barriers 8 x 4][empty barriers 8 x 4][8 floats]
  uint8_t* sA = smem;
  uint8_t* ring = smem + p.a_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + static_cast<size_t>(RING_CONSUMERS) * p.depth * p.slot_bytes);
  uint64_t* empty = full + RING_CONSUMERS * RING_MAX_DEPTH;
  float* s_red = reinterpret_cast<float*>(empty + RING_CONSUMERS * RING_MAX_DEPTH);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_local = (p.n_out - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);  // rows b, b+G, ...
  const int nseg = p.K / p.seg_elems;
  if (threadIdx.x == 0) {
    for (int s = 0; s < RING_CONSUMERS * RING_MAX_DEPTH; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_launch_dependents();  // our successor may become resident as soon as there is room; it must griddepcontrol.wait before reading our output
  const uint32_t seg_bytes = static_cast<uint32_t>(p.seg_elems) * 2;
  if (warp == RING_CONSUMERS) {
    if (lane == 0) {
      // ---------------------------------------------------------------- producer: weights do not depend on any predecessor kernel
      const int max_rows_per_warp = (rows_local + RING_CONSUMERS - 1) / RING_CONSUMERS;
      for (int i = 0; i < max_rows_per_warp * nseg; ++i) {  // i: item index inside each warp's stream (row r = i / nseg, segment i % nseg)
        const int r = i / nseg, sg = i - r * nseg;
        const int slot = i % p.depth;
        for (int w = 0; w < RING_CONSUMERS; ++w) {
          const int j = r * RING_CONSUMERS + w;  // local row
          if (j >= rows_local) break;
          uint64_t* fb = &full[w * RING_MAX_DEPTH + slot];
          if (i >= p.depth) mbar_wait(&empty[w * RING_MAX_DEPTH + slot], ((i / p.depth) - 1) & 1);
          const int oc = blockIdx.x + j * gridDim.x;
          const int n0 = GLU ? (oc / 16) * 32 + (oc % 16) : oc;
          uint8_t* dst = ring + (static_cast<size_t>(w) * p.depth + slot) * p.slot_bytes;
          const __nv_bfloat16* src = p.B + static_cast<long long>(n0) * p.ldb + sg * p.seg_elems;
          mbar_arrive_expect_tx(fb, GLU ? 2 * seg_bytes : seg_bytes);
          bulk_load_1d(dst, src, seg_bytes, fb);
          if (GLU) bulk_load_1d(dst + seg_bytes, src + 16 * p.ldb, seg_bytes, fb);
        }
      }
    }
    return;
  }
  // ------------------------------------------------------------------ consumers
  pdl_wait();  // the activation vector (and the residual / output buffers) belong to predecessor kernels
  const int chunks = p.K >> 3;  // 8 elements per chunk
  {
    // stage A in shared memory. fp32: split into lo/hi float4 planes so that a warp's 16-byte loads are bank-conflict free.
    const int t = threadIdx.x;  // 0..255
    float ssq = 0.f;
    if (AF32) {
      const float4* ag = reinterpret_cast<const float4*>(p.A);
      float4* lo = reinterpret_cast<float4*>(sA);
      float4* hi = lo + chunks;
      for (int c = t; c < chunks; c += RING_CONSUMERS * 32) {
        const float4 a0 = __ldg(ag + 2 * c), a1 = __ldg(ag + 2 * c + 1);
        lo[c] = a0;
        hi[c] = a1;
        ssq += a0.x * a0.x + a0.y * a0.y + a0.z * a0.z + a0.w * a0.w + a1.x * a1.x + a1.y * a1.y + a1.z * a1.z + a1.w * a1.w;
      }
      ssq = warp_sum(ssq);
      if (lane == 0) s_red[warp] = ssq;
    } else {
      const uint4* ag = reinterpret_cast<const uint4*>(p.A);
      uint4* sa = reinterpret_cast<uint4*>(sA);
      for (int c = t; c < chunks; c += RING_CONSUMERS * 32) sa[c] = __ldg(ag + c);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");  // consumer warps only (the producer warp never joins)
  }
  float nrm = p.alpha;
  if (AF32) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < RING_CONSUMERS; ++w) tot += s_red[w];
    nrm = rsqrtf(tot / p.K + p.rms_eps) * p.alpha;
  }
  const int seg_chunks = p.seg_elems >> 3;
  int i = 0;  // item index in this warp's stream
  for (int j = warp; j < rows_local; j += RING_CONSUMERS) {
    float acc0 = 0.f, acc1 = 0.f;
    for (int sg = 0; sg < nseg; ++sg, ++i) {
      const int slot = i % p.depth;
      mbar_wait(&full[warp * RING_MAX_DEPTH + slot], (i / p.depth) & 1);
      const uint8_t* base = ring + (static_cast<size_t>(warp) * p.depth + slot) * p.slot_bytes;
      const uint4* w0 = reinterpret_cast<const uint4*>(base);
      const uint4* w1 = reinterpret_cast<const uint4*>(base + seg_bytes);
      const int c_off = sg * seg_chunks;
      if (AF32) {
        const float4* lo = reinterpret_cast<const float4*>(sA) + c_off;
        const float4* hi = lo + chunks;
#pragma unroll 4
        for (int c = lane; c < seg_chunks; c += 32) {
          const uint4 w = w0[c];
          const float4 a0 = lo[c], a1 = hi[c];
          acc0 += bf16_lo(w.x) * a0.x + bf16_hi(w.x) * a0.y + bf16_lo(w.y) * a0.z + bf16_hi(w.y) * a0.w + bf16_lo(w.z) * a1.x + bf16_hi(w.z) * a1.y +
                  bf16_lo(w.w) * a1.z + bf16_hi(w.w) * a1.w;
          if (GLU) {
            const uint4 g = w1[c];
            acc1 += bf16_lo(g.x) * a0.x + bf16_hi(g.x) * a0.y + bf16_lo(g.y) * a0.z + bf16_hi(g.y) * a0.w + bf16_lo(g.z) * a1.x + bf16_hi(g.z) * a1.y +
                    bf16_lo(g.w) * a1.z + bf16_hi(g.w) * a1.w;
          }
        }
      } else {
        const uint4* sa = reinterpret_cast<const uint4*>(sA) + c_off;
#pragma unroll 4
        for (int c = lane; c < seg_chunks; c += 32) {
          const uint4 w = w0[c], a = sa[c];
          acc0 += bf16_lo(w.x) * bf16_lo(a.x) + bf16_hi(w.x) * bf16_hi(a.x) + bf16_lo(w.y) * bf16_lo(a.y) + bf16_hi(w.y) * bf16_hi(a.y) +
                  bf16_lo(w.z) * bf16_lo(a.z) + bf16_hi(w.z) * bf16_hi(a.z) + bf16_lo(w.w) * bf16_lo(a.w) + bf16_hi(w.w) * bf16_hi(a.w);
          if (GLU) {
            const uint4 g = w1[c];
            acc1 += bf16_lo(g.x) * bf16_lo(a.x) + bf16_hi(g.x) * bf16_hi(a.x) + bf16_lo(g.y) * bf16_lo(a.y) + bf16_hi(g.y) * bf16_hi(a.y) +
                    bf16_lo(g.z) * bf16_lo(a.z) + bf16_hi(g.z) * bf16_hi(a.z) + bf16_lo(g.w) * bf16_lo(a.w) + bf16_hi(g.w) * bf16_hi(a.w);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[warp * RING_MAX_DEPTH + slot]);  // every lane's shared-memory reads of this slot are done
    }
    acc0 = warp_sum(acc0);
    if (GLU) acc1 = warp_sum(acc1);
    if (lane == 0) {
      const int oc = blockIdx.x + j * gridDim.x;
      const int n0 = GLU ? (oc / 16) * 32 + (oc % 16) : oc;
      float v = acc0 * nrm;
      if (p.bias) v += p.bias[n0];
      if (GLU) {
        float u = acc1 * nrm;
        if (p.bias) u += p.bias[n0 + 16];
        v = act_ring(v, p.act) * u;
      } else {
        v = act_ring(v, p.act);
      }
      if (p.col_scale) v *= p.col_scale[oc];
      if (p.resid) v += p.resid_f32 ? reinterpret_cast<const float*>(p.resid)[oc] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.resid)[oc]);
      if (p.out_f32)
        reinterpret_cast<float*>(p.C)[oc] = v;
      else
        reinterpret_cast<__nv_bfloat16*>(p.C)[oc] = __float2bfloat16(v);
    }
  }
}

static int ring_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

static int env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Returns VG_OK when the ring kernel handled the call, VG_ENOSUP when the caller should use the generic skinny kernel.
int gemv_ring_try(const VgGemm* g, cudaStream_t s) {
  static const int enabled = env_flag("VG_GEMV_RING", 1);
  if (!enabled || g->M != 1 || g->batch != 1 || (g->K % 8) || g->K < 1024) return VG_ENOSUP;
  const int n_out = g->glu ? g->N / 2 : g->N;
  if (n_out < 1024 || (g->glu && (g->N % 32))) return VG_ENOSUP;
  const bool af32 = g->a_dtype == VG_F32;
  if ((reinterpret_cast<uintptr_t>(g->A) & 15) || (reinterpret_cast<uintptr_t>(g->B) & 15) || (g->ldb % 8)) return VG_ENOSUP;
  GemvRingParams p;
  p.A = g->A; p.B = static_cast<const __nv_bfloat16*>(g->B); p.ldb = g->ldb; p.C = g->C;
  p.N = g->N; p.K = g->K; p.n_out = n_out; p.glu = g->glu; p.act = g->act;
  p.out_f32 = g->out_dtype == VG_F32; p.resid_f32 = g->resid_dtype == VG_F32;
  p.bias = static_cast<const float*>(g->bias); p.col_scale = static_cast<const float*>(g->col_scale); p.resid = g->resid;
  p.alpha = g->alpha; p.rms_eps = g->a_rms_eps;
  p.a_bytes = g->K * (af32 ? 4 : 2);
  // K-segment: the largest multiple of 256 elements (whole 16-byte chunks for every lane) dividing K with a slot of <= 6 KB
  const int slot_cap = 6144 / (g->glu ? 2 : 1) / 2;  // elements
  p.seg_elems = 0;
  for (int cand = slot_cap / 256 * 256; cand >= 256; cand -= 256)
    if (g->K % cand == 0) {
      p.seg_elems = cand;
      break;
    }
  if (p.seg_elems == 0) return VG_ENOSUP;
  p.slot_bytes = p.seg_elems * 2 * (g->glu ? 2 : 1);
  const int fixed = 2 * RING_CONSUMERS * RING_MAX_DEPTH * 8 + 64;
  const int budget = 112 * 1024 - p.a_bytes - fixed;  // two blocks per SM: (228 KB - 2 x 1 KB reserved) / 2 = 113 KB each
  p.depth = budget / (RING_CONSUMERS * p.slot_bytes);
  if (p.depth > RING_MAX_DEPTH) p.depth = RING_MAX_DEPTH;
  if (p.depth < 2) return VG_ENOSUP;
  const int smem = p.a_bytes + RING_CONSUMERS * p.depth * p.slot_bytes + fixed;
  const int grid = n_out < 2 * ring_num_sms() ? n_out : 2 * ring_num_sms();
  void (*kern)(const GemvRingParams) = af32 ? (g->glu ? gemv_ring_kernel<true, true> : gemv_ring_kernel<true, false>)
                                            : (g->glu ? gemv_ring_kernel<false, true> : gemv_ring_kernel<false, false>);
  static bool attr_done[4] = {false, false, false, false};
  const int ki = (af32 ? 2 : 0) + (g->glu ? 1 : 0);
  if (!attr_done[ki]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    if (e != cudaSuccess) return set_error(VG_ECUDA, "gemv_ring smem attr: %s", cudaGetErrorString(e));
    attr_done[ki] = true;
  }
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(RING_THREADS), smem, s, p);
  if (e != cudaSuccess) return set_error(VG_ECUDA, "gemv_ring launch: %s", cudaGetErrorString(e));
  count_launch();
  return check_launch("gemv_ring_kernel");
}

}  // namespace vg
