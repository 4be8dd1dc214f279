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

// ------------------------------------------------------------------------------------------------ M = 2..8 rows (batched decode)
// S = 2..8 sequences decode together: every weight row is streamed ONCE for all of them. With M rows the fp32-FMA formulation is issue-bound
// (measured: 640 instructions per 4 KB of weights at M = 4, 2.3 TB/s), so the products go to the tensor cores: the 8 weight rows of a tile are the
// B operand (n = 8) of mma.sync.m16n8k16, the M <= 8 activation rows the upper half of the A operand (rows 8..15 are zero registers). fp32
// activations are split into bf16 hi + lo in shared memory and take two MMAs, so the result keeps fp32-activation accuracy (weights are bf16
// anyway, accumulation fp32).
//   * one block per SM, ONE in-order ring of slots (8 weight rows x seg K-elements, seg = 1536..2048: rows padded by 16 B for conflict-free
//     ldmatrix), filled by 8 producer lanes with one 3-4 KB bulk copy each per slot (PDL: the weights start streaming as soon as the block is
//     resident; only the activations wait for the predecessor kernel);
//   * the 8 consumer warps split every slot by K (seg/8 elements each: 2-3 ldmatrix.x4 + 2-4 MMAs per 32 K-elements, four independent
//     accumulator chains) and carry the 8 x M partial
//     sums of the tile in registers across its K/seg slots; at the end of a tile the 8 partials meet in shared memory and 64 threads apply the
//     epilogue (fused RMSNorm scale, bias, SwiGLU, activation, column scale, residual);
//   * GLU: a tile is 4 gate rows + 4 up rows of the 16-interleaved gate_up layout, so that one MMA produces both halves of 4 outputs.
constexpr int RM_MAXSLOTS = 24;
constexpr int RM_CW = 8;
constexpr int RM_THREADS = (RM_CW + 1) * 32;        // + one producer warp (8 lanes issue; 1, 2 or 8 producer warps measured the same)

struct GemvRingMmaParams {
  const void* A; long long lda;   // [M][lda] fp32 (AF32) or bf16
  const __nv_bfloat16* B; long long ldb;
  void* C; long long ldc;
  const void* resid; long long ldr;
  int M, N, K, n_out, act, out_f32, resid_f32, nslots, a_stride, seg;   // a_stride: elements per staged activation row (K + 8); seg: K elements per slot
  const float* bias;
  const float* col_scale;
  float alpha, rms_eps;
};

template <bool AF32, bool GLU>
__global__ void __launch_bounds__(RM_THREADS, 1) gemv_ring_mma_kernel(const GemvRingMmaParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int RM_SEG = p.seg;                        // K elements per slot: bulk copies of seg x 2 bytes (>= 2 KB: the copy engine is ~100 cycles per copy)
  const int RM_ROWB = RM_SEG * 2 + 16;             // slot rows are padded by 16 B
  const int RM_SLOT = 8 * RM_ROWB;
  // layout: [zero 128 B][A hi: M x a_stride bf16][A lo (AF32)][ring: nslots x 8 x (2 seg + 16) B][full 24][empty 24][nrm 8][red 2 x 8 x 8 x 8]
  uint8_t* zero16 = smem;
  __nv_bfloat16* sAhi = reinterpret_cast<__nv_bfloat16*>(smem + 128);
  const size_t plane = static_cast<size_t>(p.M) * p.a_stride;
  __nv_bfloat16* sAlo = sAhi + plane;
  uint8_t* ring = reinterpret_cast<uint8_t*>(sAhi + (AF32 ? 2 : 1) * plane);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + static_cast<size_t>(p.nslots) * RM_SLOT);
  uint64_t* empty = full + RM_MAXSLOTS;
  float* s_nrm = reinterpret_cast<float*>(empty + RM_MAXSLOTS);
  float* red = s_nrm + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = p.n_out / (GLU ? 4 : 8);
  const int tiles_local = (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int nseg = p.K / RM_SEG;
  if (threadIdx.x < RM_MAXSLOTS) {
    mbar_init(&full[threadIdx.x], 1);
    mbar_init(&empty[threadIdx.x], RM_CW);
  }
  if (threadIdx.x < 4) reinterpret_cast<uint32_t*>(zero16)[threadIdx.x] = 0u;
  mbar_fence_init();
  __syncthreads();
  pdl_launch_dependents();
  if (warp >= RM_CW) {
    constexpr int RPW = 8;                         // rows per slot, one producer lane each
    if (lane < RPW) {
      // ---------------------------------------------------------------- producer: lane r streams row r of every tile (weights are constants)
      const int r = lane;
      int slot = 0, round = 0;
      for (int i = 0; i < tiles_local; ++i) {
        const int T = blockIdx.x + i * gridDim.x;
        int row;
        if (GLU) {
          const int oc = T * 4 + (r & 3);
          row = (oc / 16) * 32 + (oc % 16) + (r >= 4 ? 16 : 0);
        } else {
          row = T * 8 + r;
        }
        const __nv_bfloat16* src = p.B + static_cast<long long>(row) * p.ldb;
        for (int sg = 0; sg < nseg; ++sg) {
          if (lane == 0) {
            if (round > 0) mbar_wait(&empty[slot], (round - 1) & 1);
            mbar_arrive_expect_tx(&full[slot], RPW * RM_SEG * 2);
          }
          __syncwarp((1u << RPW) - 1u);
          bulk_load_1d(ring + static_cast<size_t>(slot) * RM_SLOT + r * RM_ROWB, src + sg * RM_SEG, RM_SEG * 2, &full[slot]);
          if (++slot == p.nslots) {
            slot = 0;
            ++round;
          }
        }
      }
    }
    return;
  }
  // ------------------------------------------------------------------ consumers
  pdl_wait();  // the activation rows (and the residual / output buffers) belong to predecessor kernels
  {
    const int chunks = p.K >> 3;
    if (AF32) {
      // warp w stages rows w, w + 4: fp32 -> bf16 hi + bf16 lo, sum of squares of the row for the fused RMSNorm
      for (int m = warp; m < p.M; m += RM_CW) {
        const float4* ag = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.A) + m * p.lda);
        uint4* hi = reinterpret_cast<uint4*>(sAhi + static_cast<size_t>(m) * p.a_stride);
        uint4* lo = reinterpret_cast<uint4*>(sAlo + static_cast<size_t>(m) * p.a_stride);
        float ssq = 0.f;
        for (int c0 = lane; c0 < chunks; c0 += 32 * 6) {
          float4 v[12];
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int c = c0 + 32 * u;
            v[2 * u] = v[2 * u + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < chunks) {
              v[2 * u] = __ldg(ag + 2 * c);
              v[2 * u + 1] = __ldg(ag + 2 * c + 1);
            }
          }
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int c = c0 + 32 * u;
            if (c < chunks) {
              const float x[8] = {v[2 * u].x, v[2 * u].y, v[2 * u].z, v[2 * u].w, v[2 * u + 1].x, v[2 * u + 1].y, v[2 * u + 1].z, v[2 * u + 1].w};
              uint32_t h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat16 h0 = __float2bfloat16(x[2 * e]), h1 = __float2bfloat16(x[2 * e + 1]);
                const __nv_bfloat16 l0 = __float2bfloat16(x[2 * e] - __bfloat162float(h0)), l1 = __float2bfloat16(x[2 * e + 1] - __bfloat162float(h1));
                h[e] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
                l[e] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
                ssq += x[2 * e] * x[2 * e] + x[2 * e + 1] * x[2 * e + 1];
              }
              hi[c] = make_uint4(h[0], h[1], h[2], h[3]);
              lo[c] = make_uint4(l[0], l[1], l[2], l[3]);
            }
          }
        }
        ssq = warp_sum(ssq);
        if (lane == 0) s_nrm[m] = rsqrtf(ssq / p.K + p.rms_eps) * p.alpha;
      }
    } else {
      const int t = threadIdx.x;  // consumer threads
      const int total = p.M * chunks;
      for (int i0 = t; i0 < total; i0 += RM_CW * 32 * 8) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + RM_CW * 32 * u;
          if (i < total) {
            const int m = i / chunks, c = i - m * chunks;
            v[u] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.A) + m * p.lda) + c);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + RM_CW * 32 * u;
          if (i < total) {
            const int m = i / chunks, c = i - m * chunks;
            reinterpret_cast<uint4*>(sAhi + static_cast<size_t>(m) * p.a_stride)[c] = v[u];
          }
        }
      }
      if (t < 8) s_nrm[t] = p.alpha;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(RM_CW * 32) : "memory");  // consumer warps only (the producer warps never join)
  }
  const int g = lane >> 2, t4 = lane & 3;
  const int mrow = lane & 7, mq = lane >> 3;       // ldmatrix: this lane addresses row mrow of 8x8 matrix mq (K offset 8 mq inside a 32-K group)
  const int kw = warp * (RM_SEG / RM_CW);          // this warp's K range inside every slot
  const bool arow = mrow < p.M;
  const uint8_t* a_hi = arow ? reinterpret_cast<const uint8_t*>(sAhi + static_cast<size_t>(mrow) * p.a_stride + kw + 8 * mq) : zero16;
  const uint8_t* a_lo = arow ? reinterpret_cast<const uint8_t*>(sAlo + static_cast<size_t>(mrow) * p.a_stride + kw + 8 * mq) : zero16;
  const int a_mul = arow ? 2 : 0;                  // bytes per K element for real rows, 0 keeps the zero chunk
  const uint32_t w_off = mrow * RM_ROWB + (kw + 8 * mq) * 2;
  int slot = 0, phase = 0, buf = 0;
  for (int i = 0; i < tiles_local; ++i) {
    const int T = blockIdx.x + i * gridDim.x;
    // four independent accumulator chains (an mma.sync is ~30 cycles deep; one chain of 4 MMAs per 32 K-elements capped the kernel at 3.3 TB/s)
    float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f}, d3[4] = {0.f, 0.f, 0.f, 0.f};
    for (int sg = 0; sg < nseg; ++sg) {
      mbar_wait(&full[slot], phase);
      const uint8_t* wb = ring + static_cast<size_t>(slot) * RM_SLOT + w_off;
      const int a_off = a_mul * (sg * RM_SEG);
#pragma unroll 2
      for (int q = 0; q < RM_SEG / RM_CW / 32; q += 2) {     // two 32-K groups per iteration
        uint32_t b[2][4], ah[2][4], al[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          ldmatrix_x4(b[h], wb + (q + h) * 64);
          ldmatrix_x4(ah[h], a_hi + a_off + a_mul * ((q + h) * 32));
          if (AF32) ldmatrix_x4(al[h], a_lo + a_off + a_mul * ((q + h) * 32));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t x0[4] = {ah[h][0], 0u, ah[h][1], 0u}, x1[4] = {ah[h][2], 0u, ah[h][3], 0u};
          if (AF32) {
            const uint32_t y0[4] = {al[h][0], 0u, al[h][1], 0u}, y1[4] = {al[h][2], 0u, al[h][3], 0u};
            mma_bf16_16816(d0, x0, b[h][0], b[h][1]);
            mma_bf16_16816(d1, x1, b[h][2], b[h][3]);
            mma_bf16_16816(d2, y0, b[h][0], b[h][1]);
            mma_bf16_16816(d3, y1, b[h][2], b[h][3]);
          } else if (h == 0) {
            mma_bf16_16816(d0, x0, b[h][0], b[h][1]);
            mma_bf16_16816(d1, x1, b[h][2], b[h][3]);
          } else {
            mma_bf16_16816(d2, x0, b[h][0], b[h][1]);
            mma_bf16_16816(d3, x1, b[h][2], b[h][3]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
      if (++slot == p.nslots) {
        slot = 0;
        phase ^= 1;
      }
    }
    const float d[2] = {(d0[0] + d1[0]) + (d2[0] + d3[0]), (d0[1] + d1[1]) + (d2[1] + d3[1])};
    // the 8 K-slices of the tile meet in shared memory: red[buf][warp][m = g][n = 2 t4, 2 t4 + 1]
    float* rb = red + buf * (RM_CW * 64);
    *reinterpret_cast<float2*>(rb + warp * 64 + g * 8 + 2 * t4) = make_float2(d[0], d[1]);
    asm volatile("bar.sync 1, %0;" ::"n"(RM_CW * 32) : "memory");
    const int tid = threadIdx.x;
    if (tid < 64) {
      const int m = tid >> 3, n = tid & 7;
      if (m < p.M && (!GLU || n < 4)) {
        const float nrm = s_nrm[m];
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < RM_CW; ++w) v += rb[w * 64 + tid];
        v *= nrm;
        const int oc = GLU ? T * 4 + n : T * 8 + n;
        if (GLU) {
          const int n0 = (oc / 16) * 32 + (oc % 16);
          float u = 0.f;
#pragma unroll
          for (int w = 0; w < RM_CW; ++w) u += rb[w * 64 + tid + 4];
          u *= nrm;
          if (p.bias) {
            v += p.bias[n0];
            u += p.bias[n0 + 16];
          }
          v = act_ring(v, p.act) * u;
        } else {
          if (p.bias) v += p.bias[oc];
          v = act_ring(v, p.act);
        }
        if (p.col_scale) v *= p.col_scale[oc];
        if (p.resid) {
          const long long ro = static_cast<long long>(m) * p.ldr + oc;
          v += p.resid_f32 ? reinterpret_cast<const float*>(p.resid)[ro] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.resid)[ro]);
        }
        const long long co = static_cast<long long>(m) * p.ldc + oc;
        if (p.out_f32)
          reinterpret_cast<float*>(p.C)[co] = v;
        else
          reinterpret_cast<__nv_bfloat16*>(p.C)[co] = __float2bfloat16(v);
      }
    }
    buf ^= 1;
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

// M = 2..8 rows (batched decode). Returns VG_ENOSUP when the caller should use the generic skinny kernel.
int gemv_ring_mr_try(const VgGemm* g, cudaStream_t s) {
  static const int enabled = env_flag("VG_GEMV_RING_MR", 1);
  static const int smem_kb = env_flag("VG_GEMV_RING_MR_SMEM_KB", 224);
  if (!enabled || g->M < 2 || g->M > 8 || g->batch != 1 || (g->K % (RM_CW * 64)) || g->K < 1024) return VG_ENOSUP;
  const int n_out = g->glu ? g->N / 2 : g->N;
  if (n_out < 1024 || (g->glu && (g->N % 32)) || (!g->glu && (n_out % 8))) return VG_ENOSUP;
  const bool af32 = g->a_dtype == VG_F32;
  if (!af32 && g->a_rms_eps > 0.f) return VG_ENOSUP;
  if ((reinterpret_cast<uintptr_t>(g->A) & 15) || (reinterpret_cast<uintptr_t>(g->B) & 15) || (g->ldb % 8) || (g->lda % (af32 ? 4 : 8))) return VG_ENOSUP;
  GemvRingMmaParams p;
  p.A = g->A; p.lda = g->lda; p.B = static_cast<const __nv_bfloat16*>(g->B); p.ldb = g->ldb; p.C = g->C; p.ldc = g->ldc;
  p.resid = g->resid; p.ldr = g->ldr;
  p.M = g->M; p.N = g->N; p.K = g->K; p.n_out = n_out; p.act = g->act;
  p.out_f32 = g->out_dtype == VG_F32; p.resid_f32 = g->resid_dtype == VG_F32;
  p.bias = static_cast<const float*>(g->bias); p.col_scale = static_cast<const float*>(g->col_scale);
  p.alpha = g->alpha; p.rms_eps = g->a_rms_eps;
  p.a_stride = g->K + 8;
  static const int seg_cap = env_flag("VG_GEMV_RING_MR_SEG", 2048);
  const int a_bytes = g->M * p.a_stride * 2 * (af32 ? 2 : 1);
  const int fixed = 128 + 2 * RM_MAXSLOTS * 8 + 32 + 2 * RM_CW * 64 * 4 + 64;
  // K elements per slot: the largest multiple of 64 x consumer warps dividing K, <= seg_cap, that still leaves a ring of >= 4 slots
  p.seg = 0;
  int RM_SLOT = 0;
  for (int cand = seg_cap / (RM_CW * 64) * (RM_CW * 64); cand >= RM_CW * 64; cand -= RM_CW * 64) {
    if (g->K % cand) continue;
    RM_SLOT = 8 * (cand * 2 + 16);
    p.nslots = (smem_kb * 1024 - a_bytes - fixed) / RM_SLOT;
    if (p.nslots >= 4) {
      p.seg = cand;
      break;
    }
  }
  if (p.seg == 0) return VG_ENOSUP;
  if (p.nslots > RM_MAXSLOTS) p.nslots = RM_MAXSLOTS;
  const int smem = a_bytes + p.nslots * RM_SLOT + fixed;
  const int ntiles = n_out / (g->glu ? 4 : 8);
  const int grid = ntiles < ring_num_sms() ? ntiles : ring_num_sms();
  void (*kern)(const GemvRingMmaParams) = af32 ? (g->glu ? gemv_ring_mma_kernel<true, true> : gemv_ring_mma_kernel<true, false>)
                                               : (g->glu ? gemv_ring_mma_kernel<false, true> : gemv_ring_mma_kernel<false, false>);
  static bool attr_done[4] = {false, false, false, false};
  const int ki = (af32 ? 2 : 0) + (g->glu ? 1 : 0);
  if (!attr_done[ki]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return set_error(VG_ECUDA, "gemv_ring_mma smem attr: %s", cudaGetErrorString(e));
    attr_done[ki] = true;
  }
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(RM_THREADS), smem, s, p);
  if (e != cudaSuccess) return set_error(VG_ECUDA, "gemv_ring_mma launch: %s", cudaGetErrorString(e));
  count_launch();
  return check_launch("gemv_ring_mma_kernel");
}

}  // namespace vg
