/*
 * libvglamm — C ABI of the B200-native (sm_100a) kernels behind the VideoGLaMM grounded-generation forward pass.
 *
 * The reference (mbzuai-oryx/VideoGLaMM) has no FFI for this path: its arithmetic is torch.nn modules
 * (SURVEY.md §8b). Each entry point below therefore cites the reference call site(s) whose arithmetic it
 * replaces. Conventions, identical for every function:
 *   - all pointers are DEVICE pointers owned by the caller; kernels never allocate;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), nothing synchronises;
 *   - return 0 on success, a negative VG_E* code otherwise (no exceptions cross the boundary);
 *     vg_last_error() returns a thread-local message for the last failure;
 *   - "bf16" = __nv_bfloat16 bits, row-major, leading dimensions in ELEMENTS.
 */
#ifndef VGLAMM_H_
#define VGLAMM_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VG_OK 0
#define VG_EINVAL (-1)   /* bad argument (shape / alignment / null) */
#define VG_ECUDA (-2)    /* CUDA runtime / driver error */
#define VG_ENOSUP (-3)   /* configuration not supported by the kernels */

#define VG_BF16 0
#define VG_F32 1

#define VG_ACT_NONE 0
#define VG_ACT_GELU 1      /* exact erf GELU (nn.GELU())                      */
#define VG_ACT_QUICKGELU 2 /* x*sigmoid(1.702x) (HF CLIP "quick_gelu")        */
#define VG_ACT_RELU 3
#define VG_ACT_SILU 4
#define VG_ACT_SIGMOID 5

const char* vg_last_error(void);
int vg_version(void);
/* Number of kernel launches issued through this library since load / last reset (bench.py "gpu_launches"). */
long long vg_launch_count(void);
void vg_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------------ GEMM
 * C[b] = epilogue( alpha * A[b] (M x K)  *  B[b]^T (B is N x K, i.e. nn.Linear weight layout) )
 * epilogue, in this order:  + bias[n]  ->  act  ->  (glu: act(gate)*up)  ->  * col_scale[n]  ->  + resid[m,n]
 * Replaces every nn.Linear / 1x1 Conv2d / patch-embedding contraction on the path, e.g.
 *   HF Phi3 qkv_proj/o_proj/gate_up_proj/down_proj (called from VideoGLaMM/model/videogpt_plus/model/language_model/phi3.py:72-84),
 *   InternVideo2 Attention/Mlp (VideoGLaMM/model/videogpt_plus/model/internvideo/internvideo2.py:190-262),
 *   Hiera MultiScaleAttention.qkv/proj + MLP (VideoGLaMM/model/segment_anything_2/sam2/modeling/backbones/hieradet.py:58-168),
 *   SAM2 mask-decoder Attention projections / MLPs / ConvTranspose2d-as-GEMM (.../sam2/modeling/sam/transformer.py:196-260,
 *   .../sam/mask_decoder.py:221-240).
 * Kernel: TMA (SWIZZLE_128B) -> shared -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue.
 * Requirements: lda, ldb multiples of 8; A, B 16-byte aligned; K >= 1 (K tail zero-filled by TMA).
 * With glu != 0 the N columns are interleaved [16 gate | 16 up] per 32 and the output has N/2 columns.
 */
typedef struct VgGemm {
  const void* A; long long lda; long long strideA; /* bf16 [batch][M][lda]   */
  const void* B; long long ldb; long long strideB; /* bf16 [batch][N][ldb]   */
  void* C; long long ldc; long long strideC;       /* out  [batch][M][ldc]   */
  int M, N, K, batch;
  int out_dtype;                                   /* VG_BF16 | VG_F32       */
  const float* bias;                               /* [N] or NULL            */
  int act;                                         /* VG_ACT_*               */
  int glu;                                         /* 0 | 1                  */
  const float* col_scale;                          /* [N_out] or NULL        */
  const void* resid; long long ldr; long long strideR; int resid_dtype; /* optional residual, may alias C */
  float alpha;
  int force_block_n;                               /* 0 = auto, else 64/128/256 (tests) */
} VgGemm;
int vg_gemm(const VgGemm* g, void* stream);
/* Same contract, one thread per output element, fp32 accumulate in k order. Test/diagnostic reference ON DEVICE. */
int vg_gemm_naive(const VgGemm* g, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VGLAMM_H_ */
