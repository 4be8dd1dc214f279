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
 * batch > 1 with strideB == 0 shares one B (weights) across all batch entries.
 * With glu != 0 the N columns are interleaved [16 gate | 16 up] per 32 and the output has N/2 columns.
 * M <= 8 (LLM decode, one row per decoding sequence) is HBM-bound weight streaming and leaves the tcgen05 path: M == 1 -> gemv_ring_kernel
 * (cp.async.bulk rings, fp32 FMA), 2 <= M <= 8 -> gemv_ring_mma_kernel (the same rings feeding mma.sync, every weight row streamed once for all
 * rows), shapes those decline (N_out < 1024, K not a multiple of 512) -> gemv_kernel; force_block_n != 0 keeps M <= 8 on the tcgen05 kernel.
 */
typedef struct VgGemm {
  const void* A; long long lda; long long strideA; /* bf16 [batch][M][lda]   */
  const void* B; long long ldb; long long strideB; /* bf16 [batch][N][ldb]   */
  void* C; long long ldc; long long strideC;       /* out  [batch][M][ldc]   */
  int M, N, K, batch;
  int out_dtype;                                   /* VG_BF16 | VG_F32       */
  const float* bias; long long strideBias;         /* [N] or NULL; per-batch stride (0 = shared) */
  int act;                                         /* VG_ACT_*               */
  int glu;                                         /* 0 | 1                  */
  const float* col_scale;                          /* [N_out] or NULL        */
  const void* resid; long long ldr; long long strideR; int resid_dtype; /* optional residual, may alias C */
  float alpha;
  int force_block_n;                               /* 0 = auto, else 64/128/256 (tests) */
  int a_dtype;                                     /* VG_BF16 (default) | VG_F32: fp32 A, only for M <= 8 (decode) */
  float a_rms_eps;                                 /* with fp32 A: rows are RMS-normalised on the fly, x*rsqrt(mean(x^2)+eps)
                                                      (the norm weight must be folded into B); 0 is a valid eps */
} VgGemm;
int vg_gemm(const VgGemm* g, void* stream);
/* Same contract, one thread per output element, fp32 accumulate in k order. Test/diagnostic reference ON DEVICE. */
int vg_gemm_naive(const VgGemm* g, void* stream);

/* ------------------------------------------------------------------------------------------------ attention
 * O = softmax(scale * Q K^T + mask) V, flash style (scores never materialised), bf16 in/out, fp32 softmax.
 * Element (b,h,row,d) of X lives at X + b*x_bs + h*x_hs + row*x_rs + d, so a packed [tokens, 3*H*D] qkv GEMM output
 * is consumed in place. causal: query i sees keys j <= i + (Nk-Nq). window>0 additionally requires
 * j >= i + (Nk-Nq) - window (transformers 4.41 Phi-3 sliding-window rule).
 * 2D window addressing (Hiera): if *_win > 0 the batch index enumerates (image, window_y, window_x) and token t of a
 * window maps to row img*(nwy*win*gridw) + (wy*win + t/win)*gridw + wx*win + t%win of the un-partitioned token grid
 * (replaces window_partition/unpartition copies, .../sam2/modeling/backbones/utils.py:16-62); *_bs is ignored then.
 * Replaces F.scaled_dot_product_attention at .../backbones/hieradet.py:72-76, .../sam/transformer.py:249-255,316-322,
 * InternVideo2 _naive_attn (internvideo2.py:190-209), HF CLIP / Phi-3 attention (called from clip_encoder.py:45-72, phi3.py:72-81).
 * Requirements: D % 8 == 0, D <= 256; q/k/v strides multiples of 8 elements, 16-byte aligned bases.
 * Kernels: problems with Nq, Nk >= 128 (dense, causal, causal + sliding window, 16x16 Hiera windows) run on the tcgen05 kernels
 * (attention_tc.cu: TMA -> smem -> tcgen05.mma, S and O in TMEM; two query tiles per block when D <= 128 and the problem is >= 1024 blocks deep);
 * pooled-query windows and tiny problems use the mma.sync kernels; Nq == 1 uses the split-K decode kernel.
 */
typedef struct VgAttention {
  const void *q, *k, *v; void* o;
  long long q_bs, q_hs, q_rs, k_bs, k_hs, k_rs, v_bs, v_hs, v_rs, o_bs, o_hs, o_rs;
  int B, H, Nq, Nk, D;
  float scale;
  int causal, window;
  int q_win, q_gridw, q_nwx, q_nwy;
  int k_win, k_gridw, k_nwx, k_nwy;
  void* workspace; long long workspace_bytes;   /* optional fp32 scratch enabling split-K over keys: single-query (decode) attention needs
                                                   B*H*16*(max(D,128)+2)*4 bytes; non-causal attention with < 148 query tiles (SAM2 memory
                                                   attention) uses up to 8*B*H*Nq*(Dpad+2)*4 bytes */
  int ws_has_counters;                          /* the LAST B*H ints of `workspace` are zero-initialised ticket counters owned by the
                                                   library (kept zero between calls): single-query split-K then merges in-kernel */
  int force_legacy;                             /* 1: skip the tcgen05 kernel and use the mma.sync kernel (tests / A-B timing) */
  const int* nk_dev;                            /* optional (Nq == 1 only): the number of keys is read from this DEVICE int at run time
                                                   (Nk is then only an upper bound) so a captured decode step can be replayed */
  float fused_rope_theta;                       /* > 0 (Nq == 1, D <= 128 only; batch entry b uses the packed row at q + b*q_bs and its own caches at k/v + b*k_bs): the decode step's RoPE + KV-cache append is fused: q points
                                                   at the UN-rotated packed row [q | k_new | v_new] (3*H*D bf16) of the newest token at position
                                                   Nk-1; the kernel rotates q on the fly, rotates k_new, writes k_new / v_new into row Nk-1 of the
                                                   k / v cache views (which must be writable) and attends over keys [0, Nk). Replaces
                                                   vg_rope_kv_append + vg_attention (one launch less per layer and token). */
  int nk_per_batch;                             /* 1: nk_dev is an array of B key counts, one per batch entry (batched decode of sequences of different lengths) */
} VgAttention;
int vg_attention(const VgAttention* a, void* stream);

/* ------------------------------------------------------------------------------------------------ normalisation
 * Row-wise LayerNorm (rms=0: (x-mean)/sqrt(var+eps)*w+b) or RMSNorm (rms=1: x/sqrt(mean(x^2)+eps)*w), fp32 statistics.
 * x: [rows][ldx] (x_dtype), y: [rows][ldy] (y_dtype). LayerNorm2d on channels-last tensors is the same operation.
 * Replaces nn.LayerNorm / LayerNorm2d (.../sam2/modeling/sam2_utils.py:137-149), InternVideo2 RMSNorm (internvideo2.py:134-145),
 * HF Phi3RMSNorm. b may be NULL.
 */
int vg_norm(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, long long ldy, const float* w, const float* b,
            int rows, int cols, float eps, int rms, void* stream);

/* ------------------------------------------------------------------------------------------------ elementwise / layout
 * out[r][c] = a[r % a_rows][c] + alpha * b[r % b_rows][c]   (b NULL: plain dtype cast). Row strides in elements.
 * Covers q = queries + query_pe / k = keys + key_pe (.../sam/transformer.py:160-193), + pos_embed, + no_mem_embed
 * (.../sam2_base.py:641-645), curr + 0.1*curr_pos (.../memory_attention.py:140-141).
 */
int vg_add(const void* a, int a_dtype, long long lda, int a_rows, const void* b, int b_dtype, long long ldb, int b_rows, float alpha,
           void* out, int out_dtype, long long ldo, int rows, int cols, void* stream);
/* out[i] = resid[i] + sum_p parts[p*n + i], fp32 (resid may be NULL or alias out): sums the K-split partial products of a skinny projection
 * (batched LLM decode: o_proj / down_proj are run as `nparts` batched GEMMs over K slices so that every SM streams weights). */
int vg_sum_parts(const float* parts, long long n, int nparts, const float* resid, float* out, void* stream);
/* 2x2 max (is_max=1; Hiera q-pool, hieradet.py:21-33) or mean (is_max=0; arch.py:88-96 adaptive pooling by 2) over a
 * channels-last [n][h][w][c] grid; row stride of the input in elements (reads a column slice of a wider tensor). */
int vg_pool2x2(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, long long ldy, int n, int h, int w, int c, int is_max,
               void* stream);
/* out[n][2y+dy][2x+dx][c] = lat[n][2y+dy][2x+dx][c] + top[n][y][x][c]  (FPN nearest top-down, image_encoder.py:115-126); fp32 */
int vg_upsample2x_add(const float* lat, const float* top, float* out, int n, int h, int w, int c, void* stream);
/* ConvTranspose2d(k=2,s=2) epilogue after the [n*h*w, 4*cout] GEMM (columns ordered (dy,dx,co)): pixel-shuffle, + bias[co],
 * + feat[n][2h][2w][cout] (may be NULL; feat_bcast=1: one [2h][2w][cout] map shared by all n), optional LayerNorm2d(ln_w, ln_b, eps)
 * then GELU. (mask_decoder.py:221-225) */
int vg_convt2x2_epilogue(const void* g, int g_dtype, const float* bias, const void* feat, int feat_dtype, const float* ln_w,
                         const float* ln_b, float eps, void* out, int out_dtype, int n, int h, int w, int cout, int feat_bcast, void* stream);
/* im2col for Conv2d(kernel k, stride s, pad p) on NCHW fp32/bf16 input -> bf16 (or fp32) [n*oh*ow][kpad], column order (ci,ky,kx),
 * zero-filled to kpad (multiple of 8). Patch embeddings: Hiera 7x7/4 (backbones/utils.py:65-95), CLIP 14x14/14, IV2 1x14x14. */
int vg_im2col(const void* x, int x_dtype, void* out, int out_dtype, int n, int cin, int h, int w, int k, int s, int p, int oh, int ow, int kpad,
              void* stream);
/* Depthwise KxK conv, stride 1, pad K/2, channels-last [n][h][w][c]; weight [c][K][K] fp32 (CXBlock.dwconv, memory_encoder.py:62-118). */
int vg_dwconv(const void* x, int x_dtype, const float* w, const float* bias, void* y, int y_dtype, int n, int h, int wd, int c, int k,
              void* stream);
/* Conv2d 3x3 stride 2 pad 1, channels-last, weights [ (ky,kx,ci) ][cout] fp32, + bias, then LayerNorm2d + GELU
 * (MaskDownSampler stage, memory_encoder.py:17-58). x: [n][h][w][cin] -> y: [n][h/2][w/2][cout]. */
int vg_conv3x3s2_ln_gelu(const void* x, int x_dtype, const float* w, const float* bias, const float* ln_w, const float* ln_b, float eps,
                         void* y, int y_dtype, int n, int h, int wd, int cin, int cout, void* stream);
/* The same convolution as a tensor-core GEMM, for cin >= 16: (1) im2col of the channels-last [n][h][w][cin] input (cin % 8 == 0) to bf16
 * [n*(h/2)*(w/2)][9*cin], column order (ky,kx,ci); (2) vg_gemm with the weights as [cout][(ky,kx,ci)] + bias, fp32 out;
 * (3) LayerNorm2d + GELU on the fp32 rows (cols % 4 == 0, cols <= 640), bf16 or fp32 out. (MaskDownSampler stages 16->64 and 64->256,
 * memory_encoder.py:17-58) */
int vg_im2col3x3s2_cl(const void* x, int x_dtype, void* out, int n, int h, int w, int cin, void* stream);
int vg_layernorm_gelu(const float* x, long long ldx, void* y, int y_dtype, long long ldy, const float* w, const float* b, int rows, int cols,
                      float eps, void* stream);
/* Bilinear resize (align_corners=False, no antialias) of fp32 planes [n][ih][iw] -> [n][oh][ow] with a fused output map:
 * mode 0: identity; 1: sigmoid(x)*scale+bias; 2: (x>0)*scale+bias (sam2_base.py:366-374, 687-697; sam2_video_predictor.py:498-518);
 * mode 3: write uint8 (x > 0) instead of fp32 (VideoGLaMM.py:870-875); mode 4: write (x > 0) BIT-PACKED, 8 pixels per byte MSB first (numpy.packbits
 * order), each [oh][ow] plane padded to ceil(oh*ow/8) bytes — the form in which masks leave the GPU (D2H and the multi-GPU gather). */
int vg_bilinear(const float* x, void* y, int n, int ih, int iw, int oh, int ow, int mode, float scale, float bias, void* stream);
/* 2D axial RoPE on rows of a [rows][ld] bf16 / fp32 tensor, in place: row r has grid position (r % gw, (r / gw) % gh); the first
 * dim/2 features rotate with x, the rest with y, consecutive (even,odd) pairs are one complex number
 * (position_encoding.py:174-216). Rows [0, nrows). theta 10000. */
int vg_rope2d(void* x, int x_dtype, long long ld, int nrows, int dim, int gw, int gh, void* stream);
/* The same rotation from a table: vg_rope2d_table fills table[gw*gh][dim/2] = (cos, sin) pairs (fp32, 2 floats each) once; vg_rope2d_cached
 * rotates `nblk` adjacent column blocks of width dim of every row (q and k of a packed qkv row in one launch), row r at table position r % period. */
int vg_rope2d_table(float* table, int dim, int gw, int gh, void* stream);
int vg_rope2d_cached(void* x, int x_dtype, long long ld, int nrows, int dim, int nblk, const float* table, int period, void* stream);
/* Phi-3 / Llama rotate-half RoPE in place on q and k slices of a packed qkv tensor [rows][ld] (bf16 / fp32): for each head, pairs (i, i+hd/2).
 * pos[r] is the position of row r. */
int vg_rope_neox(void* qkv, int x_dtype, long long ld, const int* pos, int rows, int n_heads_total, int head_dim, float theta, void* stream);
/* out[r][:] = table[ids[r]][:] (nn.Embedding), bf16 table -> out dtype */
int vg_embedding(const void* table, long long ldt, const long long* ids, void* out, int out_dtype, long long ldo, int rows, int cols, void* stream);
/* cache[*pos_dev][col0 : col0+cols] = src[0][col0 : col0+cols] for a bf16 row (KV-cache append at a device-side position, so that
 * the decode step is position-independent at launch level and can be captured in a CUDA graph and replayed). */
int vg_row_scatter(const void* src, void* cache, long long ld_cache, const int* pos_dev, int col0, int cols, void* stream);
/* Decode-step helper (one launch instead of RoPE + two scatters): rotate-half RoPE at position *pos_dev on the q and k heads of the
 * packed bf16 row qkv[0 : 3*H*hd] in place, then cache[*pos_dev][H*hd : 3*H*hd] = (k, v). */
int vg_rope_kv_append(void* qkv, void* cache, long long ld_cache, const int* pos_dev, int n_heads, int head_dim, float theta, void* stream);
/* MaskDecoder._dynamic_multimask_via_stability (.../sam2/modeling/sam/mask_decoder.py:247-295): per object, index 0 if the stability score
 * of mask 0 (area(x > delta) / area(x > -delta), 1 if the union is empty) is >= thresh, else the best-IoU index among masks 1..3.
 * masks fp32 [n_obj][4][hw], iou fp32 [n_obj][4] -> out_idx int64 [n_obj]. */
int vg_multimask_select(const float* masks, const float* iou, int n_obj, int hw, float delta, float thresh, long long* out_idx, void* stream);
/* adds n to the launch counter (replayed CUDA-graph nodes are launches too) */
void vg_launch_count_add(long long n);
/* 8-connected components of binary masks: the reference's only native extension, sam2._C.get_connected_componnets
 * (.../sam2/csrc/connected_components.cu:213-289, bound in .../sam2/utils/misc.py:47-63). mask uint8 [n][h][w] (nonzero = foreground; the
 * reference's [N,1,H,W] contiguous tensor), h and w even (the reference asserts the same) -> labels int32 [n][h][w]: 0 on the background, else
 * 1 + min over the component of the linear index of the top-left pixel of each member's 2x2 block (exactly the reference's label values);
 * counts int32 [n][h][w]: component area on foreground pixels, 0 elsewhere. Caller-owned outputs and workspace (>= 2*n*h*w*4 bytes), runs on
 * `stream`, errors are return codes (the reference raises through AT_ASSERTM). */
int vg_connected_components(const unsigned char* mask, int* labels, int* counts, int n, int h, int w, void* workspace, long long workspace_bytes,
                            void* stream);
/* ------------------------------------------------------------------------------------------------ fp32-grade arithmetic
 * The mask logits are sensitive to the rounding of the mask decoder's operands (north_star: 1e-3 on mask logits; bf16 operands give ~5e-3,
 * scripts/precision_study.py), so its contractions run as split-bf16 GEMMs and its small attentions in exact fp32:
 * vg_split3: x fp32 [rows][ldx] (cols used) -> bf16 [rows][ldo >= 3*kp]: columns [0,kp) = hi = bf16(x), then (A side, b_side=0) lo = bf16(x - hi), hi
 * or (B side, b_side=1) hi, lo; columns [cols,kp) of each third are zero (kp % 8 == 0). One vg_gemm over K' = 3*kp on an A-side and a B-side
 * operand computes A_hi B_hi + A_lo B_hi + A_hi B_lo with fp32 accumulation (relative error ~2^-17).
 * Replaces (at fp32 fidelity) the nn.Linear / ConvTranspose2d / hyper_in @ upscaled contractions of .../sam2/modeling/sam/mask_decoder.py:167-245,
 * .../sam/transformer.py:160-260. */
int vg_split3(const float* x, long long ldx, void* out, long long ldo, long long rows, int cols, int kp, int b_side, void* stream);
/* vg_attention with fp32 q/k/v/o (same struct, strides in fp32 elements; no alignment requirement; D <= 256): exact fp32 softmax(scale QK^T)V with
 * online softmax, expf and FMA accumulation; optional split over keys through `workspace` (ns*B*H*Nq*(D+2)*4 bytes) for few-query problems.
 * Product path: the two-way transformer's attentions (.../sam/transformer.py:233-260: 9x9, 9x4096, 4096x9 tokens). Verification mode: every
 * attention of the path (causal, sliding window and Hiera 2D windows supported; nk_dev / fused_rope_theta are not). */
int vg_attention_f32(const VgAttention* a, void* stream);
/* ------------------------------------------------------------------------------------------------ fused SAM2 mask decoder
 * ONE launch for: src = image_embed + dense prompt (no_mask_embed), TwoWayTransformer (2 x {token self-attention, token->image attention, MLP 2048,
 * image->token attention} + final token->image attention + LayerNorm), ConvTranspose2d(256->64,k2,s2) + feat_s1 + LayerNorm2d + GELU,
 * ConvTranspose2d(64->32,k2,s2) + feat_s0 + GELU, the 4 hypernetwork MLPs, masks = hyper_in @ upscaled, IoU head (sigmoid) and object-score head —
 * MaskDecoder.predict_masks .../sam2/modeling/sam/mask_decoder.py:167-245, TwoWayTransformer .../sam/transformer.py:63-115, 160-193, 233-260 —
 * for all B objects of a frame (grid = (HW/32 row blocks, B); the blocks of one object synchronise through a counter in `workspace`).
 * fp32-grade arithmetic throughout: the token chain, softmaxes, norms and GELUs in fp32; the row-local image-side contractions either in fp32 FMA
 * (force_exact) or on the tensor cores as split-bf16 (hi/lo) products with fp32 accumulation (default; ~2^-17 relative). Weight pointers are fp32 device arrays; "T" marks arrays stored TRANSPOSED ([in][out]) because the image-side
 * contractions read them coalesced over the output column; the rest are nn.Linear layout ([out][in]). pek / peq are the load-time constant products
 * dense_pe @ Wk^T (token->image k projections of layer 0, layer 1, final) and dense_pe @ Wq^T (image->token q projections of layer 0, 1), [HW][128].
 */
typedef struct VgLinear { const float* w; const float* b; } VgLinear;
typedef struct VgMdAttn { VgLinear q, k, v, out; } VgMdAttn;
typedef struct VgMdLayer {
  VgMdAttn sa;                 /* token self-attention: all [256][256] */
  VgMdAttn t2i;                /* q [128][256]; k T[256][128]; v T[256][128]; out [256][128] */
  VgMdAttn i2t;                /* q T[256][128]; k [128][256]; v [128][256]; out T[128][256] */
  VgLinear n1, n2, n3, n4;     /* LayerNorm weight / bias [256] */
  VgLinear fc1, fc2;           /* [2048][256], [256][2048] */
} VgMdLayer;
typedef struct VgMaskDecoderWeights {
  VgMdLayer layer[2];
  VgMdAttn final_t2i;          /* like t2i */
  VgLinear norm_final;
  VgLinear up1;                /* T[256][4*64], column (dy*2+dx)*64 + co; bias [64] */
  VgLinear up_ln;              /* LayerNorm2d(64) */
  VgLinear up2;                /* T[64][4*32], column (dy*2+dx)*32 + co; bias [32] */
  VgLinear hyper[4][3];        /* [256][256], [256][256], [32][256] per mask token */
  VgLinear iou[3];             /* [256][256], [256][256], [4][256] */
  VgLinear obj[3];             /* [256][256], [256][256], [1][256] */
  const float* pek[3];
  const float* peq[2];
  /* tensor-core path of the row-local contractions: the same 12 image-side matrices in nn.Linear layout [out][in], split into bf16 hi = bf16(w) and
   * lo = bf16(w - hi): index 4*layer + {0: t2i.k, 1: t2i.v, 2: i2t.q, 3: i2t.out}, 8: final k, 9: final v, 10: up1 [(dy,dx,co)][256], 11: up2 [(dy,dx,co)][64] */
  const void* img_hi[12];
  const void* img_lo[12];
} VgMaskDecoderWeights;
typedef struct VgMaskDecoder {
  const float* pix;            /* fp32 [B][HW][256] image embedding rows per object ([HW][256] shared by all objects when pix_shared != 0) */
  int pix_shared;
  const float* dense_pe;       /* fp32 [HW][256] (get_dense_pe; only used through pek / peq, kept for validation) */
  const float* no_mask;        /* fp32 [256] dense prompt embedding */
  const float* s0;             /* fp32 [16*HW][32] high-res feature of the frame (conv_s0), shared by the objects */
  const float* s1;             /* fp32 [4*HW][64]  (conv_s1) */
  const float* tokens;         /* fp32 [B][Nt][256]: [obj_score, iou, 4 mask tokens] + sparse prompt embeddings */
  int B, Nt, fs;               /* objects, tokens per object (<= 12), feature-map side (HW = fs*fs, a multiple of 32) */
  int C, heads, mlp, depth;    /* must be 256, 8, 2048, 2 (the SAM2 decoder geometry this kernel is built for) */
  const VgMaskDecoderWeights* weights;   /* HOST pointer to the table of device pointers */
  float* masks;                /* out fp32 [B][4][(4*fs)^2] */
  float* iou;                  /* out fp32 [B][4] (sigmoid) */
  float* mask_tokens;          /* out fp32 [B][4][256] (hs[:, 2:6]) */
  float* obj;                  /* out fp32 [B] object-score logits */
  void* workspace; long long workspace_bytes;   /* >= vg_mask_decoder_workspace_bytes(B, fs) */
  int force_exact;             /* 1: fp32-FMA row-local contractions (also selected by the environment variable VG_MD_EXACT=1); 0: bf16-split tensor-core path */
} VgMaskDecoder;
long long vg_mask_decoder_workspace_bytes(int B, int fs);
int vg_mask_decoder(const VgMaskDecoder* d, void* stream);
/* diagnostic: `dev_buf` = device buffer of 64 uint64 that following launches fill with the %globaltimer of block (0,0) at kernel start, after each
 * grid barrier and at the end ([63] = number of barriers); NULL disables. */
void vg_mask_decoder_set_timeline(void* dev_buf);
/* ------------------------------------------------------------------------------------------------ GPU pre-processing of the SAM2 frames
 * vg_resize_u8_pil: PIL.Image.resize(BILINEAR) on uint8 HWC frames, BIT-EXACT (what torchvision's resize(to_pil_image(x)) does inside
 * ResizeLongestSide.apply_image, VideoGLaMM/model/segment_anything/utils/transforms.py:26-36): horizontal pass src [n][h][w][c] -> tmp [n][h][ow][c],
 * vertical pass -> dst [n][oh][ow][c], both with PIL's fixed-point coefficients. kh / kv: int32 [ow][ksize_h] / [oh][ksize_v] coefficients (22 fractional
 * bits), bh / bv: int32 [ow][2] / [oh][2] = (first input index, tap count) — PIL's precompute_coeffs + normalize_coeffs_8bpc, evaluated by the caller
 * (videoglamm_b200/preprocess.py:pil_bilinear_coeffs). All pointers device memory.
 * vg_sam_normalize_squash: src uint8 [n][h][w][3] -> dst fp32 [n][3][S][S] = F.interpolate((x - mean) / std, (S, S), mode="bilinear") —
 * sam_preprocess(model_type="sam2"), VideoGLaMM/utils/sam_transforms.py:54-58. mean3 / std3 are HOST pointers to 3 floats. */
int vg_resize_u8_pil(const unsigned char* src, unsigned char* tmp, unsigned char* dst, int n, int h, int w, int c, int oh, int ow, const int* kh, const int* bh,
                     int ksize_h, const int* kv, const int* bv, int ksize_v, void* stream);
int vg_sam_normalize_squash(const unsigned char* src, float* dst, int n, int h, int w, int S, const float* mean3, const float* std3, void* stream);
/* GPU pre-processing of the encoder inputs (EncPreprocessor_VideoGPTPlus.preprocess, VideoGLaMM/utils/enc_preprocessors.py:108-166):
 * vg_resize_u8_cv: cv2.resize(x, (ow, oh)) = INTER_LINEAR on uint8 HWC frames, BIT-EXACT (VideoTrainProcessor.frames2tensor,
 * model/videogpt_plus/model/internvideo/utils.py:120-131). xi / yi: int32 [ow] / [oh] left / upper source index, xc / yc: int32 [ow][2] / [oh][2]
 * 11-bit coefficients (videoglamm_b200/preprocess.py:cv_linear_tables). The PIL BICUBIC resize of the CLIP branch is vg_resize_u8_pil with
 * bicubic coefficient tables (preprocess.py:pil_coeffs).
 * vg_u8_normalize_chw: src uint8 [n][h][w][3], crop [top, top+oh) x [left, left+ow) -> dst fp32 [n][3][oh][ow]. mode 0: float32(x * (1/255)), then
 * (v - mean) / std in float32 (HF CLIPImageProcessor rescale + normalize); mode 1: (x / 255 - mean) / std in float64, rounded once
 * (VideoTrainProcessor.normalize). mean3 / std3 are HOST pointers to 3 doubles. */
int vg_resize_u8_cv(const unsigned char* src, unsigned char* dst, int n, int h, int w, int c, int oh, int ow, const int* xi, const int* xc, const int* yi,
                    const int* yc, void* stream);
int vg_u8_normalize_chw(const unsigned char* src, float* dst, int n, int h, int w, int oh, int ow, int top, int left, int mode, const double* mean3,
                        const double* std3, void* stream);
/* argmax over each row of a [rows][cols] fp32/bf16 matrix -> int64 index (first maximal index, like torch.argmax) */
int vg_argmax(const void* x, int x_dtype, long long ldx, long long* out, int rows, int cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VGLAMM_H_ */
