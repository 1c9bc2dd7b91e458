/*
 * lsvs_b200.h — C ABI of the B200-native chunk-encoder / feature-alignment hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The reference
 * (ruppelb/Large-Scale-ViT-SLAM) is pure PyTorch, so "what its FFI would bind" are the module
 * forwards / free functions on the hot path; each entry point cites the reference
 * interface it replaces.  The Python mirror of those interfaces (large-scale-vit-slam_b200/aligned_vggt)
 * binds this header through ctypes (large-scale-vit-slam_b200/lsvs_b200/native.py); INTEGRATION.md shows the
 * reference-side stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - tensors are dense row-major; bf16 is the raw 16-bit pattern (uint16_t);
 *   - every call returns 0 on success, a negative LSVS_E* code otherwise; lsvs_last_error() gives text;
 *   - calls only enqueue work on `stream` (no host sync) unless stated.
 */
#ifndef LSVS_B200_H_
#define LSVS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSVS_OK 0
#define LSVS_EINVAL (-1)   /* bad argument / shape (the reference raises AssertionError / ValueError) */
#define LSVS_ECUDA (-2)    /* CUDA runtime / driver error */
#define LSVS_EUNSUPPORTED (-3)

typedef uint16_t lsvs_bf16;

/* ---- library ------------------------------------------------------------------------------- */
const char* lsvs_last_error(void);
int lsvs_version(void);
/* number of kernels this library has launched since load (monotonic; for bench.py's gpu_launches) */
unsigned long long lsvs_launch_count(void);

/* per-kernel-class CUDA-event profiler (used by bench.py for the roofline object; adds two event records per
 * launch while enabled).  Classes: 0 tcgen05 GEMM, 1 tcgen05 attention with < 2048 keys (frame / DINO / head passes),
 * 2 LayerNorm/cast, 3 fp32 tail (decode, camera head), 4 Sim(3) apply, 5 tcgen05 attention with >= 2048 keys (global pass). */
#define LSVS_PROF_NCAT 6
int lsvs_profile_enable(int on);
int lsvs_profile_read(double* ms, double* flops, double* bytes, long long* launches);

/* ---- Sim(3) apply (memory-bound) -------------------------------------------------------------
 * replaces apply_sim3_alignment_on_point_maps  aligned_vggt/utils/alignment.py:491-526
 *          and the inlined copies              aligned_vggt/models/featureAligned_vggt.py:198-207,
 *                                              aligned_vggt/models/poseAligned_wrapped_vggt.py:180-187
 * out[b,n,:] = T[b,:3,:3] * (s[b] * pts[b,n,:]) + T[b,:3,3]      pts/out: (B, n_points, 3) fp32
 * T: (B,4,4) fp32 row-major, s: (B) fp32.  out may alias pts.  24 B/point of HBM traffic. */
int lsvs_sim3_apply_points(const float* pts, const float* T, const float* s, float* out,
                           int batch, long long n_points, void* stream);
/* replaces `depth *= chunk_scale.view(B,1,1,1,1)`  featureAligned_vggt.py:171, alignment.py:487,
 * pointAligned_wrapped_vggt.py:137-138.   out[b,i] = s[b] * x[b,i];  8 B/pixel. */
int lsvs_scale_rows(const float* x, const float* s, float* out, int batch, long long n_per_batch, void* stream);
/* replaces apply_sim3_alignment_on_c2w alignment.py:558-594 : out[b,f] = T[b] @ [R | s[b]*t]  (B,S,4,4) */
int lsvs_sim3_apply_c2w(const float* poses, const float* T, const float* s, float* out, int batch, int frames, void* stream);
/* replaces apply_sim3_alignment_on_w2c alignment.py:528-556 : extr (B,S,rows,4), rows in {3,4} -> (B,S,4,4) */
int lsvs_sim3_apply_w2c(const float* extr, int rows, const float* T, const float* s, float* out, int batch, int frames, void* stream);

/* ---- bf16 tensor-core GEMM (tcgen05 / TMEM / TMA) ----------------------------------------------
 * C[M,N] = A[M,K] . W[N,K]^T with fp32 accumulation; A, W bf16 K-major (nn.Linear weight layout).
 * replaces the cuBLASLt calls behind UPSTREAM Attention.qkv/proj and Mlp.fc1/fc2,
 * alignment_head.py:242 (project_in) and cross_attention.py:55-57,76 (q/k/v/proj), fused with the
 * elementwise work that follows them in the reference (bias, GELU, LayerScale + residual add,
 * q_norm/k_norm LayerNorm, RoPE).  K % 64 == 0, N % 128 == 0. */
enum { LSVS_EPI_BIAS_BF16 = 0,      /* out bf16 [M,ldo]  = acc + bias                                      */
       LSVS_EPI_BIAS_GELU_BF16 = 1, /* out bf16          = gelu_erf(acc + bias)                            */
       LSVS_EPI_BIAS_F32 = 2,       /* out fp32          = acc + bias                                      */
       LSVS_EPI_RESID_F32 = 3,      /* resid fp32 [M,ldr] += gamma * (acc + bias); out2 (optional) = resid */
       LSVS_EPI_HEADNORM64_BF16 = 4,  /* out bf16: per 64-wide head LayerNorm (+RoPE) on q/k columns, bias only on the rest */
       LSVS_EPI_HEADNORM128_BF16 = 5,
       LSVS_EPI_CONV_BF16 = 6 };    /* internal to lsvs_conv2d_nhwc_bf16: bias + residuals + ReLU + border mask on a padded grid */
enum { LSVS_ROPE_NONE = 0, LSVS_ROPE_2D = 1, LSVS_ROPE_1D = 2 };

typedef struct lsvs_gemm_epilogue {
  const float* bias;  void* out;  int ldo;
  const float* gamma; float* resid; int ldr; float* out2; int ld2;
  const float* qn_w; const float* qn_b; const float* kn_w; const float* kn_b;
  int n_q_cols; int n_k_cols; float ln_eps;
  int rope_mode; const float* rope_tab;   /* [pos][n_freq][2] = (cos, sin) */
  int tokens_per_frame; int n_special; int grid_w;   /* LSVS_ROPE_2D: row -> (y,x) */
  const int* pos_ids; int pos_period;                /* LSVS_ROPE_1D: pos_ids[row % pos_period] */
} lsvs_gemm_epilogue;

int lsvs_gemm_bf16(const lsvs_bf16* A, int lda, const lsvs_bf16* W, int ldw, int M, int N, int K, int epilogue_kind,
                   const lsvs_gemm_epilogue* epilogue, void* stream);
/* Determinism: results are bit-reproducible from run to run except for the residual GEMMs (LSVS_EPI_RESID_F32) of SHORT chunks
 * (fewer than 37 output tiles, i.e. about 4-8 frames of 518x154), where K is sliced over idle CTA pairs and the slices are added
 * into the fp32 residual by L2 atomics in arrival order.  LSVS_DETERMINISTIC=1 in the environment disables the slicing.
 * M <= 128 rows (UPSTREAM CameraHead trunk on the S frame tokens of a chunk): a weight-streaming kernel with swapped operands
 * whose K slices are summed in slice order over a thread-block cluster — bit-reproducible, HBM-bound (N*K*2 bytes per call). */
/* cos/sin table used by the RoPE epilogues: tab[p][j] = (cos, sin)(p * base^(-2j/(2*n_freq))), p < n_pos.
 * (rope.py:46-58; fp32 angles)  `tab` holds n_pos*n_freq*2 floats. */
int lsvs_rope_table(float* tab, int n_pos, int n_freq, float base, void* stream);

/* ---- fused attention (tcgen05 / TMEM / TMA, online softmax) --------------------------------------
 * O = softmax(Q K^T * scale) V per (batch, head); no mask (the reference never masks: SDPA without mask in
 * UPSTREAM Attention; the all-true mask of cross_attention.py:66-67 is a no-op).
 * q/k/v/o: bf16, row = token, head h occupies columns [h*head_dim, (h+1)*head_dim); batch b owns rows
 * [b*Lq,(b+1)*Lq) of q/o and [b*Lk,(b+1)*Lk) of k/v.  head_dim in {64,128}.  q,k,v may be column slices of
 * one fused qkv buffer (pass the slice pointers and the common row stride). */
int lsvs_attention_bf16(const lsvs_bf16* q, int ldq, const lsvs_bf16* k, int ldk, const lsvs_bf16* v, int ldv,
                        lsvs_bf16* o, int ldo, int batches, int heads, int head_dim, int Lq, int Lk, float scale,
                        void* stream);

/* ---- fp32-class operators behind lsvs_engine_config::precision (csrc/precise.cu), exported for parity tests ----------------
 * An fp32 GEMM operand x is represented by two bf16 terms hi = bf16(x), lo = bf16(x - hi).  Activations are laid out as "split
 * rows" [hi | lo | hi] (three sections of `cols` columns), weights as [hi | hi | lo]; lsvs_gemm_bf16 over the 3x longer K axis
 * then returns A W^T to ~2^-16 relative with fp32 accumulation.
 *   lsvs_layernorm_split   LayerNorm of (rows, D) fp32 (D in 512/1024/2048; w, b nullable) -> split rows, row stride ld_out >= 3D
 *   lsvs_cast_split        (rows, cols) fp32 -> split rows; gelu != 0: through the exact-erf GELU first
 *   lsvs_headnorm_rope_f32 per-head LayerNorm (+ 2-D / 1-D RoPE as in lsvs_gemm_epilogue) in place on the fp32 columns
 *                          [col0, col0 + n_heads*head_dim) of buf (rows, ld)
 *   lsvs_attention_f32     O = softmax(Q K^T scale) V with fp32 q/k/v (layout as lsvs_attention_bf16) on the CUDA cores; o: split
 *                          rows with sections `section` elements apart (the A operand of the projection GEMM) */
int lsvs_layernorm_split(const float* x, long long ld_in, const float* w, const float* b, float eps, lsvs_bf16* out, long long ld_out,
                         long long rows, int D, void* stream);
int lsvs_cast_split(const float* x, long long ld_in, lsvs_bf16* out, long long ld_out, long long rows, int cols, int gelu, void* stream);
int lsvs_headnorm_rope_f32(float* buf, long long ld, long long rows, int col0, int n_heads, int head_dim, const float* w, const float* b,
                           float eps, int rope_mode, const float* rope_tab, int tokens_per_frame, int n_special, int grid_w,
                           const int* pos_ids, int pos_period, void* stream);
int lsvs_attention_f32(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv, lsvs_bf16* o,
                       long long ldo, long long section, int batches, int heads, int head_dim, int Lq, int Lk, float scale, void* stream);

/* ---- training of the alignment head (reference: only the head trains, its blocks run under
 * torch.utils.checkpoint, alignment_head.py:361,385,498,527).  The autograd graph is torch's; its heavy nodes are these kernels
 * (lsvs_b200/train.py): the tensor-core GEMM above for every Linear forward / input-gradient / weight-gradient product, and
 *   lsvs_attention_f32_train     forward with fp32 output o (rows, ldo) and lse (batches, heads, Lq) = log2-sum-exp of the scaled scores
 *   lsvs_attention_f32_backward  dq, dk, dv from q, k, v, o, d_o (o and d_o share ldo) and lse; d_buf: batches*heads*Lq floats scratch */
int lsvs_attention_f32_train(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv, float* o,
                             long long ldo, float* lse, int batches, int heads, int head_dim, int Lq, int Lk, float scale, void* stream);
int lsvs_attention_f32_backward(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv,
                                const float* o, const float* d_o, long long ldo, const float* lse, float* d_buf, float* dq,
                                long long lddq, float* dk, long long lddk, float* dv, long long lddv, int batches, int heads,
                                int head_dim, int Lq, int Lk, float scale, void* stream);

/* ---- engine: packed weights + kernel sequencing for the module forwards ---------------------------
 * One engine per process / GPU.  Parameters are pushed by state_dict name (the names are the reference's
 * checkpoint contract) from DEVICE fp32 pointers; the engine keeps its own copies (bf16 for the
 * tensor-core GEMM weights, fp32 otherwise), so the caller's tensors may be freed or moved afterwards. */
typedef struct lsvs_engine lsvs_engine;
typedef struct lsvs_engine_config {
  int embed_dim, num_heads, patch_size, num_register_tokens; /* VGGT-1B geometry: 1024, 16, 14, 4          */
  int depth, dino_depth;                                     /* alternating-attention pairs, DINOv2 blocks */
  int head_depth_aa, num_memory_tokens;                      /* alignment head: 4, 8                       */
  int with_alignment_head, with_camera_head;
  float rope_base;                                           /* 100                                        */
  int precision;  /* 0: bf16 operands / fp32 accumulate everywhere (the reference's bf16-mixed inference, run_model.py:472).
                   * 1: the alignment-head blocks, project_in and the camera-head trunk run fp32-class (every GEMM operand split
                   *    into two bf16 terms = 16 significant bits, fp32 LayerNorm / RoPE / attention): what the reference computes
                   *    in fp32 (camera head with autocast disabled, featureAligned_vggt.py:103-104; decode alignment_head.py:340).
                   * 2: every block of the path fp32-class (3x the GEMM work, attention on the CUDA cores): verification mode that
                   *    shows the bf16 pipeline differs from the fp32 reference by rounding only.  Fixed at lsvs_engine_create. */
} lsvs_engine_config;

int lsvs_engine_create(const lsvs_engine_config* cfg, lsvs_engine** out);
int lsvs_engine_destroy(lsvs_engine* e);
/* rows/cols: 2-D view of weight matrices (rows = out features, cols = in features, conv kernels flattened);
 * 0,0 for everything else. */
int lsvs_engine_set_param(lsvs_engine* e, const char* name, const float* data, long long numel, int rows, int cols, void* stream);
/* resolve all names into per-block weight tables; fails with the first missing / mis-shaped key */
int lsvs_engine_finalize(lsvs_engine* e, void* stream);
/* DINOv2 position embedding already interpolated to the (gh, gw) patch grid: (1 + gh*gw, 1024) fp32
 * (row 0 = class token).  Weight preprocessing, done once per image size by the host side. */
int lsvs_engine_set_pos_embed(lsvs_engine* e, const float* pos, int gh, int gw, void* stream);

/* replaces UPSTREAM vggt Aggregator.forward (call site featureAligned_vggt.py:78; poseAligned_wrapped_vggt.py:62;
 * pointAligned_wrapped_vggt.py:60).  images (B,S,3,H,W) fp32 in [0,1].  For each requested layer id the
 * concatenated [frame | global] block outputs are written to taps[i]: (B,S,P,2048) fp32, P = 5 + (H/14)(W/14). */
int lsvs_aggregator_forward(lsvs_engine* e, const float* images, int B, int S, int H, int W, float* const* taps,
                            const int* tap_layers, int n_taps, void* stream);
/* replaces AlignmentHead.forward  aligned_vggt/heads/alignment_head.py:224-345 (eval path).
 * tokens (B,S,P,2048) fp32; overlap_in (B,T,P+1,1024) fp32 or NULL (first chunk); memory_in (B,8,512) or NULL.
 * out: chunk_sim3 (B,1,8), frame_se3 (B,S-1,7), memory_out (B,8,512), overlap_out (B,1+next_overlap,P+1,1024).
 * An engine created with num_memory_tokens = 0 (alignment_head.py:211,468,504) ignores memory_in / memory_out (may be NULL). */
int lsvs_alignment_head_forward(lsvs_engine* e, const float* tokens, int B, int S, int P, int H, int W, int next_overlap,
                                const float* overlap_in, int T, const float* memory_in, float* chunk_sim3,
                                float* frame_se3, float* memory_out, float* overlap_out, void* stream);
/* the same with the tokens already in bf16 — the form in which the chunk scheduler ships a chunk's last-layer tokens to the
 * alignment rank (54 MB instead of 108 MB per 32-frame chunk); they are the first GEMM's operand as they stand (precision 0 only). */
int lsvs_alignment_head_forward_bf16(lsvs_engine* e, const lsvs_bf16* tokens, int B, int S, int P, int H, int W, int next_overlap,
                                     const float* overlap_in, int T, const float* memory_in, float* chunk_sim3,
                                     float* frame_se3, float* memory_out, float* overlap_out, void* stream);
/* The head split at its only context-free boundary (project_in + token_norm + the first frame block need neither the
 * previous chunk's overlap tokens nor its memory).  lsvs_alignment_head_prefix runs that part — on the rank that encoded the chunk —
 * and writes the fp32 token stream prefix_out (B,S,P+1,1024); lsvs_alignment_head_resume continues from it on the alignment rank
 * (first temporal block ... decode).  prefix + resume produce bit for bit what lsvs_alignment_head_forward does. */
int lsvs_alignment_head_prefix(lsvs_engine* e, const float* tokens, int B, int S, int P, int H, int W, float* prefix_out, void* stream);
int lsvs_alignment_head_resume(lsvs_engine* e, const float* prefix, int B, int S, int P, int H, int W, int next_overlap,
                               const float* overlap_in, int T, const float* memory_in, float* chunk_sim3, float* frame_se3,
                               float* memory_out, float* overlap_out, void* stream);
/* the fp32 decode stage alone: AlignmentHead._decode_alignments alignment_head.py:427-540 (+ GatedUpdate,
 * gated_update.py:43-78).  align_tokens (B,S,1024) fp32 = the processed per-frame alignment tokens. */
int lsvs_alignment_decode_forward(lsvs_engine* e, const float* align_tokens, int B, int S, const float* memory_in,
                                  float* chunk_sim3, float* frame_se3, float* memory_out, void* stream);
/* replaces UPSTREAM vggt CameraHead.forward (call site featureAligned_vggt.py:106): tokens_last (B,S,P,2048)
 * -> pose_enc (B,S,9) of the last refinement iteration; pose_enc_iters (optional, NULL to skip): (num_iterations,B,S,9), the
 * activated encoding after every iteration (UPSTREAM returns that list; the reference reads only [-1], :109). */
int lsvs_camera_head_forward(lsvs_engine* e, const float* tokens_last, int B, int S, int P, int num_iterations,
                             float* pose_enc, float* pose_enc_iters, void* stream);
/* ---- DPT dense-prediction heads (depth_head / point_head) --------------------------------------------
 * replaces UPSTREAM vggt/heads/dpt_head.py DPTHead.forward (construction featureAligned_vggt.py:28-29, calls :166-168 and
 * :183-185).  prefix: "depth_head." or "point_head." (state_dict prefix of the parameters pushed
 * with lsvs_engine_set_param).  Weight layouts expected by lsvs_engine_set_param for these prefixes (rows, cols):
 *   Conv2d k x k        (oc, ic, k, k)  ->  rows = oc, cols = k*k*ic ordered (ky, kx, ic)
 *   ConvTranspose2d k=s (ic, oc, k, k)  ->  rows = k*k*oc ordered (ky, kx, oc), cols = ic
 *   scratch.output_conv2.0: 32 output channels zero-padded to 64 rows (bias to 64); output_conv2.2: (od, 32) fp32.
 * taps: 4 device pointers to the tapped Aggregator layers 4/11/17/23, each (frames, P, 2048) fp32, P = 5 + (H/14)*(W/14).
 * pred (frames, H', W', output_dim-1) fp32, conf (frames, H', W') fp32 with H' = 14*(H/14), W' = 14*(W/14);
 * activation: 0 = exp (depth), 1 = inv_log (points); confidence 1 + exp.  frames_chunk frames are processed at a time
 * (upstream default 8; results do not depend on it), <= 0: all at once.  bf16 tensor-core convolutions, fp32 accumulate. */
int lsvs_dpt_head_forward(lsvs_engine* e, const char* prefix, const float* const* taps, int frames, int P, int H, int W,
                          int output_dim, int activation, float* pred, float* conf, int frames_chunk, void* stream);
/* convolution on a zero-padded NHWC grid (building block of the DPT head, exported for parity tests):
 * x (frames*hp*wp, C) bf16 with a one-pixel zero border, w (OC, taps*C) bf16 ordered (ky, kx, c), taps = 1 or 9 (3x3, pad 1),
 * out = mask(relu?(conv + bias + res1 + res2)) on the same grid (frames*hp*wp, OC); mask_border writes zeros on the border. */
int lsvs_conv2d_nhwc_bf16(const lsvs_bf16* x, const lsvs_bf16* w, const float* bias, const lsvs_bf16* res1, const lsvs_bf16* res2,
                          lsvs_bf16* out, int frames, int hp, int wp, int C, int OC, int taps, int relu, int mask_border,
                          void* stream);
/* resampling kernels of the DPT head (exported for parity tests).  h, w: input grid; "padded" = one-pixel zero border.
 *   POS_EMBED      out (frames,h,w,C) += ratio * uv sincos embedding, aspect = W_img / H_img (DPTHead._apply_pos_embed)
 *   PAD            in (frames,h,w,C) -> out padded (frames,h+2,w+2,C)
 *   CONVT_SHUFFLE  in (frames*h*w, a*a*C) GEMM output of a ConvTranspose2d(kernel = stride = a) -> out padded (frames,a*h+2,a*w+2,C)
 *   IM2COL_S2      in (frames,h,w,C) -> out (frames*ho*wo, 9*C) patch matrix of a 3x3 / stride 2 / pad 1 convolution
 *   BILINEAR       in padded (frames,h+2,w+2,C) -> out padded (frames,a+2,b+2,C), align_corners=True (+ embedding if ratio > 0) */
enum { LSVS_DPT_POS_EMBED = 0, LSVS_DPT_PAD = 1, LSVS_DPT_CONVT_SHUFFLE = 2, LSVS_DPT_IM2COL_S2 = 3, LSVS_DPT_BILINEAR = 4 };
int lsvs_dpt_resample(int op, const lsvs_bf16* in, lsvs_bf16* out, int frames, int h, int w, int C, int a, int b, float aspect,
                      float ratio, void* stream);

/* replaces the pose / Sim(3) composition featureAligned_vggt.py:97-143 and :190-196 (and data.py:12-52,
 * geometry.py:4-37 underneath).  chunk_sim3 (B,1,8), frame_se3 (B,S-1,7), cam_enc (B,S,9) camera-head output,
 * prev_pose_enc (B,S_prev,9) aligned poses of the previous chunk or NULL.
 * out: pose_enc_out (B,S,9), point_T (B,4,4) transform for the (scaled) point maps, scale_out (B). */
int lsvs_pose_chain(const float* chunk_sim3, const float* frame_se3, const float* cam_enc, const float* prev_pose_enc,
                    int S_prev, int overlap, int B, int S, int H, int W, float* pose_enc_out, float* point_T,
                    float* scale_out, void* stream);
/* the same composition when the caller passes ground-truth poses (sample modes chunk_gt / two_chunks, run_model.py:332,
 * training_metrics.py:645).  gt_poses (B,S,gt_rows,4) world-to-camera, gt_rows 3 (padded with [0 0 0 1] like
 * poseAligned_wrapped_vggt.py:86-87) or 4.  gt_mode is a bit set:
 *   LSVS_GT_MEAN   with a previous chunk, gt_poses[:,0] replaces the averaged overlap transform
 *                  (featureAligned_vggt.py:123-124, poseAligned_wrapped_vggt.py:108-109);
 *   LSVS_GT_SCALE  S > 1: the camera translations (and scale_out, which the caller applies to depth / point maps) are multiplied by
 *                  |sum x.y / sum x.x| over the re-based predicted positions x and the first-frame-centred gt positions y
 *                  (poseAligned_wrapped_vggt.py:84-104 with scale_lse_solver, alignment.py:113-129; a CPU numpy round trip there). */
#define LSVS_GT_MEAN 1
#define LSVS_GT_SCALE 2
int lsvs_pose_chain_gt(const float* chunk_sim3, const float* frame_se3, const float* cam_enc, const float* prev_pose_enc,
                       int S_prev, int overlap, int B, int S, int H, int W, const float* gt_poses, int gt_rows, int gt_mode,
                       float* pose_enc_out, float* point_T, float* scale_out, void* stream);
/* replaces pointAligned_wrapped_vggt.py:113-122: pose_enc (B,S,9) -> extrinsics -> apply_sim3_alignment_on_w2c (alignment.py:528)
 * -> pose_enc (B,S,9), FoV entries passed through the reference's tan/atan round trip. */
int lsvs_pose_enc_apply_sim3(const float* pose_enc, const float* T, const float* s, float* out, int B, int S, int H, int W, void* stream);

/* ---- IRLS weighted Umeyama Sim(3) (point-aligned baseline) ---------------------------------------
 * replaces irls_sim3_umeyama / weighted_umeyama_sim3  aligned_vggt/models/pointAligned_wrapped_vggt.py:159-305.
 * src, dst: (n_points,3) fp32; conf_src, conf_dst: (n_points) fp32.  Finds (R (3,3), t (3), s ()) with
 * dst ~ s R src + t: w0 = sqrt(conf_src*conf_dst), points with w0 < factor*median(w0) dropped, one weighted Umeyama
 * solve, then up to max_iters Huber(delta)-reweighted solves (stops early when |dR|,|dt|,|ds| < tol).
 * No host synchronisation; *status (device int) is set to 1 if the total weight is too small (the reference raises
 * ValueError).  workspace: lsvs_irls_umeyama_workspace_bytes() bytes of device memory. */
size_t lsvs_irls_umeyama_workspace_bytes(void);
int lsvs_irls_umeyama(const float* src, const float* dst, const float* conf_src, const float* conf_dst, long long n_points,
                      float conf_threshold_factor, float delta, int max_iters, float tol, float* R, float* t, float* s,
                      int* status, void* workspace, void* stream);

/* ---- evaluation-side geometry right after the path --------------------------------
 * replaces unproject_depth_map_to_point_map  aligned_vggt/utils/geometry.py:39-75: depth (frames,H,W) fp32, extrinsics
 * (frames,3,4) world-to-camera, intrinsics (frames,3,3) -> world_points (frames,H,W,3) = R^T (K^-1 (u,v,1) d) - R^T t. */
int lsvs_unproject_depth(const float* depth, const float* extrinsics, const float* intrinsics, float* world_points, int frames,
                         int H, int W, void* stream);
/* replaces the solver of scale_align_from_depths  aligned_vggt/utils/alignment.py:244-323: per batch element the L1-optimal
 * scale a = argmin sum_i w_i |a x_i - y_i| = weighted median of y_i / x_i with weights w_i x_i, w_i = mask * conf / max(y, 0.1 *
 * mean valid depth).  depth_pred, depth_gt, mask (0/1 as fp32), conf: (B, N) fp32; scales: (B) fp32 (made positive).  Radix select
 * over the ratio bit patterns instead of the reference's sort + cumsum + searchsorted; no host synchronisation.
 * workspace: lsvs_depth_scale_align_workspace_bytes(B) bytes of device memory. */
size_t lsvs_depth_scale_align_workspace_bytes(int B);
int lsvs_depth_scale_align(const float* depth_pred, const float* depth_gt, const float* mask, const float* conf, int B, long long N,
                           float* scales, void* workspace, void* stream);

/* ---- peer mailboxes over NVLink (transport of the chunk scheduler) -------------------------
 * replaces the per-chunk hand-over of the reference's sequential loop (training/training_metrics.py:636-657: one process,
 * the previous chunk's predictions are simply passed to the next call) once chunks are dealt to several GPUs: the
 * owner's last-layer tokens + camera encodings go to the alignment rank, the decoded Sim(3) packet comes back.
 * Buffers are plain cudaMalloc memory shared between the per-GPU processes of one box through CUDA IPC; payloads move
 * with the copy engines, a sequence number published by lsvs_peer_signal / awaited by lsvs_peer_wait orders them.
 * No SM is held while waiting for a peer (unlike an NCCL send/recv kernel), no host synchronisation.
 *   lsvs_peer_alloc    zero-initialised device buffer that can be exported (synchronises the device once)
 *   lsvs_peer_export   64-byte handle to hand to another process;  lsvs_peer_open maps it there (peer access enabled lazily)
 *   lsvs_peer_put      dst (local or mapped peer memory) <- src, `bytes` bytes, in stream order
 * A mailbox starts with a two-word header: flag[0] = sequence number, flag[1] = poison.  `status` is the calling rank's health
 * word — device memory or pinned (mapped) host memory, so that the host can poll it without synchronising: 0 healthy, 1 one of
 * this rank's waits timed out, 2 a peer published poison.
 *   lsvs_peer_signal   flag[0] <- value with system-scope release, after everything earlier in the stream; if *status != 0
 *                      (status may be NULL) it publishes flag[1] <- 1 instead: a rank that lost a message never hands results
 *                      computed from stale mailboxes to its peers
 *   lsvs_peer_wait     the stream continues once (int)(flag[0] - value) >= 0; it gives up — the stream continues, nothing
 *                      hangs — when flag[1] != 0 (*status <- 2), after timeout_s seconds (*status <- 1), or at once when
 *                      *status is already non-zero */
#define LSVS_PEER_HANDLE_BYTES 64
int lsvs_peer_alloc(size_t bytes, void** ptr);
int lsvs_peer_free(void* ptr);
int lsvs_peer_export(const void* ptr, unsigned char* handle64);
int lsvs_peer_open(const unsigned char* handle64, void** ptr);
int lsvs_peer_close(void* ptr);
int lsvs_peer_put(void* dst, const void* src, size_t bytes, void* stream);
int lsvs_peer_signal(unsigned* flag, unsigned value, const unsigned* status, void* stream);
int lsvs_peer_wait(const unsigned* flag, unsigned value, unsigned* status, double timeout_s, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LSVS_B200_H_ */
