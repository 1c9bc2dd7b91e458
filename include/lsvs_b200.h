/*
 * lsvs_b200.h — C ABI of the B200-native chunk-encoder / feature-alignment hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The reference
 * (ruppelb/Large-Scale-ViT-SLAM) is pure PyTorch, so "what its FFI would bind" are the module
 * forwards / free functions on the hot path; each entry point cites the reference
 * interface it replaces.  The Python mirror of those interfaces (large-scale-vit-slam_b200/aligned_vggt)
 * binds this header through ctypes (large-scale-vit-slam_b200/_native.py); INTEGRATION.md shows the
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
       LSVS_EPI_HEADNORM128_BF16 = 5 };
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

#ifdef __cplusplus
}
#endif
#endif /* LSVS_B200_H_ */
