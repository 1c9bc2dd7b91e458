// Internal interface of csrc/small_f32.cu (fp32 latency-bound tail: alignment decode, GatedUpdate, camera head).
#pragma once
#include <cuda_runtime.h>

namespace lsvs {

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2, ACT_SIGMOID = 3 };

// y[m,n] = out_act(sum_k in_act(x[m,k]) W[n,k] + b[n]); optional y = (residual ? y_old : 0) + gamma[n] * value
int linear_f32(const float* x, long long ldx, const float* W, const float* b, float* y, long long ldy, int M, int N, int K,
               int in_act, int out_act, const float* gamma, bool residual, cudaStream_t st);

struct SmallAttnArgs {
  const float* q; long long ldq; const float* k; long long ldk; const float* v; long long ldv; float* out; long long ldo;
  int B, H, hd, Nq, Nk;
  const float* qn_w; const float* qn_b; const float* kn_w; const float* kn_b;  // per-head LayerNorm (nullable)
  const int* pos_q; const int* pos_k;                                          // 1-D RoPE position ids (nullable)
  float rope_base; float scale;
};
int attn_small_f32(const SmallAttnArgs& a, cudaStream_t st);

int mean_row_norm(const float* x, int B, int S, int D, float* out, cudaStream_t st);
int memory_prepare(const float* tokens, const float* mem_param, const float* mem_in, const float* frame_init, const float* alpha,
                   const float* mean_norm, float* kv, float* directional, int B, int S, int NM, int D, cudaStream_t st);
int gu_prepare(const float* mem, const float* upd, float* inp, float* mem_scaled, int B, int NM, int D, cudaStream_t st);
int gu_gate_input(const float* deltas, const float* mem, const float* mem_scaled, float* gate_in, int rows, int D, cudaStream_t st);
int gu_finish(const float* gate_in, const float* mem, const float* gate, float* out, int rows, int D, cudaStream_t st);
int modulate(const float* normed, const float* tok, const float* mod, float* out, long long rows, int D, cudaStream_t st);
int combine_rows(const float* a, long long lda, const float* b, long long ldb, float* out, long long ldo, long long rows, int cols,
                 int relu_from, int exp_col, cudaStream_t st);

// Pose / Sim(3) composition of one chunk (featureAligned_vggt.py:97-143,190-196).  See csrc/pose.cu.
int pose_chain(const float* chunk_sim3, const float* frame_se3, const float* cam_enc, const float* prev_pose_enc, int S_prev,
               int overlap, int B, int S, int H, int W, float* pose_enc_out, float* point_T, float* scale_out, cudaStream_t st,
               const float* gt_poses = nullptr, int gt_rows = 4, int gt_mode = 0);

int pose_enc_apply_sim3(const float* enc, const float* T, const float* s, float* out, int B, int S, int H, int W, cudaStream_t st);
// IRLS weighted Umeyama (csrc/umeyama.cu); status: 0 ok, 1 = total weight too small
size_t irls_umeyama_workspace_bytes();
int irls_umeyama(const float* src, const float* dst, const float* conf_src, const float* conf_dst, long long M, float factor, float delta,
                 int max_iters, float tol, float* R, float* t, float* s, int* status, void* workspace, cudaStream_t stream);

}  // namespace lsvs
