// Internal interface of csrc/precise.cu: fp32-class block operators (split-bf16 GEMM operands, fp32 attention).
#pragma once
#include <cuda_runtime.h>

#include "gemm.h"

namespace lsvs {
// LayerNorm of (rows, D) fp32 -> split bf16 rows [hi | lo | hi] (3*D wide, row stride ld_out)
int layernorm_split(const float* x, long long ld_in, const float* w, const float* b, float eps, void* out, long long ld_out,
                    long long rows, int D, cudaStream_t st);
// (rows, cols) fp32 -> split bf16 (rows, 3*cols); gelu: through the exact-erf GELU first
int cast_split(const float* x, long long ld_in, void* out, long long ld_out, long long rows, int cols, bool gelu, cudaStream_t st);
// act: 0 none, 1 GELU, 2 SiLU
int cast_split_act(const float* x, long long ld_in, void* out, long long ld_out, long long rows, int cols, int act, cudaStream_t st);
// weights (rows, k_in) fp32 -> (rows, 3*k_pad) bf16 = [hi | hi | lo]
int pack_weight_split(const float* w, void* out, long long rows, int k_in, int k_pad, cudaStream_t st);
// images -> fp32 im2col of the patch convolution: (frames*gh*gw, 640)
int patch_unfold_f32(const float* img, float* out, int frames, int H, int W, cudaStream_t st);
// per-head LayerNorm (+ RoPE) in place on fp32 columns [col0, col0 + n_heads*hd) of buf
int headnorm_rope_f32(float* buf, long long ld, long long rows, int col0, int n_heads, int hd, const float* w, const float* b, float eps,
                      int rope_mode, const float2* tab, int tpf, int nsp, int gw, const int* pos_ids, int period, cudaStream_t st);
struct AttentionF32Args {
  const float* q; const float* k; const float* v;   // fp32, row = token, head h at columns [h*head_dim, (h+1)*head_dim)
  void* o;                                            // bf16 split rows [hi | lo | hi], sections `section` elements apart
  long long ldq, ldk, ldv, ldo, section;
  int batches, heads, head_dim, Lq, Lk;
  float scale;
  // training forward (optional): fp32 output rows (stride ldo32) and per (batch, head, query) log2-sum-exp of the scaled scores
  float* o32 = nullptr; long long ldo32 = 0; float* lse = nullptr;
};
int attention_f32(const AttentionF32Args& a, cudaStream_t st);
// backward of attention_f32: q/k/v/o/d_o fp32 (o and d_o share the stride ldo), lse from the forward; d_buf: scratch of
// batches*heads*Lq floats (dO . O per row); dq/dk/dv fp32 with their own strides
struct AttentionF32BwdArgs {
  const float* q; const float* k; const float* v; const float* o; const float* d_o; const float* lse;
  float* d_buf; float* dq; float* dk; float* dv;
  long long ldq, ldk, ldv, ldo, lddq, lddk, lddv;
  int batches, heads, head_dim, Lq, Lk;
  float scale;
};
int attention_f32_backward(const AttentionF32BwdArgs& a, cudaStream_t st);
}  // namespace lsvs
