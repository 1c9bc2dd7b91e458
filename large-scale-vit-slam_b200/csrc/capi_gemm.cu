// C-ABI entry points for the GEMM and its helper tables.
#include "gemm.h"
#include "host_common.h"

namespace {
__global__ void rope_table_kernel(float2* tab, int n_pos, int n_freq, float base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pos * n_freq) return;
  const int p = i / n_freq, j = i % n_freq;
  // inv_freq_j = 1 / base^(2j / dim), dim = 2 * n_freq  (rope.py:46-48), angle in fp32 (rope.py:51-55)
  const float expo = (float)(2 * j) / (float)(2 * n_freq);
  const float inv_freq = 1.0f / powf(base, expo);
  const float ang = (float)p * inv_freq;
  tab[i] = make_float2(cosf(ang), sinf(ang));
}
}  // namespace

extern "C" int lsvs_rope_table(float* tab, int n_pos, int n_freq, float base, void* stream) {
  LSVS_CHECK_ARG(tab && n_pos > 0 && n_freq > 0 && base > 0, "rope_table: bad arguments");
  const int n = n_pos * n_freq;
  rope_table_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2*>(tab), n_pos, n_freq, base);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

extern "C" int lsvs_gemm_bf16(const lsvs_bf16* A, int lda, const lsvs_bf16* W, int ldw, int M, int N, int K, int kind,
                              const lsvs_gemm_epilogue* ep, void* stream) {
  LSVS_CHECK_ARG(ep, "gemm: epilogue descriptor is null");
  lsvs::GemmEpilogue e;
  e.bias = ep->bias; e.out = ep->out; e.ldo = ep->ldo; e.gamma = ep->gamma; e.resid = ep->resid; e.ldr = ep->ldr;
  e.out2 = ep->out2; e.ld2 = ep->ld2; e.qn_w = ep->qn_w; e.qn_b = ep->qn_b; e.kn_w = ep->kn_w; e.kn_b = ep->kn_b;
  e.n_q_cols = ep->n_q_cols; e.n_k_cols = ep->n_k_cols; e.ln_eps = ep->ln_eps; e.rope_mode = ep->rope_mode;
  e.rope_tab = reinterpret_cast<const float2*>(ep->rope_tab); e.tokens_per_frame = ep->tokens_per_frame;
  e.n_special = ep->n_special; e.grid_w = ep->grid_w; e.pos_ids = ep->pos_ids; e.pos_period = ep->pos_period;
  if (kind == LSVS_EPI_RESID_F32) LSVS_CHECK_ARG(e.resid && e.ldr >= N, "gemm: residual epilogue needs resid with ldr >= N");
  else LSVS_CHECK_ARG(e.out && e.ldo >= N, "gemm: epilogue needs out with ldo >= N");
  return lsvs::gemm_bf16(A, lda, W, ldw, M, N, K, kind, e, (cudaStream_t)stream);
}

#include "attention.h"
extern "C" int lsvs_attention_bf16(const lsvs_bf16* q, int ldq, const lsvs_bf16* k, int ldk, const lsvs_bf16* v, int ldv,
                                   lsvs_bf16* o, int ldo, int batches, int heads, int head_dim, int Lq, int Lk, float scale,
                                   void* stream) {
  lsvs::AttentionArgs a{q, k, v, o, ldq, ldk, ldv, ldo, batches, heads, head_dim, Lq, Lk, scale};
  return lsvs::attention_fwd(a, (cudaStream_t)stream);
}

#include "precise.h"
// ---- fp32-class operators (csrc/precise.cu), exported for parity tests of the precision modes
extern "C" int lsvs_layernorm_split(const float* x, long long ld_in, const float* w, const float* b, float eps, lsvs_bf16* out,
                                    long long ld_out, long long rows, int D, void* stream) {
  return lsvs::layernorm_split(x, ld_in, w, b, eps, out, ld_out, rows, D, (cudaStream_t)stream);
}
extern "C" int lsvs_cast_split(const float* x, long long ld_in, lsvs_bf16* out, long long ld_out, long long rows, int cols, int gelu,
                               void* stream) {
  return lsvs::cast_split(x, ld_in, out, ld_out, rows, cols, gelu != 0, (cudaStream_t)stream);
}
extern "C" int lsvs_headnorm_rope_f32(float* buf, long long ld, long long rows, int col0, int n_heads, int head_dim, const float* w,
                                      const float* b, float eps, int rope_mode, const float* rope_tab, int tokens_per_frame,
                                      int n_special, int grid_w, const int* pos_ids, int pos_period, void* stream) {
  return lsvs::headnorm_rope_f32(buf, ld, rows, col0, n_heads, head_dim, w, b, eps, rope_mode, reinterpret_cast<const float2*>(rope_tab),
                                 tokens_per_frame, n_special, grid_w, pos_ids, pos_period, (cudaStream_t)stream);
}
extern "C" int lsvs_attention_f32(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv,
                                  lsvs_bf16* o, long long ldo, long long section, int batches, int heads, int head_dim, int Lq, int Lk,
                                  float scale, void* stream) {
  lsvs::AttentionF32Args a{q, k, v, o, ldq, ldk, ldv, ldo, section, batches, heads, head_dim, Lq, Lk, scale};
  return lsvs::attention_f32(a, (cudaStream_t)stream);
}
extern "C" int lsvs_attention_f32_train(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv, float* o,
                                        long long ldo, float* lse, int batches, int heads, int head_dim, int Lq, int Lk, float scale,
                                        void* stream) {
  LSVS_CHECK_ARG(o && lse, "attention_f32_train: null output");
  lsvs::AttentionF32Args a{q, k, v, nullptr, ldq, ldk, ldv, 0, 0, batches, heads, head_dim, Lq, Lk, scale};
  a.o32 = o; a.ldo32 = ldo; a.lse = lse;
  return lsvs::attention_f32(a, (cudaStream_t)stream);
}
extern "C" int lsvs_attention_f32_backward(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv,
                                           const float* o, const float* d_o, long long ldo, const float* lse, float* d_buf, float* dq,
                                           long long lddq, float* dk, long long lddk, float* dv, long long lddv, int batches, int heads,
                                           int head_dim, int Lq, int Lk, float scale, void* stream) {
  lsvs::AttentionF32BwdArgs a{q, k, v, o, d_o, lse, d_buf, dq, dk, dv, ldq, ldk, ldv, ldo, lddq, lddk, lddv, batches, heads, head_dim, Lq, Lk, scale};
  return lsvs::attention_f32_backward(a, (cudaStream_t)stream);
}

#ifdef LSVS_MEASURE   // measurement builds only: the product ABI has no switch that changes results
namespace lsvs { extern int g_gemm_mode; }
extern "C" int lsvs_debug_gemm_mode(int mode) {
  lsvs::g_gemm_mode = mode;
  return LSVS_OK;
}
#endif
