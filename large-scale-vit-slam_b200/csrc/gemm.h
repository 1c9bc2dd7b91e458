// Internal interface of the tcgen05 GEMM (csrc/gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include "../../include/lsvs_b200.h"

namespace lsvs {

enum { EPI_BIAS_BF16 = LSVS_EPI_BIAS_BF16, EPI_BIAS_GELU_BF16 = LSVS_EPI_BIAS_GELU_BF16, EPI_BIAS_F32 = LSVS_EPI_BIAS_F32,
       EPI_RESID_F32 = LSVS_EPI_RESID_F32, EPI_HEADNORM64_BF16 = LSVS_EPI_HEADNORM64_BF16,
       EPI_HEADNORM128_BF16 = LSVS_EPI_HEADNORM128_BF16, EPI_CONV_BF16 = LSVS_EPI_CONV_BF16 };
enum { ROPE_NONE = LSVS_ROPE_NONE, ROPE_2D = LSVS_ROPE_2D, ROPE_1D = LSVS_ROPE_1D };

// same POD as the public struct, with typed pointers
struct GemmEpilogue {
  const float* bias = nullptr;   // [N]
  void* out = nullptr;           // bf16 or fp32 [M, ldo]
  int ldo = 0;
  const float* gamma = nullptr;  // [N] LayerScale (RESID)
  float* resid = nullptr;        // fp32 [M, ldr], updated in place: resid += gamma * (acc + bias)
  int ldr = 0;
  float* out2 = nullptr;         // optional second copy of the updated residual (tapped layer output)
  int ld2 = 0;
  const float* qn_w = nullptr; const float* qn_b = nullptr;   // per-head LayerNorm of q columns
  const float* kn_w = nullptr; const float* kn_b = nullptr;   // per-head LayerNorm of k columns
  int n_q_cols = 0, n_k_cols = 0;  // [0,n_q) q heads | [n_q, n_q+n_k) k heads | rest: bias only
  float ln_eps = 1e-5f;
  int rope_mode = ROPE_NONE;
  const float2* rope_tab = nullptr;  // [pos][n_freq] (cos, sin)
  int tokens_per_frame = 0, n_special = 0, grid_w = 0;  // ROPE_2D: position from the row index
  const int* pos_ids = nullptr; int pos_period = 0;     // ROPE_1D: position = pos_ids[row % pos_period]
  // EPI_CONV_BF16 — convolution on a zero-padded NHWC grid (csrc/dpt.cu).  A rows are the pixels of the padded grid
  // [frames][conv_hp][conv_wp], A columns the conv_c input channels.  With conv_taps == 9 the reduction runs over
  // K = 9 * conv_c and k-block kb reads the A rows shifted by its tap's offset (dy * conv_wp + dx), i.e. the 3x3
  // convolution is nine shifted 1x1 convolutions accumulated in tensor memory; W is [N][(ky, kx, c)].
  // out = mask(relu?(acc + bias + res1 + res2)): border pixels of the grid are written as zeros (they are the next
  // convolution's padding).
  int conv_taps = 0, conv_c = 0, conv_hp = 0, conv_wp = 0, conv_relu = 0, conv_mask = 0;
  const void* res1 = nullptr; const void* res2 = nullptr;   // optional bf16 residual inputs, same layout / stride as out
};

// The few-row kernel (M <= 128, gemm_fewrows_tcgen05) slices K over a cluster: bit-reproducible, but a different summation order
// than the other kernels.  A caller whose results must not depend on WHICH kernel its row count selects (the DPT heads: a pass over
// 1 or 8 frames crosses M = 128 on the coarse levels) switches that kernel off for its scope; the default is on.
struct FewRowsKernel {
  bool prev;
  explicit FewRowsKernel(bool on);
  ~FewRowsKernel();
};

// Promise that the W operands of the GEMMs launched in this scope are constants (model weights packed at load time): the few-row
// kernel then streams its first W tiles before griddepcontrol.wait, under the tail of the preceding kernel.  Off by default: through
// the bare C ABI W may be the output of the previous kernel on the stream.
struct ConstWeights {
  bool prev;
  explicit ConstWeights(bool on);
  ~ConstWeights();
};

int gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int epi_kind, const GemmEpilogue& e,
              cudaStream_t st);

}  // namespace lsvs
