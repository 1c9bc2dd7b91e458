// Internal interface of the tcgen05 flash attention (csrc/attention.cu).
#pragma once
#include <cuda_runtime.h>

namespace lsvs {

struct AttentionArgs {
  const void* q; const void* k; const void* v;  // bf16, row = token, head h at columns [h*head_dim, (h+1)*head_dim)
  void* o;                                       // bf16 (batches*Lq, ldo)
  int ldq, ldk, ldv, ldo;                        // row strides in elements
  int batches, heads, head_dim;
  int Lq, Lk;                                    // tokens per sequence; batch b owns rows [b*L, (b+1)*L)
  float scale;                                   // softmax scale (head_dim^-0.5)
};

int attention_fwd(const AttentionArgs& a, cudaStream_t st);

}  // namespace lsvs
