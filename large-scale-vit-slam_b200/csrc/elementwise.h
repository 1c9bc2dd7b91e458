// Internal interface of csrc/elementwise.cu.
#pragma once
#include <cuda_runtime.h>

namespace lsvs {

// row m of the logical (rows, D) view lives at physical row (m / group) * stride + offset + m % group (group 0: identity)
struct RowMap { int group = 0; int stride = 0; int offset = 0; };

int layernorm(const float* x, long long ld_in, RowMap in_map, const float* w, const float* b, float eps, void* out,
              long long ld_out, RowMap out_map, bool out_bf16, long long rows, int D, cudaStream_t st);
int cast_rows_bf16(const float* x, long long ld_in, void* out, long long ld_out, long long rows, int cols, cudaStream_t st);
// images (frames,3,H,W) fp32 in [0,1] -> (frames*(H/14)*(W/14), 640) bf16, ImageNet-normalised, taps in Conv2d order
int patch_unfold(const float* img, void* out, int frames, int H, int W, cudaStream_t st);
int dino_assemble(const float* conv, const float* cls, const float* reg, const float* pos, float* x, int frames, int Pp,
                  int n_reg, int D, cudaStream_t st);
int fill_special(const float* tok, float* x, int frames, int frames_per_seq, int P, int row_off, int n_sp, int D, cudaStream_t st);
int pack_weight_bf16(const float* w, void* out, long long rows, int k_in, int k_out, cudaStream_t st);

}  // namespace lsvs
