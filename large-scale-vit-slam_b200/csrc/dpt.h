// Internal interface of csrc/dpt.cu: data-movement / resampling kernels of the DPT dense-prediction head.
// All activations are NHWC, bf16 (or fp32 with the trailing `f32` flag: precision modes, csrc/engine.cu).  "Padded" tensors carry a one-pixel zero border, [frames][h+2][w+2][C], so that a 3x3
// convolution is nine row-shifted GEMMs over the flattened pixel index (gemm.h, EPI_CONV_BF16).
#pragma once
#include <cuda_runtime.h>

namespace lsvs {

// x[(f,y,x), c] += ratio * sincos_embed(uv grid)  (UPSTREAM DPTHead._apply_pos_embed); x unpadded (frames*h*w, C)
// U [w][C/2], V [h][C/2] (optional, from dpt_uv_tables, already scaled by ratio) replace the in-kernel sin/cos evaluation
int dpt_uv_tables(float* U, float* V, int h, int w, int C, float aspect, float ratio, cudaStream_t st);
int dpt_add_pos_embed(void* x, int frames, int h, int w, int C, float aspect, float ratio, const float* U, const float* V, cudaStream_t st, bool f32 = false);
// unpadded (frames,h,w,C) -> padded (frames,h+2,w+2,C)
int dpt_pad(const void* in, void* out, int frames, int h, int w, int C, cudaStream_t st, bool f32 = false);
// ConvTranspose2d(kernel = stride = k) as GEMM output [(f,y,x)][(i,j,c)] (+ bias[c]) -> padded (frames, k*h+2, k*w+2, C)
int dpt_convt_shuffle(const void* in, const float* bias, void* out, int frames, int h, int w, int C, int k, cudaStream_t st, bool f32 = false);
// 3x3 / stride 2 / pad 1 patch matrix: unpadded (frames,h,w,C) -> (frames*ho*wo, 9*C), taps in (ky,kx,c) order
int dpt_im2col_s2(const void* in, void* out, int frames, int h, int w, int C, cudaStream_t st, bool f32 = false);
// bilinear, align_corners=True: padded (frames,hi+2,wi+2,C) -> padded (frames,ho+2,wo+2,C); optionally adds the uv
// position embedding (ratio > 0) to the result
int dpt_bilinear(const void* in, void* out, int frames, int hi, int wi, int ho, int wo, int C, float aspect, float ratio,
                 const float* U, const float* V, cudaStream_t st, bool f32 = false);
// last 1x1 convolution (32 -> od channels, fp32 weights) + activate_head: padded (frames,H+2,W+2,ldc) bf16 ->
// pred (frames,H,W,od-1) fp32, conf (frames,H,W) fp32.  activation: 0 exp, 1 inv_log; conf: 1 + exp
int dpt_final(const void* in, int ldc, const float* w, const float* b, int od, int activation, float* pred, float* conf,
              int frames, int H, int W, cudaStream_t st, bool f32 = false);
// fp32-class 3x3 convolution (precision modes): padded (frames,hp,wp,C) fp32 -> split bf16 patch rows (rows, 27 C) =
// [hi taps | lo taps | hi taps] for a GEMM against [hi | hi | lo] weights; and the convolution tail on its fp32 result
int dpt_split_im2col3(const float* in, void* out, int frames, int hp, int wp, int C, cudaStream_t st);
int dpt_post_f32(float* y, const float* r1, const float* r2, int relu, int frames, int hp, int wp, int OC, cudaStream_t st);

}  // namespace lsvs
