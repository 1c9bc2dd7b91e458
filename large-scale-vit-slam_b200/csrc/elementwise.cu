// Memory-bound kernels around the GEMM / attention pair: LayerNorm (fp32 in, bf16 or fp32 out, optional row
// re-mapping), casts, patch unfold (+ImageNet normalisation), DINOv2 token assembly, special-token fill.
// All are one-warp-per-row or one-block-per-patch-row, 16-byte vectorised, coalesced.
//
// Replaces ATen LayerNorm / cat / expand / conv-unfold kernels the reference path launches around every Linear
// (norm1/norm2, token_norm alignment_head.py:247, DINOv2 prepare_tokens, Aggregator token concat).
#include "elementwise.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lsvs {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ long long map_row(const RowMap& r, long long m) {
  return r.group > 0 ? (m / r.group) * (long long)r.stride + r.offset + (m % r.group) : m;
}

// One warp per row; the row (D = 128*NV floats) lives in registers: lane holds NV float4 at stride 32.
template <int NV, bool OUT_BF16>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long ld_in, RowMap in_map,
                                                        const float* __restrict__ w, const float* __restrict__ b, float eps,
                                                        void* __restrict__ out, long long ld_out, RowMap out_map, long long rows) {
  pdl_launch_dependents();
  pdl_wait();  // programmatic dependent launch: the grid may be resident before its predecessor has finished
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= rows) return;
  constexpr int D = NV * 128;
  const float4* src = reinterpret_cast<const float4*>(x + map_row(in_map, m) * ld_in);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = src[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
    q += (a * a + c * c) + (d * d + e * e);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  const long long orow = map_row(out_map, m);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f), h = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w) g = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
    if (b) h = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + h.x;
    o.y = (v[i].y - mean) * rstd * g.y + h.y;
    o.z = (v[i].z - mean) * rstd * g.z + h.z;
    o.w = (v[i].w - mean) * rstd * g.w + h.w;
    if constexpr (OUT_BF16) {
      uint2 p = make_uint2(ptx::pack_bf16(o.x, o.y), ptx::pack_bf16(o.z, o.w));
      reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + orow * ld_out)[lane + 32 * i] = p;
    } else {
      reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + orow * ld_out)[lane + 32 * i] = o;
    }
  }
}

__global__ void __launch_bounds__(256) cast_rows_kernel(const float* __restrict__ x, long long ld_in, __nv_bfloat16* __restrict__ out,
                                                        long long ld_out, long long rows, int cols4) {
  const long long total = rows * cols4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols4;
    const int c = (int)(i % cols4);
    const float4 v = reinterpret_cast<const float4*>(x + r * ld_in)[c];
    reinterpret_cast<uint2*>(out + r * ld_out)[c] = make_uint2(ptx::pack_bf16(v.x, v.y), ptx::pack_bf16(v.z, v.w));
  }
}

// Patch unfold for the 14x14/14 conv: one block per (frame, patch-row).  Stage the 3 x 14 image rows (normalised,
// bf16) in shared memory with coalesced reads, then emit one 640-wide bf16 row per patch (588 taps in (c,ky,kx)
// order = Conv2d weight layout, zero padded to a multiple of 64 for the GEMM's K blocking).
constexpr int PATCH = 14;
constexpr int KPATCH = 3 * PATCH * PATCH;  // 588
constexpr int KPAD = 640;

__global__ void __launch_bounds__(256) patch_unfold_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int H,
                                                           int W, int gh, int gw) {
  extern __shared__ __nv_bfloat16 tile[];  // [3][14][W]
  const int f = blockIdx.x / gh, gy = blockIdx.x % gh;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  const int n_in = 3 * PATCH * W;
  for (int i = threadIdx.x; i < n_in; i += blockDim.x) {
    const int c = i / (PATCH * W), rem = i % (PATCH * W);
    const int ky = rem / W, xx = rem % W;
    const float v = img[(((size_t)f * 3 + c) * H + (gy * PATCH + ky)) * W + xx];
    tile[i] = __float2bfloat16_rn((v - mean[c]) / stdv[c]);
  }
  __syncthreads();
  __nv_bfloat16* dst = out + ((size_t)f * gh * gw + (size_t)gy * gw) * KPAD;
  const int n_out = gw * KPAD;
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) {
    const int gx = i / KPAD, k = i % KPAD;
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (k < KPATCH) {
      const int c = k / (PATCH * PATCH), rem = k % (PATCH * PATCH);
      const int ky = rem / PATCH, kx = rem % PATCH;
      v = tile[(c * PATCH + ky) * W + gx * PATCH + kx];
    }
    dst[i] = v;
  }
}

// DINOv2 prepare_tokens: row t of frame f = cls+pos[0] | register tokens (no pos) | conv(patch)+pos[1+p].
__global__ void __launch_bounds__(256) dino_assemble_kernel(const float* __restrict__ conv, const float* __restrict__ cls,
                                                            const float* __restrict__ reg, const float* __restrict__ pos,
                                                            float* __restrict__ x, int frames, int Pp, int n_reg, int D4) {
  const int P = 1 + n_reg + Pp;
  const long long total = (long long)frames * P * D4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % D4);
    const long long row = i / D4;
    const int t = (int)(row % P);
    const long long f = row / P;
    float4 v;
    if (t == 0) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(cls) + c), p = __ldg(reinterpret_cast<const float4*>(pos) + c);
      v = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    } else if (t <= n_reg) {
      v = __ldg(reinterpret_cast<const float4*>(reg) + (size_t)(t - 1) * D4 + c);
    } else {
      const int pp = t - 1 - n_reg;
      const float4 a = reinterpret_cast<const float4*>(conv)[((size_t)f * Pp + pp) * D4 + c];
      const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + (size_t)(1 + pp) * D4 + c);
      v = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    }
    reinterpret_cast<float4*>(x)[i] = v;
  }
}

// Special tokens (1,2,n_sp,D): variant 0 for the first frame of each sequence, 1 for the others
// (slice_expand_and_flatten, alignment_head.py:543-568).  Writes rows [row_off, row_off+n_sp) of every frame.
__global__ void __launch_bounds__(256) fill_special_kernel(const float* __restrict__ tok, float* __restrict__ x, int frames,
                                                           int frames_per_seq, int P, int row_off, int n_sp, int D4) {
  const long long total = (long long)frames * n_sp * D4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % D4);
    const int t = (int)((i / D4) % n_sp);
    const long long f = i / ((long long)D4 * n_sp);
    const int variant = (f % frames_per_seq) == 0 ? 0 : 1;
    const float4 v = __ldg(reinterpret_cast<const float4*>(tok) + ((size_t)variant * n_sp + t) * D4 + c);
    reinterpret_cast<float4*>(x)[((size_t)f * P + row_off + t) * D4 + c] = v;
  }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n,
                                                          int k_in, int k_out) {
  // (rows, k_in) fp32 -> (rows, k_out) bf16 with zero padding of the extra columns (weight packing)
  const long long total = n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / k_out;
    const int c = (int)(i % k_out);
    out[i] = __float2bfloat16_rn(c < k_in ? x[r * k_in + c] : 0.f);
  }
}

int blocks_for(long long items, int threads = 256) {
  long long b = (items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int layernorm(const float* x, long long ld_in, RowMap in_map, const float* w, const float* b, float eps, void* out,
              long long ld_out, RowMap out_map, bool out_bf16, long long rows, int D, cudaStream_t st) {
  LSVS_CHECK_ARG(x && out && rows >= 0, "layernorm: null pointer");
  LSVS_CHECK_ARG(D == 512 || D == 1024 || D == 2048, "layernorm: D=%d unsupported (512/1024/2048)", D);
  if (rows == 0) return LSVS_OK;
  ProfScope prof(PROF_ELEMENTWISE, st, 0, (double)rows * D * (4 + (out_bf16 ? 2 : 4)));
  const int warps = 8;
  const unsigned grid = (unsigned)((rows + warps - 1) / warps);
#define LSVS_LN(NV)                                                                                                      \
  if (out_bf16) LSVS_CUDA(launch_pdl(layernorm_kernel<NV, true>, dim3(grid), dim3(warps * 32), 0, st, x, ld_in, in_map, w, b, eps, out, ld_out, out_map, rows)); \
  else LSVS_CUDA(launch_pdl(layernorm_kernel<NV, false>, dim3(grid), dim3(warps * 32), 0, st, x, ld_in, in_map, w, b, eps, out, ld_out, out_map, rows));
  if (D == 512) { LSVS_LN(4) } else if (D == 1024) { LSVS_LN(8) } else { LSVS_LN(16) }
#undef LSVS_LN
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int cast_rows_bf16(const float* x, long long ld_in, void* out, long long ld_out, long long rows, int cols, cudaStream_t st) {
  LSVS_CHECK_ARG(x && out && cols % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0, "cast_rows: bad arguments");
  if (rows == 0) return LSVS_OK;
  ProfScope prof(PROF_ELEMENTWISE, st, 0, (double)rows * cols * 6);
  cast_rows_kernel<<<blocks_for(rows * (cols / 4)), 256, 0, st>>>(x, ld_in, reinterpret_cast<__nv_bfloat16*>(out), ld_out, rows, cols / 4);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int patch_unfold(const float* img, void* out, int frames, int H, int W, cudaStream_t st) {
  LSVS_CHECK_ARG(img && out && frames > 0 && H % PATCH == 0 && W % PATCH == 0, "patch_unfold: image size must be a multiple of 14");
  const int gh = H / PATCH, gw = W / PATCH;
  const size_t smem = (size_t)3 * PATCH * W * sizeof(__nv_bfloat16);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    LSVS_CUDA(cudaFuncSetAttribute(patch_unfold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  patch_unfold_kernel<<<frames * gh, 256, smem, st>>>(img, reinterpret_cast<__nv_bfloat16*>(out), H, W, gh, gw);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int dino_assemble(const float* conv, const float* cls, const float* reg, const float* pos, float* x, int frames, int Pp,
                  int n_reg, int D, cudaStream_t st) {
  dino_assemble_kernel<<<blocks_for((long long)frames * (1 + n_reg + Pp) * (D / 4)), 256, 0, st>>>(conv, cls, reg, pos, x, frames, Pp, n_reg, D / 4);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int fill_special(const float* tok, float* x, int frames, int frames_per_seq, int P, int row_off, int n_sp, int D, cudaStream_t st) {
  fill_special_kernel<<<blocks_for((long long)frames * n_sp * (D / 4)), 256, 0, st>>>(tok, x, frames, frames_per_seq, P, row_off, n_sp, D / 4);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int pack_weight_bf16(const float* w, void* out, long long rows, int k_in, int k_out, cudaStream_t st) {
  f32_to_bf16_kernel<<<blocks_for(rows * k_out), 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16*>(out), rows * k_out, k_in, k_out);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

}  // namespace lsvs
