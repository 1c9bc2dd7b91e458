// DPT dense-prediction head (UPSTREAM vggt/heads/dpt_head.py; call sites featureAligned_vggt.py:28-29,166,183):
// resampling / data-movement kernels around the tensor-core convolutions.  The convolutions themselves are GEMMs
// (csrc/gemm.cu): 1x1 and transposed convolutions directly, 3x3 convolutions as nine row-shifted GEMM k-slabs over a
// zero-padded NHWC grid (EPI_CONV_BF16).  Everything here is HBM-bound element-wise work: 16-byte accesses, one
// thread per (pixel, 8 channels).
#include "dpt.h"
#include "gemm.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lsvs {
namespace {

constexpr int TPB = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) { f[2 * j] = ptx::bf16_lo(w[j]); f[2 * j + 1] = ptx::bf16_hi(w[j]); }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(ptx::pack_bf16(f[0], f[1]), ptx::pack_bf16(f[2], f[3]), ptx::pack_bf16(f[4], f[5]), ptx::pack_bf16(f[6], f[7]));
}

// Eight consecutive channels of one pixel: 16 bytes of bf16 (the benchmarked path) or 32 bytes of fp32 (precision modes,
// where the DPT heads keep fp32 activations like the reference: featureAligned_vggt.py:103, autocast disabled).
template <class T> struct V8;
template <> struct V8<__nv_bfloat16> {
  uint4 u;
  __device__ static V8 zero() { V8 v; v.u = make_uint4(0u, 0u, 0u, 0u); return v; }
  __device__ static V8 load(const __nv_bfloat16* p) { V8 v; v.u = *reinterpret_cast<const uint4*>(p); return v; }
  __device__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = u; }
  __device__ void to_float(float* f) const { unpack8(u, f); }
  __device__ static V8 from_float(const float* f) { V8 v; v.u = pack8(f); return v; }
};
template <> struct V8<float> {
  float4 a, b;
  __device__ static V8 zero() { V8 v; v.a = v.b = make_float4(0.f, 0.f, 0.f, 0.f); return v; }
  __device__ static V8 load(const float* p) { V8 v; v.a = *reinterpret_cast<const float4*>(p); v.b = *reinterpret_cast<const float4*>(p + 4); return v; }
  __device__ void store(float* p) const { *reinterpret_cast<float4*>(p) = a; *reinterpret_cast<float4*>(p + 4) = b; }
  __device__ void to_float(float* f) const { f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w; }
  __device__ static V8 from_float(const float* f) { V8 v; v.a = make_float4(f[0], f[1], f[2], f[3]); v.b = make_float4(f[4], f[5], f[6], f[7]); return v; }
};

// torch.linspace(-span*(n-1)/n, +span*(n-1)/n, n)[i] (symmetric evaluation like ATen)
__device__ __forceinline__ float uv_coord(int i, int n, float span) {
  const float end = span * (float)(n - 1) / (float)n, start = -end;
  if (n == 1) return start;
  const float step = (end - start) / (float)(n - 1);
  return i < n / 2 ? start + step * (float)i : end - step * (float)(n - 1 - i);
}
// channel c of position_grid_to_embed: [sin(u w_k), cos(u w_k), sin(v w_k), cos(v w_k)], w_k = 100^(-k / (C/4))
__device__ __forceinline__ float uv_embed(int c, int C, float u, float v) {
  const int half = C >> 1, quarter = C >> 2;
  const float coord = c < half ? u : v;
  const int cc = c < half ? c : c - half;
  const int k = cc < quarter ? cc : cc - quarter;
  const float omega = exp2f(-6.643856189774724f * (float)k / (float)quarter);  // log2(100)
  const float a = coord * omega;
  return cc < quarter ? sinf(a) : cosf(a);
}
__device__ __forceinline__ void uv_spans(float aspect, float& sx, float& sy) {
  const float diag = sqrtf(aspect * aspect + 1.0f);
  sx = aspect / diag;
  sy = 1.0f / diag;
}

// Launch geometry of the resampling kernels: blockIdx.x = row of the (padded) output grid (frame * rows_per_frame + y),
// blockIdx.y * TPB + threadIdx.x = (x, 8-channel group) within the row — 32-bit index math only.
// uv position embedding: either from per-axis tables (U [w][C/2], V [h][C/2], already scaled by ratio; engine path) or
// evaluated in place (tables == nullptr).
__device__ __forceinline__ void add_embed8(float* f, int c8, int C, int xx, int yy, int w, int h, float aspect, float ratio,
                                           const float* U, const float* V) {
  const int c0 = c8 * 8, half = C >> 1;
  if (U) {
    const float* t = c0 < half ? U + (size_t)xx * half + c0 : V + (size_t)yy * half + (c0 - half);
    const float4 a = *reinterpret_cast<const float4*>(t), b = *reinterpret_cast<const float4*>(t + 4);
    f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w; f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
  } else {
    float sx, sy;
    uv_spans(aspect, sx, sy);
    const float u = uv_coord(xx, w, sx), v = uv_coord(yy, h, sy);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += ratio * uv_embed(c0 + j, C, u, v);
  }
}

__global__ void uv_table_kernel(float* U, float* V, int h, int w, int C, float aspect, float ratio) {
  const int half = C >> 1;
  const int i = blockIdx.x * TPB + threadIdx.x;
  if (i >= (w + h) * half) return;
  float sx, sy;
  uv_spans(aspect, sx, sy);
  if (i < w * half) {
    const int xx = i / half, c = i - xx * half;
    U[i] = ratio * uv_embed(c, C, uv_coord(xx, w, sx), 0.f);
  } else {
    const int k = i - w * half, yy = k / half, c = k - yy * half;
    V[k] = ratio * uv_embed(half + c, C, 0.f, uv_coord(yy, h, sy));
  }
}

template <class T>
__global__ void pos_embed_kernel(T* x, int h, int w, int C8, float aspect, float ratio, const float* U, const float* V) {
  const int idx = blockIdx.y * TPB + threadIdx.x;
  if (idx >= w * C8) return;
  const int xx = idx / C8, c8 = idx - xx * C8;
  const int yy = blockIdx.x % h;
  T* p = x + ((size_t)blockIdx.x * w * C8 + idx) * 8;
  float f[8];
  V8<T>::load(p).to_float(f);
  add_embed8(f, c8, C8 * 8, xx, yy, w, h, aspect, ratio, U, V);
  V8<T>::from_float(f).store(p);
}

template <class T>
__global__ void pad_kernel(const T* in, T* out, int h, int w, int C8) {
  const int wp = w + 2, hp = h + 2;
  const int idx = blockIdx.y * TPB + threadIdx.x;
  if (idx >= wp * C8) return;
  const int xp = idx / C8, c8 = idx - xp * C8;
  const int f = blockIdx.x / hp, yp = blockIdx.x - f * hp;
  V8<T> v = V8<T>::zero();
  if (xp >= 1 && xp <= w && yp >= 1 && yp <= h) v = V8<T>::load(in + ((((size_t)f * h + (yp - 1)) * w + (xp - 1)) * C8 + c8) * 8);
  v.store(out + ((size_t)blockIdx.x * wp * C8 + idx) * 8);
}

template <class T>
__global__ void convt_shuffle_kernel(const T* in, const float* bias, T* out, int h, int w, int C8, int k) {
  const int wp = k * w + 2, hp = k * h + 2;
  const int idx = blockIdx.y * TPB + threadIdx.x;
  if (idx >= wp * C8) return;
  const int xp = idx / C8, c8 = idx - xp * C8;
  const int f = blockIdx.x / hp, yp = blockIdx.x - f * hp;
  V8<T> v = V8<T>::zero();
  if (xp >= 1 && xp <= k * w && yp >= 1 && yp <= k * h) {
    const int oy = yp - 1, ox = xp - 1;
    const int y = oy / k, ii = oy - y * k, x = ox / k, jj = ox - x * k;
    v = V8<T>::load(in + (((((size_t)f * h + y) * w + x) * (k * k) + ii * k + jj) * C8 + c8) * 8);
    if (bias) {
      float t[8];
      v.to_float(t);
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] += bias[c8 * 8 + j];
      v = V8<T>::from_float(t);
    }
  }
  v.store(out + ((size_t)blockIdx.x * wp * C8 + idx) * 8);
}

// blockIdx.x = output pixel row (f, oy), idx = (ox, tap, c8)
template <class T>
__global__ void im2col_s2_kernel(const T* in, T* out, int h, int w, int ho, int wo, int C8) {
  const int idx = blockIdx.y * TPB + threadIdx.x;
  if (idx >= wo * 9 * C8) return;
  const int c8 = idx % C8, r = idx / C8, tap = r % 9, ox = r / 9;
  const int f = blockIdx.x / ho, oy = blockIdx.x - f * ho;
  const int y = 2 * oy - 1 + tap / 3, x = 2 * ox - 1 + tap % 3;
  V8<T> v = V8<T>::zero();
  if (y >= 0 && y < h && x >= 0 && x < w) v = V8<T>::load(in + ((((size_t)f * h + y) * w + x) * C8 + c8) * 8);
  v.store(out + ((size_t)blockIdx.x * wo * 9 * C8 + idx) * 8);
}

template <class T>
__global__ void bilinear_kernel(const T* in, T* out, int hi, int wi, int ho, int wo, int C8, float aspect, float ratio,
                                const float* U, const float* V) {
  const int wpo = wo + 2, hpo = ho + 2, wpi = wi + 2, hpi = hi + 2;
  const int idx = blockIdx.y * TPB + threadIdx.x;
  if (idx >= wpo * C8) return;
  const int xp = idx / C8, c8 = idx - xp * C8;
  const int f = blockIdx.x / hpo, yp = blockIdx.x - f * hpo;
  T* dst = out + ((size_t)blockIdx.x * wpo * C8 + idx) * 8;
  if (xp < 1 || xp > wo || yp < 1 || yp > ho) { V8<T>::zero().store(dst); return; }
  const int oy = yp - 1, ox = xp - 1;
  // align_corners=True (ATen area_pixel_compute_scale / source index)
  const float sy = ho > 1 ? (float)(hi - 1) / (float)(ho - 1) * (float)oy : 0.f;
  const float sx = wo > 1 ? (float)(wi - 1) / (float)(wo - 1) * (float)ox : 0.f;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < hi - 1 ? 1 : 0), x1 = x0 + (x0 < wi - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const T* r0 = in + (((size_t)f * hpi + y0 + 1) * wpi * C8 + c8) * 8;
  const T* r1 = in + (((size_t)f * hpi + y1 + 1) * wpi * C8 + c8) * 8;
  float a[8], b[8], c[8], d[8], o[8];
  V8<T>::load(r0 + (size_t)(x0 + 1) * C8 * 8).to_float(a);
  V8<T>::load(r0 + (size_t)(x1 + 1) * C8 * 8).to_float(b);
  V8<T>::load(r1 + (size_t)(x0 + 1) * C8 * 8).to_float(c);
  V8<T>::load(r1 + (size_t)(x1 + 1) * C8 * 8).to_float(d);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = w00 * a[j] + w01 * b[j] + w10 * c[j] + w11 * d[j];
  if (ratio > 0.f) add_embed8(o, c8, C8 * 8, ox, oy, wo, ho, aspect, ratio, U, V);
  V8<T>::from_float(o).store(dst);
}

// one thread per pixel: 32 input channels (post-ReLU) x od fp32 weights, then activate_head
template <class T>
__global__ void final_kernel(const T* in, int ldc, const float* w, const float* b, int od, int activation, float* pred,
                             float* conf, long long total, int H, int W) {
  __shared__ float sw[4 * 32 + 4];
  if (threadIdx.x < od * 32) sw[threadIdx.x] = w[threadIdx.x];
  if (threadIdx.x < od) sw[128 + threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  const long long i = (long long)blockIdx.x * TPB + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const long long f = i / ((long long)W * H);
  const T* src = in + ((f * (H + 2) + y + 1) * (long long)(W + 2) + x + 1) * ldc;
  float acc[4] = {sw[128], sw[129], sw[130], sw[131]};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float v[8];
    V8<T>::load(src + 8 * q).to_float(v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int o = 0; o < 4; ++o) if (o < od) acc[o] += v[j] * sw[o * 32 + q * 8 + j];
    }
  }
  for (int o = 0; o < od - 1; ++o) {
    const float t = acc[o];
    pred[i * (od - 1) + o] = activation == 0 ? expf(t) : copysignf(expm1f(fabsf(t)), t);
  }
  conf[i] = 1.0f + expf(acc[od - 1]);
}

// fp32-class 3x3 convolution operand: row q of the padded grid -> [hi(tap 0..8) | lo(tap 0..8) | hi(tap 0..8)], 27 C bf16 wide,
// against weights [hi | hi | lo] (pack_weight_split of the [oc][(ky,kx,ic)] matrix): x = hi + lo to 16 bits, and
// hi w_hi + lo w_hi + hi w_lo drops only lo w_lo (2^-16 relative).  blockIdx.x = row of the padded grid, idx = (tap, c8).
__global__ void split_im2col3_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long rows, int wp, int C8) {
  const int idx = blockIdx.y * TPB + threadIdx.x;
  if (idx >= 9 * C8) return;
  const int tap = idx / C8, c8 = idx - tap * C8;
  const long long q = blockIdx.x;
  const long long src = q + (long long)(tap / 3 - 1) * wp + (tap % 3 - 1);
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (src >= 0 && src < rows) V8<float>::load(in + ((size_t)src * C8 + c8) * 8).to_float(f);
  float hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { hi[j] = __bfloat162float(__float2bfloat16_rn(f[j])); lo[j] = f[j] - hi[j]; }
  const size_t K9 = (size_t)9 * C8 * 8;
  __nv_bfloat16* o = out + (size_t)q * 3 * K9 + (size_t)idx * 8;
  const V8<__nv_bfloat16> vh = V8<__nv_bfloat16>::from_float(hi);
  vh.store(o);
  V8<__nv_bfloat16>::from_float(lo).store(o + K9);
  vh.store(o + 2 * K9);
}

// y = mask(relu?(y + r1 + r2)) on the padded grid (rows, OC) fp32: the tail of EPI_CONV_BF16 for the fp32-class path
__global__ void post_f32_kernel(float* y, const float* r1, const float* r2, int relu, int hp, int wp, int OC8) {
  const int idx = blockIdx.y * TPB + threadIdx.x;
  if (idx >= wp * OC8) return;
  const int xp = idx / OC8;
  const int yp = blockIdx.x % hp;
  float* p = y + ((size_t)blockIdx.x * wp * OC8 + idx) * 8;
  if (xp < 1 || xp > wp - 2 || yp < 1 || yp > hp - 2) { V8<float>::zero().store(p); return; }
  float f[8], t[8];
  V8<float>::load(p).to_float(f);
  if (r1) { V8<float>::load(r1 + (p - y)).to_float(t);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += t[j]; }
  if (r2) { V8<float>::load(r2 + (p - y)).to_float(t);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += t[j]; }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  V8<float>::from_float(f).store(p);
}

int blocks_for(long long total) { return (int)((total + TPB - 1) / TPB); }
dim3 row_grid(long long rows, int per_row) { return dim3((unsigned)rows, (unsigned)((per_row + TPB - 1) / TPB)); }

}  // namespace

int dpt_uv_tables(float* U, float* V, int h, int w, int C, float aspect, float ratio, cudaStream_t st) {
  LSVS_CHECK_ARG(U && V && C % 16 == 0, "dpt_uv_tables: bad arguments");
  uv_table_kernel<<<blocks_for((long long)(w + h) * (C / 2)), TPB, 0, st>>>(U, V, h, w, C, aspect, ratio);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
// `f32`: the activations are fp32 instead of bf16 (precision modes)
template <class T> const T* in_as(const void* p) { return reinterpret_cast<const T*>(p); }

int dpt_add_pos_embed(void* x, int frames, int h, int w, int C, float aspect, float ratio, const float* U, const float* V, cudaStream_t st, bool f32) {
  LSVS_CHECK_ARG(x && C % 16 == 0, "dpt_add_pos_embed: bad arguments");
  ProfScope prof(PROF_ELEMENTWISE, st, 0, 4.0 * frames * h * (double)w * C);
  const dim3 grid = row_grid((long long)frames * h, w * (C / 8));
  if (f32) pos_embed_kernel<float><<<grid, TPB, 0, st>>>(reinterpret_cast<float*>(x), h, w, C / 8, aspect, ratio, U, V);
  else pos_embed_kernel<__nv_bfloat16><<<grid, TPB, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(x), h, w, C / 8, aspect, ratio, U, V);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
int dpt_pad(const void* in, void* out, int frames, int h, int w, int C, cudaStream_t st, bool f32) {
  ProfScope prof(PROF_ELEMENTWISE, st, 0, 2.0 * frames * ((double)h * w + (double)(h + 2) * (w + 2)) * C);
  const dim3 grid = row_grid((long long)frames * (h + 2), (w + 2) * (C / 8));
  if (f32) pad_kernel<float><<<grid, TPB, 0, st>>>(in_as<float>(in), reinterpret_cast<float*>(out), h, w, C / 8);
  else pad_kernel<__nv_bfloat16><<<grid, TPB, 0, st>>>(in_as<__nv_bfloat16>(in), reinterpret_cast<__nv_bfloat16*>(out), h, w, C / 8);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
int dpt_convt_shuffle(const void* in, const float* bias, void* out, int frames, int h, int w, int C, int k, cudaStream_t st, bool f32) {
  ProfScope prof(PROF_ELEMENTWISE, st, 0, 2.0 * frames * ((double)k * k * h * w + (double)(k * h + 2) * (k * w + 2)) * C);
  const dim3 grid = row_grid((long long)frames * (k * h + 2), (k * w + 2) * (C / 8));
  if (f32) convt_shuffle_kernel<float><<<grid, TPB, 0, st>>>(in_as<float>(in), bias, reinterpret_cast<float*>(out), h, w, C / 8, k);
  else convt_shuffle_kernel<__nv_bfloat16><<<grid, TPB, 0, st>>>(in_as<__nv_bfloat16>(in), bias, reinterpret_cast<__nv_bfloat16*>(out), h, w, C / 8, k);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
int dpt_im2col_s2(const void* in, void* out, int frames, int h, int w, int C, cudaStream_t st, bool f32) {
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  ProfScope prof(PROF_ELEMENTWISE, st, 0, 2.0 * frames * ((double)h * w + 9.0 * ho * wo) * C);
  const dim3 grid = row_grid((long long)frames * ho, wo * 9 * (C / 8));
  if (f32) im2col_s2_kernel<float><<<grid, TPB, 0, st>>>(in_as<float>(in), reinterpret_cast<float*>(out), h, w, ho, wo, C / 8);
  else im2col_s2_kernel<__nv_bfloat16><<<grid, TPB, 0, st>>>(in_as<__nv_bfloat16>(in), reinterpret_cast<__nv_bfloat16*>(out), h, w, ho, wo, C / 8);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
int dpt_bilinear(const void* in, void* out, int frames, int hi, int wi, int ho, int wo, int C, float aspect, float ratio, const float* U,
                 const float* V, cudaStream_t st, bool f32) {
  ProfScope prof(PROF_ELEMENTWISE, st, 0, 2.0 * frames * ((double)(hi + 2) * (wi + 2) + (double)(ho + 2) * (wo + 2)) * C);
  const dim3 grid = row_grid((long long)frames * (ho + 2), (wo + 2) * (C / 8));
  if (f32) bilinear_kernel<float><<<grid, TPB, 0, st>>>(in_as<float>(in), reinterpret_cast<float*>(out), hi, wi, ho, wo, C / 8, aspect, ratio, U, V);
  else bilinear_kernel<__nv_bfloat16><<<grid, TPB, 0, st>>>(in_as<__nv_bfloat16>(in), reinterpret_cast<__nv_bfloat16*>(out), hi, wi, ho, wo, C / 8, aspect, ratio, U, V);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
int dpt_final(const void* in, int ldc, const float* w, const float* b, int od, int activation, float* pred, float* conf, int frames,
              int H, int W, cudaStream_t st, bool f32) {
  ProfScope prof(PROF_ELEMENTWISE, st, 0, (double)frames * H * W * (64.0 + 4.0 * od));
  LSVS_CHECK_ARG(od >= 2 && od <= 4 && ldc >= 32 && ldc % 8 == 0, "dpt_final: output_dim must be 2..4");
  const long long total = (long long)frames * H * W;
  if (f32) final_kernel<float><<<blocks_for(total), TPB, 0, st>>>(in_as<float>(in), ldc, w, b, od, activation, pred, conf, total, H, W);
  else final_kernel<__nv_bfloat16><<<blocks_for(total), TPB, 0, st>>>(in_as<__nv_bfloat16>(in), ldc, w, b, od, activation, pred, conf, total, H, W);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
int dpt_split_im2col3(const float* in, void* out, int frames, int hp, int wp, int C, cudaStream_t st) {
  LSVS_CHECK_ARG(in && out && C % 8 == 0, "dpt_split_im2col3: bad arguments");
  const long long rows = (long long)frames * hp * wp;
  ProfScope prof(PROF_ELEMENTWISE, st, 0, (double)rows * C * (9.0 * 4 + 27.0 * 2));
  split_im2col3_kernel<<<row_grid(rows, 9 * (C / 8)), TPB, 0, st>>>(in, reinterpret_cast<__nv_bfloat16*>(out), rows, wp, C / 8);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
int dpt_post_f32(float* y, const float* r1, const float* r2, int relu, int frames, int hp, int wp, int OC, cudaStream_t st) {
  LSVS_CHECK_ARG(y && OC % 8 == 0, "dpt_post_f32: bad arguments");
  ProfScope prof(PROF_ELEMENTWISE, st, 0, (double)frames * hp * wp * OC * 4.0 * (2 + (r1 != nullptr) + (r2 != nullptr)));
  post_f32_kernel<<<row_grid((long long)frames * hp, wp * (OC / 8)), TPB, 0, st>>>(y, r1, r2, relu, hp, wp, OC / 8);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

}  // namespace lsvs

// ---------------------------------------------------------------------------------------------------- C ABI
extern "C" int lsvs_conv2d_nhwc_bf16(const lsvs_bf16* x, const lsvs_bf16* w, const float* bias, const lsvs_bf16* res1,
                                     const lsvs_bf16* res2, lsvs_bf16* out, int frames, int hp, int wp, int C, int OC, int taps,
                                     int relu, int mask_border, void* stream) {
  using namespace lsvs;
  LSVS_CHECK_ARG(x && w && out && frames > 0 && hp > 2 && wp > 2, "conv2d: bad arguments");
  LSVS_CHECK_ARG(taps == 1 || taps == 9, "conv2d: taps must be 1 (1x1) or 9 (3x3, pad 1)");
  LSVS_CHECK_ARG(C % 64 == 0, "conv2d: input channels must be a multiple of 64");
  GemmEpilogue e;
  e.bias = bias; e.out = out; e.ldo = OC;
  e.conv_taps = taps == 9 ? 9 : 0; e.conv_c = C; e.conv_hp = hp; e.conv_wp = wp; e.conv_relu = relu; e.conv_mask = mask_border;
  e.res1 = res1; e.res2 = res2;
  const long long M = (long long)frames * hp * wp;
  LSVS_CHECK_ARG(M < (1ll << 31), "conv2d: too many pixels for one launch");
  return gemm_bf16(x, C, w, taps * C, (int)M, OC, taps * C, EPI_CONV_BF16, e, (cudaStream_t)stream);
}

extern "C" int lsvs_dpt_resample(int op, const lsvs_bf16* in, lsvs_bf16* out, int frames, int h, int w, int C, int a, int b, float aspect,
                                 float ratio, void* stream) {
  using namespace lsvs;
  cudaStream_t st = (cudaStream_t)stream;
  LSVS_CHECK_ARG(out && frames > 0 && h > 0 && w > 0 && C > 0 && C % 8 == 0, "dpt_resample: bad arguments");
  switch (op) {
    case LSVS_DPT_POS_EMBED: return dpt_add_pos_embed(out, frames, h, w, C, aspect, ratio, nullptr, nullptr, st);
    case LSVS_DPT_PAD: LSVS_CHECK_ARG(in, "dpt_resample: null input"); return dpt_pad(in, out, frames, h, w, C, st);
    case LSVS_DPT_CONVT_SHUFFLE: LSVS_CHECK_ARG(in && a > 0, "dpt_resample: bad stride"); return dpt_convt_shuffle(in, nullptr, out, frames, h, w, C, a, st);
    case LSVS_DPT_IM2COL_S2: LSVS_CHECK_ARG(in, "dpt_resample: null input"); return dpt_im2col_s2(in, out, frames, h, w, C, st);
    case LSVS_DPT_BILINEAR: LSVS_CHECK_ARG(in && a > 0 && b > 0, "dpt_resample: bad output size"); return dpt_bilinear(in, out, frames, h, w, a, b, C, aspect, ratio, nullptr, nullptr, st);
    default: return fail(LSVS_EINVAL, "dpt_resample: unknown op %d", op);
  }
}
