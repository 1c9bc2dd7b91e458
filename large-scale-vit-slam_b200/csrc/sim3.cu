// Sim(3) application kernels (memory-bound, fp32).
//   points : reference apply_sim3_alignment_on_point_maps, aligned_vggt/utils/alignment.py:491-526
//   depth  : featureAligned_vggt.py:171
//   poses  : apply_sim3_alignment_on_c2w / _on_w2c, alignment.py:528-594
// The reference materialises scaled copy + homogeneous cat + an expanded (N,4,4) transform + bmm
// (>= 200 B/point); here one pass reads 12 B and writes 12 B per point.
#include <cstdlib>

#include "host_common.h"

namespace {

struct Sim3 {  // 3x3 rotation pre-multiplied by the scale, plus translation
  float r[9];
  float t[3];
};

__device__ __forceinline__ Sim3 load_sim3(const float* __restrict__ T, const float* __restrict__ s, int b) {
  Sim3 m;
  const float* Tb = T + (size_t)b * 16;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) m.r[i * 3 + j] = __ldg(Tb + i * 4 + j);
    m.t[i] = __ldg(Tb + i * 4 + 3);
  }
  (void)s;  // the scale is applied to the point first (reference rounding order), see xform()
  return m;
}

__device__ __forceinline__ void xform(const Sim3& m, float sc, float x, float y, float z, float& ox, float& oy, float& oz) {
  // reference: p*s, then row-dot with [R|t] over the homogeneous vector (x,y,z,1) accumulated left to right
  x *= sc; y *= sc; z *= sc;
  ox = fmaf(m.r[2], z, fmaf(m.r[1], y, m.r[0] * x)) + m.t[0];
  oy = fmaf(m.r[5], z, fmaf(m.r[4], y, m.r[3] * x)) + m.t[1];
  oz = fmaf(m.r[8], z, fmaf(m.r[7], y, m.r[6] * x)) + m.t[2];
}

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Vector path: each thread owns groups of 4 consecutive points = 12 floats = 3 x 16 B.
// grid.y = batch, grid.x * block strides over the groups of one batch element.
__global__ void __launch_bounds__(256) sim3_points_vec4(const float* __restrict__ pts, const float* __restrict__ T,
                                                        const float* __restrict__ s, float* __restrict__ out,
                                                        long long n_points) {
  const int b = blockIdx.y;
  const Sim3 m = load_sim3(T, s, b);
  const float sc = __ldg(s + b);
  const long long groups = n_points >> 2;
  const float4* in4 = reinterpret_cast<const float4*>(pts + (size_t)b * n_points * 3);
  float4* out4 = reinterpret_cast<float4*>(out + (size_t)b * n_points * 3);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const float4 a = ld_stream(in4 + 3 * g), c = ld_stream(in4 + 3 * g + 1), d = ld_stream(in4 + 3 * g + 2);
    float4 oa, oc, od;
    xform(m, sc, a.x, a.y, a.z, oa.x, oa.y, oa.z);
    xform(m, sc, a.w, c.x, c.y, oa.w, oc.x, oc.y);
    xform(m, sc, c.z, c.w, d.x, oc.z, oc.w, od.x);
    xform(m, sc, d.y, d.z, d.w, od.y, od.z, od.w);
    st_stream(out4 + 3 * g, oa);
    st_stream(out4 + 3 * g + 1, oc);
    st_stream(out4 + 3 * g + 2, od);
  }
}

// Coalesced path: a warp moves 128 points (384 floats) at a time with three fully coalesced 16-byte accesses per lane in each
// direction; the float3 -> per-lane regrouping goes through a 1.5 KB shared-memory slab (a lane's 4 points are 48 contiguous
// bytes: quarter-warps hit all 32 banks once, no conflicts).  The per-lane strided variant above touches every 32-byte sector
// from two different instructions.
__global__ void __launch_bounds__(256) sim3_points_coalesced(const float* __restrict__ pts, const float* __restrict__ T,
                                                             const float* __restrict__ s, float* __restrict__ out,
                                                             long long n_points) {
  __shared__ float4 slab[8][96];
  const int b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const Sim3 m = load_sim3(T, s, b);
  const float sc = __ldg(s + b);
  const long long chunks = n_points >> 7;  // 128 points = 96 float4
  const float4* in4 = reinterpret_cast<const float4*>(pts + (size_t)b * n_points * 3);
  float4* out4 = reinterpret_cast<float4*>(out + (size_t)b * n_points * 3);
  const long long stride = (long long)gridDim.x * 8;
  float4* sl = slab[w];
  for (long long c = (long long)blockIdx.x * 8 + w; c < chunks; c += stride) {
    const float4* src = in4 + c * 96;
    const float4 v0 = ld_stream(src + lane), v1 = ld_stream(src + 32 + lane), v2 = ld_stream(src + 64 + lane);
    sl[lane] = v0; sl[32 + lane] = v1; sl[64 + lane] = v2;
    __syncwarp();
    const float4 a = sl[3 * lane], cc = sl[3 * lane + 1], d = sl[3 * lane + 2];
    float4 oa, oc, od;
    xform(m, sc, a.x, a.y, a.z, oa.x, oa.y, oa.z);
    xform(m, sc, a.w, cc.x, cc.y, oa.w, oc.x, oc.y);
    xform(m, sc, cc.z, cc.w, d.x, oc.z, oc.w, od.x);
    xform(m, sc, d.y, d.z, d.w, od.y, od.z, od.w);
    __syncwarp();
    sl[3 * lane] = oa; sl[3 * lane + 1] = oc; sl[3 * lane + 2] = od;
    __syncwarp();
    float4* dst = out4 + c * 96;
    st_stream(dst + lane, sl[lane]); st_stream(dst + 32 + lane, sl[32 + lane]); st_stream(dst + 64 + lane, sl[64 + lane]);
    __syncwarp();
  }
}

// Scalar path for n_points % 4 != 0 (batch rows then lose 16-byte alignment) and for the tail.
__global__ void __launch_bounds__(256) sim3_points_scalar(const float* __restrict__ pts, const float* __restrict__ T,
                                                          const float* __restrict__ s, float* __restrict__ out,
                                                          long long n_points, long long first) {
  const int b = blockIdx.y;
  const Sim3 m = load_sim3(T, s, b);
  const float sc = __ldg(s + b);
  const float* in = pts + (size_t)b * n_points * 3;
  float* o = out + (size_t)b * n_points * 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_points; i += stride) {
    float ox, oy, oz;
    xform(m, sc, in[3 * i], in[3 * i + 1], in[3 * i + 2], ox, oy, oz);
    o[3 * i] = ox; o[3 * i + 1] = oy; o[3 * i + 2] = oz;
  }
}

__global__ void __launch_bounds__(256) scale_rows_kernel(const float* __restrict__ x, const float* __restrict__ s,
                                                         float* __restrict__ out, long long n, int vec_ok) {
  const int b = blockIdx.y;
  const float sc = __ldg(s + b);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float* in = x + (size_t)b * n;
  float* o = out + (size_t)b * n;
  long long done = 0;
  if (vec_ok) {
    const long long n4 = n >> 2;
    const float4* in4 = reinterpret_cast<const float4*>(in);
    float4* o4 = reinterpret_cast<float4*>(o);
    for (long long i = tid; i < n4; i += stride) {
      float4 v = ld_stream(in4 + i);
      v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
      st_stream(o4 + i, v);
    }
    done = n4 << 2;
  }
  for (long long i = done + tid; i < n; i += stride) o[i] = in[i] * sc;
}

// One thread per (batch, frame) 4x4 pose.  mode 0: c2w in -> T @ [R | s t];  mode 1: w2c (rows x 4) in ->
// inverse, scale, T@, inverse back (closed-form SE(3) inverses, as the reference does).
__device__ __forceinline__ void se3_inverse(const float* m, float* o) {  // m,o: 4x4 row-major, bottom row assumed 0 0 0 1
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) o[i * 4 + j] = m[j * 4 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i * 4 + 3] = -(o[i * 4 + 0] * m[3] + o[i * 4 + 1] * m[7] + o[i * 4 + 2] * m[11]);
  o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
}

__global__ void sim3_poses_kernel(const float* __restrict__ in, int rows, const float* __restrict__ T,
                                  const float* __restrict__ s, float* __restrict__ out, int batch, int frames, int mode) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * frames) return;
  const int b = idx / frames;
  float m[16], c2w[16], r[16];
  const float* src = in + (size_t)idx * rows * 4;
#pragma unroll
  for (int i = 0; i < 16; ++i) m[i] = (i < rows * 4) ? src[i] : (i == 15 ? 1.f : 0.f);
  if (mode == 1) se3_inverse(m, c2w);
  else {
#pragma unroll
    for (int i = 0; i < 16; ++i) c2w[i] = m[i];
  }
  const float sc = s[b];
  c2w[3] *= sc; c2w[7] *= sc; c2w[11] *= sc;
  const float* Tb = T + (size_t)b * 16;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fmaf(Tb[i * 4 + k], c2w[k * 4 + j], acc);
      r[i * 4 + j] = acc;
    }
  float* dst = out + (size_t)idx * 16;
  if (mode == 1) {
    float inv[16];
    se3_inverse(r, inv);
#pragma unroll
    for (int i = 0; i < 16; ++i) dst[i] = inv[i];
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) dst[i] = r[i];
  }
}

int grid_for(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)lsvs::num_sms() * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" int lsvs_sim3_apply_points(const float* pts, const float* T, const float* s, float* out, int batch,
                                      long long n_points, void* stream) {
  LSVS_CHECK_ARG(pts && T && s && out, "sim3_apply_points: null pointer");
  LSVS_CHECK_ARG(batch > 0 && batch <= 65535 && n_points >= 0, "sim3_apply_points: Inputs must have matching batch dimension (batch=%d n=%lld)", batch, n_points);
  if (n_points == 0) return LSVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  lsvs::ProfScope prof(lsvs::PROF_SIM3, st, 0, 24.0 * batch * (double)n_points);
  const bool aligned = ((n_points & 3) == 0 || batch == 1) && ((uintptr_t)pts % 16 == 0) && ((uintptr_t)out % 16 == 0);
  long long tail_first = 0;
  if (aligned && (n_points >> 7) > 0) {  // measured: 3.7 -> 4.8 TB/s at 2.55 M points, 4.6 -> 5.7 TB/s at 17 M (tools/ub_sim3.py)
    dim3 grid(grid_for((n_points >> 7) * 32, 256), batch);
    sim3_points_coalesced<<<grid, 256, 0, st>>>(pts, T, s, out, n_points);
    LSVS_LAUNCH_CHECK();
    tail_first = (n_points >> 7) << 7;
  } else if (aligned && (n_points >> 2) > 0) {
    dim3 grid(grid_for(n_points >> 2, 256), batch);
    sim3_points_vec4<<<grid, 256, 0, st>>>(pts, T, s, out, n_points);
    LSVS_LAUNCH_CHECK();
    tail_first = (n_points >> 2) << 2;
  }
  if (tail_first < n_points) {
    dim3 grid(grid_for(n_points - tail_first, 256), batch);
    sim3_points_scalar<<<grid, 256, 0, st>>>(pts, T, s, out, n_points, tail_first);
    LSVS_LAUNCH_CHECK();
  }
  return LSVS_OK;
}

extern "C" int lsvs_scale_rows(const float* x, const float* s, float* out, int batch, long long n, void* stream) {
  LSVS_CHECK_ARG(x && s && out, "scale_rows: null pointer");
  LSVS_CHECK_ARG(batch > 0 && batch <= 65535 && n >= 0, "scale_rows: bad shape");
  if (n == 0) return LSVS_OK;
  const int vec_ok = ((n & 3) == 0 || batch == 1) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0);
  dim3 grid(grid_for(vec_ok ? (n >> 2) + 1 : n, 256), batch);
  scale_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, s, out, n, vec_ok);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

extern "C" int lsvs_sim3_apply_c2w(const float* poses, const float* T, const float* s, float* out, int batch, int frames, void* stream) {
  LSVS_CHECK_ARG(poses && T && s && out && batch > 0 && frames > 0, "sim3_apply_c2w: Inputs must have matching batch dimension");
  const int n = batch * frames;
  sim3_poses_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(poses, 4, T, s, out, batch, frames, 0);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

extern "C" int lsvs_sim3_apply_w2c(const float* extr, int rows, const float* T, const float* s, float* out, int batch, int frames, void* stream) {
  LSVS_CHECK_ARG(extr && T && s && out && batch > 0 && frames > 0, "sim3_apply_w2c: Inputs must have matching batch dimension");
  LSVS_CHECK_ARG(rows == 3 || rows == 4, "sim3_apply_w2c: extrinsics must be (B,S,3,4) or (B,S,4,4)");
  const int n = batch * frames;
  sim3_poses_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(extr, rows, T, s, out, batch, frames, 1);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
