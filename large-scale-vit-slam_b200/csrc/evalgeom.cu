// Evaluation-side geometry right after the path, fp32, HBM-bound:
//   lsvs_unproject_depth          aligned_vggt/utils/geometry.py:39-75  (unproject_depth_map_to_point_map)
//   lsvs_depth_scale_align        aligned_vggt/utils/alignment.py:244-323 (scale_align_from_depths: weighted-median L1 scale)
// The reference materialises pixel grids, rays, homogeneous copies and a full sort of S*H*W ratios; here the unprojection is one
// pass (4 B read + 12 B written per pixel) and the weighted median is a 4-pass radix select over the ratio bit patterns
// (24 B read per pixel and pass, no sort, no host synchronisation).
#include <cstdint>

#include "host_common.h"

namespace {

constexpr int TPB = 256;

// ---------------------------------------------------------------------------------------------------- unprojection
struct Cam { float kinv[9]; float r[9]; float t[3]; };  // K^-1, R^T, -R^T t  (camera-to-world)

__device__ __forceinline__ Cam load_cam(const float* __restrict__ E, const float* __restrict__ K) {
  Cam c;
  // general 3x3 inverse by cofactors (torch.inverse on the intrinsics, geometry.py:58)
  const float a = K[0], b = K[1], cc = K[2], d = K[3], e = K[4], f = K[5], g = K[6], h = K[7], i = K[8];
  const float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  const float det = a * A + b * B + cc * C;
  const float id = 1.0f / det;
  c.kinv[0] = A * id; c.kinv[1] = -(b * i - cc * h) * id; c.kinv[2] = (b * f - cc * e) * id;
  c.kinv[3] = B * id; c.kinv[4] = (a * i - cc * g) * id;  c.kinv[5] = -(a * f - cc * d) * id;
  c.kinv[6] = C * id; c.kinv[7] = -(a * h - b * g) * id;  c.kinv[8] = (a * e - b * d) * id;
  // closed_form_inverse_se3 of the world-to-camera [R|t]: R^T, -R^T t (geometry.py:69)
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) c.r[r * 3 + q] = E[q * 4 + r];
#pragma unroll
  for (int r = 0; r < 3; ++r) c.t[r] = -(c.r[r * 3] * E[3] + c.r[r * 3 + 1] * E[7] + c.r[r * 3 + 2] * E[11]);
  return c;
}

// A warp handles 128 consecutive pixels of one frame per step: each lane reads 4 depths (one 16-byte load), computes its 4
// points and stages the 12 floats in shared memory so that the float3 output leaves as three fully coalesced 16-byte stores
// per lane (same regrouping as the Sim(3) kernel, csrc/sim3.cu).  grid.y = frame; H*W % 4 == 0 on this path.
__global__ void __launch_bounds__(TPB) unproject_kernel(const float* __restrict__ depth, const float* __restrict__ extr,
                                                        const float* __restrict__ intr, float* __restrict__ out, int H, int W) {
  __shared__ Cam cam;
  __shared__ float4 slab[TPB / 32][96];
  const int f = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) cam = load_cam(extr + (size_t)f * 12, intr + (size_t)f * 9);
  __syncthreads();
  const int n = H * W, chunks = (n + 127) >> 7;
  const float* dsrc = depth + (size_t)f * n;
  float4* dst4 = reinterpret_cast<float4*>(out + (size_t)f * n * 3);
  float* sl = reinterpret_cast<float*>(slab[w]);
  for (int c = blockIdx.x * (TPB / 32) + w; c < chunks; c += gridDim.x * (TPB / 32)) {
    const int p0 = (c << 7) + 4 * lane;
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    if (p0 + 3 < n) { const float4 v = __ldg(reinterpret_cast<const float4*>(dsrc + p0)); d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
    else { for (int j = 0; j < 4; ++j) if (p0 + j < n) d[j] = __ldg(dsrc + p0 + j); }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = p0 + j, v = p / W, u = p - v * W;
      // ray = K^-1 (u, v, 1); cam = ray * depth; world = R^T cam - R^T t
      const float fu = (float)u, fv = (float)v;
      const float rx = (cam.kinv[0] * fu + cam.kinv[1] * fv + cam.kinv[2]) * d[j];
      const float ry = (cam.kinv[3] * fu + cam.kinv[4] * fv + cam.kinv[5]) * d[j];
      const float rz = (cam.kinv[6] * fu + cam.kinv[7] * fv + cam.kinv[8]) * d[j];
      sl[12 * lane + 3 * j + 0] = cam.r[0] * rx + cam.r[1] * ry + cam.r[2] * rz + cam.t[0];
      sl[12 * lane + 3 * j + 1] = cam.r[3] * rx + cam.r[4] * ry + cam.r[5] * rz + cam.t[1];
      sl[12 * lane + 3 * j + 2] = cam.r[6] * rx + cam.r[7] * ry + cam.r[8] * rz + cam.t[2];
    }
    __syncwarp();
    const long long base4 = (long long)c * 96;       // float4 index of this chunk's first output
    const long long end4 = ((long long)n * 3) >> 2;  // H*W % 4 == 0: the frame's output is a whole number of float4
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const long long i4 = base4 + 32 * k + lane;
      if (i4 < end4) dst4[i4] = slab[w][32 * k + lane];
    }
    __syncwarp();
  }
}

// pixel-per-thread variant for H*W % 4 != 0 (frames then lose 16-byte alignment)
__global__ void __launch_bounds__(TPB) unproject_scalar_kernel(const float* __restrict__ depth, const float* __restrict__ extr,
                                                               const float* __restrict__ intr, float* __restrict__ out, int H, int W) {
  __shared__ Cam cam;
  const int f = blockIdx.y;
  if (threadIdx.x == 0) cam = load_cam(extr + (size_t)f * 12, intr + (size_t)f * 9);
  __syncthreads();
  const int p = blockIdx.x * TPB + threadIdx.x;
  if (p >= H * W) return;
  const int v = p / W, u = p - v * W;
  const float d = __ldg(depth + (size_t)f * H * W + p);
  const float fu = (float)u, fv = (float)v;
  const float rx = (cam.kinv[0] * fu + cam.kinv[1] * fv + cam.kinv[2]) * d;
  const float ry = (cam.kinv[3] * fu + cam.kinv[4] * fv + cam.kinv[5]) * d;
  const float rz = (cam.kinv[6] * fu + cam.kinv[7] * fv + cam.kinv[8]) * d;
  float* o = out + ((size_t)f * H * W + p) * 3;
  o[0] = cam.r[0] * rx + cam.r[1] * ry + cam.r[2] * rz + cam.t[0];
  o[1] = cam.r[3] * rx + cam.r[4] * ry + cam.r[5] * rz + cam.t[1];
  o[2] = cam.r[6] * rx + cam.r[7] * ry + cam.r[8] * rz + cam.t[2];
}

// ---------------------------------------------------------------------------------------------------- weighted-median scale
struct ScaleState {
  double sum_valid, sum_depth;   // sum(m), sum(y * m)
  double total;                  // sum of effective weights
  double below;                  // effective weight of all ratios below the current radix prefix
  unsigned long long count_below;
  unsigned int prefix;
  float min_depth;
  float scale;
  double hist_w[256];
  unsigned long long hist_n[256];
};

__device__ __forceinline__ unsigned int key_of(float r) {  // order-preserving map float -> uint32
  const unsigned int u = __float_as_uint(r);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float val_of(unsigned int k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// ratio and effective L1 weight of one pixel (alignment.py:283-297)
__device__ __forceinline__ void ratio_weight(float x, float y, float m, float conf, float min_depth, float& r, float& w) {
  const float y_cl = fmaxf(y, min_depth);
  const float w_depth = 1.0f / fmaxf(y_cl, 1e-6f);
  const float wt = m * conf * w_depth;
  const float sign = x < 0.f ? -1.f : 1.f;  // sign(0) -> 1
  const float xp = x * sign, yp = y * sign;
  r = yp / fmaxf(xp, 1e-6f);
  w = wt * xp;
}

__global__ void scale_init_kernel(ScaleState* st, int B) {
  const int b = blockIdx.x;
  if (b >= B) return;
  ScaleState& s = st[b];
  if (threadIdx.x == 0) { s.sum_valid = 0; s.sum_depth = 0; s.total = 0; s.below = 0; s.count_below = 0; s.prefix = 0; s.min_depth = 0; s.scale = 0; }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) { s.hist_w[i] = 0; s.hist_n[i] = 0; }
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;
}

// pass A: sum(m), sum(y*m)
__global__ void __launch_bounds__(TPB) scale_mean_kernel(const float* __restrict__ y, const float* __restrict__ m, long long N, ScaleState* st) {
  __shared__ double sh[8];
  const int b = blockIdx.y;
  const float* yb = y + (size_t)b * N; const float* mb = m + (size_t)b * N;
  double sv = 0, sd = 0;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < N; i += (long long)gridDim.x * TPB) { const float mm = mb[i]; sv += mm; sd += (double)yb[i] * mm; }
  sv = block_sum(sv, sh); sd = block_sum(sd, sh);
  if (threadIdx.x == 0) { atomicAdd(&st[b].sum_valid, sv); atomicAdd(&st[b].sum_depth, sd); }
}
__global__ void scale_min_depth_kernel(ScaleState* st, int B) {
  const int b = threadIdx.x;
  if (b < B) st[b].min_depth = 0.1f * (float)(st[b].sum_depth / fmax(st[b].sum_valid, 1.0));
}

// pass B: total effective weight
__global__ void __launch_bounds__(TPB) scale_total_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ m,
                                                          const float* __restrict__ conf, long long N, ScaleState* st) {
  __shared__ double sh[8];
  const int b = blockIdx.y;
  const float md = st[b].min_depth;
  const size_t o = (size_t)b * N;
  double tw = 0;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < N; i += (long long)gridDim.x * TPB) {
    float r, w;
    ratio_weight(x[o + i], y[o + i], m[o + i], conf[o + i], md, r, w);
    tw += w;
  }
  tw = block_sum(tw, sh);
  if (threadIdx.x == 0) atomicAdd(&st[b].total, tw);
}

// radix pass p (0 = most significant byte): weight / count histogram of the keys that match the decided prefix
__global__ void __launch_bounds__(TPB) scale_hist_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ m,
                                                         const float* __restrict__ conf, long long N, ScaleState* st, int pass) {
  __shared__ double hw[256];
  __shared__ unsigned int hn[256];
  const int b = blockIdx.y;
  hw[threadIdx.x] = 0; hn[threadIdx.x] = 0;
  __syncthreads();
  const float md = st[b].min_depth;
  const unsigned int prefix = st[b].prefix;
  const int shift = 24 - 8 * pass;
  const unsigned int mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
  const size_t o = (size_t)b * N;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < N; i += (long long)gridDim.x * TPB) {
    float r, w;
    ratio_weight(x[o + i], y[o + i], m[o + i], conf[o + i], md, r, w);
    const unsigned int k = key_of(r);
    if ((k & mask) == prefix) {
      const unsigned int bin = (k >> shift) & 255u;
      atomicAdd(&hw[bin], (double)w);
      atomicAdd(&hn[bin], 1u);
    }
  }
  __syncthreads();
  if (hn[threadIdx.x]) { atomicAdd(&st[b].hist_w[threadIdx.x], hw[threadIdx.x]); atomicAdd(&st[b].hist_n[threadIdx.x], (unsigned long long)hn[threadIdx.x]); }
}

// pick the bin that holds the weighted median: the first non-empty bin whose cumulative weight reaches total / 2
// (torch.searchsorted(cumsum, 0.5 * total, side="left") on the sorted ratios, alignment.py:303-310)
__global__ void scale_select_kernel(ScaleState* st, int B, int pass, float* scales) {
  const int b = threadIdx.x;
  if (b >= B) return;
  ScaleState& s = st[b];
  const double target = 0.5 * s.total;
  const int shift = 24 - 8 * pass;
  double cum = s.below;
  int pick = -1, last = -1;
  for (int i = 0; i < 256; ++i) {
    if (s.hist_n[i] == 0) continue;
    last = i;
    if (cum + s.hist_w[i] >= target) { pick = i; break; }
    cum += s.hist_w[i];
  }
  if (pick < 0) { pick = last < 0 ? 0 : last; cum -= (last < 0 ? 0.0 : s.hist_w[last]); }  // rounding: clamp to the last element (idx_med.clamp(max=N-1))
  s.below = cum;
  s.prefix |= (unsigned int)pick << shift;
  for (int i = 0; i < 256; ++i) { s.hist_w[i] = 0; s.hist_n[i] = 0; }
  if (pass == 3) {
    float sc = val_of(s.prefix);
    if (sc <= 0.f) sc = -sc;  // scales[scales <= 0] *= -1
    s.scale = sc;
    scales[b] = sc;
  }
}

}  // namespace

extern "C" int lsvs_unproject_depth(const float* depth, const float* extrinsics, const float* intrinsics, float* world_points, int frames,
                                    int H, int W, void* stream) {
  LSVS_CHECK_ARG(depth && extrinsics && intrinsics && world_points && frames > 0 && H > 0 && W > 0, "unproject_depth: bad arguments");
  LSVS_CHECK_ARG(frames <= 65535, "unproject_depth: more than 65535 frames in one call");
  lsvs::ProfScope prof(lsvs::PROF_SIM3, (cudaStream_t)stream, 0, 16.0 * frames * H * (double)W);
  const int n = H * W;
  if ((n & 3) == 0 && ((uintptr_t)depth % 16 == 0) && ((uintptr_t)world_points % 16 == 0)) {
    const int chunks = (n + 127) >> 7, per_cta = TPB / 32;
    int bx = (chunks + per_cta - 1) / per_cta;
    const int cap = (lsvs::num_sms() * 8 + frames - 1) / frames;
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    unproject_kernel<<<dim3(bx, frames), TPB, 0, (cudaStream_t)stream>>>(depth, extrinsics, intrinsics, world_points, H, W);
  } else {
    unproject_scalar_kernel<<<dim3((n + TPB - 1) / TPB, frames), TPB, 0, (cudaStream_t)stream>>>(depth, extrinsics, intrinsics, world_points, H, W);
  }
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

extern "C" size_t lsvs_depth_scale_align_workspace_bytes(int B) { return sizeof(ScaleState) * (size_t)(B > 0 ? B : 1); }

extern "C" int lsvs_depth_scale_align(const float* depth_pred, const float* depth_gt, const float* mask, const float* conf, int B,
                                      long long N, float* scales, void* workspace, void* stream) {
  LSVS_CHECK_ARG(depth_pred && depth_gt && mask && conf && scales && workspace && B > 0 && B <= 1024 && N > 0, "depth_scale_align: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ScaleState* s = reinterpret_cast<ScaleState*>(workspace);
  lsvs::ProfScope prof(lsvs::PROF_ELEMENTWISE, st, 0, (8.0 + 5 * 16.0) * B * (double)N);
  const int bx = (int)((N + TPB * 8 - 1) / (TPB * 8) < 1184 ? (N + TPB * 8 - 1) / (TPB * 8) : 1184);
  dim3 grid(bx > 0 ? bx : 1, B);
  scale_init_kernel<<<B, 256, 0, st>>>(s, B);
  scale_mean_kernel<<<grid, TPB, 0, st>>>(depth_gt, mask, N, s);
  scale_min_depth_kernel<<<1, 1024, 0, st>>>(s, B);
  scale_total_kernel<<<grid, TPB, 0, st>>>(depth_pred, depth_gt, mask, conf, N, s);
  for (int pass = 0; pass < 4; ++pass) {
    scale_hist_kernel<<<grid, TPB, 0, st>>>(depth_pred, depth_gt, mask, conf, N, s, pass);
    scale_select_kernel<<<1, 1024, 0, st>>>(s, B, pass, scales);
  }
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
