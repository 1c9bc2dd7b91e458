// fp32-class ("precise") variants of the block operators, for the precision modes of the engine (lsvs_engine_config::precision):
//   1 = alignment head + camera-head trunk (the reference runs the camera head with autocast disabled, featureAligned_vggt.py:103-104,
//       and decodes the alignments in fp32, alignment_head.py:340), 2 = every block of the path (verification of the bf16 pipeline).
//
// GEMMs stay on the tcgen05 kernel of csrc/gemm.cu: an fp32 operand is split into two bf16 terms x = hi + lo (hi = bf16(x),
// lo = bf16(x - hi), 16 significant bits together) and the three significant partial products are one GEMM over a K axis three
// times as long:      A' = [A_hi | A_lo | A_hi]  (M, 3K),   W' = [W_hi | W_hi | W_lo]  (N, 3K)   =>   A' W'^T = A W^T (1 + O(2^-16))
// with fp32 accumulation in tensor memory.  The kernels here produce the split activations (LayerNorm, cast / GELU, attention
// output) and replace the bf16 flash attention by an fp32 one on the CUDA cores (q/k LayerNorm + RoPE in fp32 as well).
#include <cuda_bf16.h>

#include "host_common.h"
#include "precise.h"
#include "ptx.cuh"

namespace lsvs {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void split2(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// four consecutive values -> the three sections of a split row (section stride `sec` elements)
__device__ __forceinline__ void store_split4(__nv_bfloat16* row, long long sec, int col, float4 v) {
  __nv_bfloat16 h[4], l[4];
  split2(v.x, h[0], l[0]); split2(v.y, h[1], l[1]); split2(v.z, h[2], l[2]); split2(v.w, h[3], l[3]);
  const uint2 hv = *reinterpret_cast<uint2*>(h), lv = *reinterpret_cast<uint2*>(l);
  *reinterpret_cast<uint2*>(row + col) = hv;
  *reinterpret_cast<uint2*>(row + sec + col) = lv;
  *reinterpret_cast<uint2*>(row + 2 * sec + col) = hv;
}

// LayerNorm, one warp per row (D = 128*NV), output split [hi | lo | hi] with row stride ld_out (>= 3*D)
template <int NV>
__global__ void __launch_bounds__(256) layernorm_split_kernel(const float* __restrict__ x, long long ld_in, const float* __restrict__ w,
                                                              const float* __restrict__ b, float eps, __nv_bfloat16* __restrict__ out,
                                                              long long ld_out, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= rows) return;
  constexpr int D = NV * 128;
  const float4* src = reinterpret_cast<const float4*>(x + m * ld_in);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) { v[i] = src[lane + 32 * i]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
    q += (a * a + c * c) + (d * d + e * e);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = w ? __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 h = b ? __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + h.x; o.y = (v[i].y - mean) * rstd * g.y + h.y;
    o.z = (v[i].z - mean) * rstd * g.z + h.z; o.w = (v[i].w - mean) * rstd * g.w + h.w;
    store_split4(out + m * ld_out, D, 4 * (lane + 32 * i), o);
  }
}

// (rows, cols) fp32 -> split (rows, 3*cols) bf16, optionally through the exact GELU (act 1) or SiLU (act 2)
__global__ void __launch_bounds__(256) cast_split_kernel(const float* __restrict__ x, long long ld_in, __nv_bfloat16* __restrict__ out,
                                                         long long ld_out, long long rows, int cols, int gelu) {
  const int c4 = cols / 4;
  const long long total = rows * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4;
    const int c = (int)(i % c4);
    float4 v = reinterpret_cast<const float4*>(x + r * ld_in)[c];
    if (gelu == 1) {
      v.x = 0.5f * v.x * (1.f + erff(v.x * 0.70710678118654752f)); v.y = 0.5f * v.y * (1.f + erff(v.y * 0.70710678118654752f));
      v.z = 0.5f * v.z * (1.f + erff(v.z * 0.70710678118654752f)); v.w = 0.5f * v.w * (1.f + erff(v.w * 0.70710678118654752f));
    } else if (gelu == 2) {
      v.x = v.x / (1.f + expf(-v.x)); v.y = v.y / (1.f + expf(-v.y)); v.z = v.z / (1.f + expf(-v.z)); v.w = v.w / (1.f + expf(-v.w));
    }
    store_split4(out + r * ld_out, cols, 4 * c, v);
  }
}

// weights (rows, k_in) fp32 -> (rows, 3*k_pad) bf16 = [hi | hi | lo], zero padded
__global__ void __launch_bounds__(256) pack_split_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, long long rows,
                                                         int k_in, int k_pad) {
  const long long total = rows * k_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / k_pad;
    const int c = (int)(i % k_pad);
    __nv_bfloat16 hi, lo;
    split2(c < k_in ? w[r * k_in + c] : 0.f, hi, lo);
    __nv_bfloat16* o = out + r * 3 * k_pad;
    o[c] = hi; o[k_pad + c] = hi; o[2 * k_pad + c] = lo;
  }
}

// im2col of the 14x14/14 patch convolution in fp32: (frames,3,H,W) in [0,1] -> (frames*gh*gw, 640), ImageNet-normalised,
// taps in Conv2d (c, ky, kx) order, zero padded from 588
__global__ void __launch_bounds__(256) patch_unfold_f32_kernel(const float* __restrict__ img, float* __restrict__ out, int H, int W,
                                                               int gh, int gw, long long total) {
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % 640);
    const long long patch = i / 640;
    float v = 0.f;
    if (k < 588) {
      const int c = k / 196, rem = k % 196, ky = rem / 14, kx = rem % 14;
      const int gx = (int)(patch % gw), gy = (int)((patch / gw) % gh);
      const long long f = patch / ((long long)gw * gh);
      v = (img[((f * 3 + c) * H + (gy * 14 + ky)) * W + gx * 14 + kx] - mean[c]) / stdv[c];
    }
    out[i] = v;
  }
}

// per-head LayerNorm + RoPE in place on fp32 columns [col0, col0 + n_heads*HD) of `buf` (one warp per (row, head))
template <int HD>
__global__ void __launch_bounds__(256) headnorm_rope_f32_kernel(float* __restrict__ buf, long long ld, long long rows, int col0, int n_heads,
                                                                const float* __restrict__ w, const float* __restrict__ b, float eps,
                                                                int rope_mode, const float2* __restrict__ tab, int tpf, int nsp, int gw,
                                                                const int* __restrict__ pos_ids, int period) {
  constexpr int VPL = HD / 32;
  const int lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= rows * n_heads) return;
  const long long m = item / n_heads;
  const int head = (int)(item % n_heads);
  float* p = buf + m * ld + col0 + head * HD + VPL * lane;
  float v[VPL];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < VPL; ++c) { v[c] = p[c]; s += v[c]; }
  const float mean = warp_sum(s) * (1.0f / HD);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < VPL; ++c) { const float d = v[c] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / HD) + eps);
#pragma unroll
  for (int c = 0; c < VPL; ++c) v[c] = (v[c] - mean) * rstd * __ldg(w + VPL * lane + c) + __ldg(b + VPL * lane + c);
  if (rope_mode != ROPE_NONE) {
    int pos, R, e0, xr;
    if (rope_mode == ROPE_2D) {
      const int t = (int)(m % tpf);
      int py = 0, px = 0;
      if (t >= nsp) { const int pp = t - nsp; py = pp / gw + 1; px = pp % gw + 1; }
      pos = (lane >> 4) ? px : py; R = HD / 2; e0 = VPL * (lane & 15); xr = 8;
    } else {
      pos = __ldg(pos_ids + (int)(m % period)); R = HD; e0 = VPL * lane; xr = 16;
    }
    const int nf = R / 2;
#pragma unroll
    for (int c = 0; c < VPL; ++c) {
      const float other = __shfl_xor_sync(0xffffffffu, v[c], xr);
      const int e = e0 + c;
      const float2 cs = __ldg(tab + (size_t)pos * nf + (e % nf));
      v[c] = e < nf ? v[c] * cs.x - other * cs.y : v[c] * cs.x + other * cs.y;
    }
  }
#pragma unroll
  for (int c = 0; c < VPL; ++c) p[c] = v[c];
}

// fp32 flash attention on the CUDA cores.  CTA = 32 queries x 4 threads; thread (q, part) owns the float4 columns 4*i + part of
// the head dim (conflict-free shared-memory reads of the key / value rows, broadcast over the 8 queries of a warp).  Keys are
// staged 32 at a time; scores of a tile are reduced over the 4 parts with two shuffles, the running maximum / sum / output are
// rescaled once per tile.  Output: split bf16 rows [hi | lo | hi] (sections `sec` apart) = the A operand of the projection GEMM.
template <int HD>
__global__ void __launch_bounds__(128) attn_f32_kernel(const float* __restrict__ Q, long long ldq, const float* __restrict__ K, long long ldk,
                                                       const float* __restrict__ V, long long ldv, __nv_bfloat16* __restrict__ O,
                                                       long long ldo, long long sec, int Lq, int Lk, float scale_log2e,
                                                       float* __restrict__ O32, long long ldo32, float* __restrict__ LSE) {
  constexpr int KT = 32, NF = HD / 16;   // keys per tile, float4 per thread
  __shared__ float4 sK[KT][HD / 4], sV[KT][HD / 4];
  const int tid = threadIdx.x, part = tid & 3, qi = tid >> 2;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int q_idx = blockIdx.x * 32 + qi;
  const bool q_ok = q_idx < Lq;
  float4 q[NF], o[NF];
  {
    const float4* qp = reinterpret_cast<const float4*>(Q + ((long long)batch * Lq + (q_ok ? q_idx : 0)) * ldq + head * HD);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      const float4 t = qp[4 * i + part];
      q[i] = make_float4(t.x * scale_log2e, t.y * scale_log2e, t.z * scale_log2e, t.w * scale_log2e);
      o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float m_run = -INFINITY, l_run = 0.f;
  for (int k0 = 0; k0 < Lk; k0 += KT) {
    __syncthreads();
    for (int i = tid; i < KT * (HD / 4); i += 128) {
      const int r = i / (HD / 4), c = i % (HD / 4);
      const bool ok = k0 + r < Lk;
      const long long row = (long long)batch * Lk + (ok ? k0 + r : 0);
      sK[r][c] = ok ? reinterpret_cast<const float4*>(K + row * ldk + head * HD)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      sV[r][c] = ok ? reinterpret_cast<const float4*>(V + row * ldv + head * HD)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    float s[KT];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const float4 kk = sK[j][4 * i + part];
        a = fmaf(q[i].x, kk.x, a); a = fmaf(q[i].y, kk.y, a); a = fmaf(q[i].z, kk.z, a); a = fmaf(q[i].w, kk.w, a);
      }
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      s[j] = (k0 + j < Lk) ? a : -INFINITY;
      mx = fmaxf(mx, s[j]);
    }
    const float m_new = fmaxf(m_run, mx);
    const float alpha = exp2f(m_run - m_new);   // first tile: exp2(-inf) = 0
    l_run *= alpha;
#pragma unroll
    for (int i = 0; i < NF; ++i) { o[i].x *= alpha; o[i].y *= alpha; o[i].z *= alpha; o[i].w *= alpha; }
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float p = exp2f(s[j] - m_new);
      l_run += p;
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const float4 vv = sV[j][4 * i + part];
        o[i].x = fmaf(p, vv.x, o[i].x); o[i].y = fmaf(p, vv.y, o[i].y); o[i].z = fmaf(p, vv.z, o[i].z); o[i].w = fmaf(p, vv.w, o[i].w);
      }
    }
    m_run = m_new;
  }
  if (q_ok) {
    const float inv = 1.0f / l_run;
    if (O) {
      __nv_bfloat16* row = O + ((long long)batch * Lq + q_idx) * ldo;
#pragma unroll
      for (int i = 0; i < NF; ++i)
        store_split4(row, sec, head * HD + 4 * (4 * i + part), make_float4(o[i].x * inv, o[i].y * inv, o[i].z * inv, o[i].w * inv));
    }
    if (O32) {  // training forward: fp32 output and the row's log2-sum-exp of the scaled scores, for the backward kernels
      float4* row = reinterpret_cast<float4*>(O32 + ((long long)batch * Lq + q_idx) * ldo32 + head * HD);
#pragma unroll
      for (int i = 0; i < NF; ++i) row[4 * i + part] = make_float4(o[i].x * inv, o[i].y * inv, o[i].z * inv, o[i].w * inv);
    }
    if (LSE && part == 0) LSE[((long long)batch * gridDim.y + head) * Lq + q_idx] = m_run + log2f(l_run);
  }
}

// ---- backward of the fp32 attention (training of the alignment head, alignment_head.py:361,385: only the head trains).
// With s_ij = scale * q_i . k_j (in log2 units), P_ij = 2^(s_ij - lse_i), D_i = dO_i . O_i:
//   dV_j = sum_i P_ij dO_i     dS_ij = P_ij (dO_i . V_j - D_i)     dQ_i = scale sum_j dS_ij K_j     dK_j = scale sum_i dS_ij Q_i
// Same thread layout as the forward: 32 rows x 4 threads, each thread owns every 4th float4 of the head dim.
template <int HD>
__global__ void __launch_bounds__(128) attn_f32_bwd_dq_kernel(const float* __restrict__ Q, long long ldq, const float* __restrict__ K, long long ldk,
                                                              const float* __restrict__ V, long long ldv, const float* __restrict__ O,
                                                              const float* __restrict__ dO, long long ldo, const float* __restrict__ LSE,
                                                              float* __restrict__ Dbuf, float* __restrict__ dQ, long long lddq, int Lq, int Lk,
                                                              float scale) {
  constexpr int KT = 32, NF = HD / 16;
  __shared__ float4 sK[KT][HD / 4], sV[KT][HD / 4];
  const int tid = threadIdx.x, part = tid & 3, qi = tid >> 2;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int q_idx = blockIdx.x * 32 + qi;
  const bool q_ok = q_idx < Lq;
  const long long qrow = (long long)batch * Lq + (q_ok ? q_idx : 0);
  const float sl = scale * 1.4426950408889634f;
  float4 q[NF], go[NF], dq[NF];
  float Di = 0.f;
  {
    const float4* qp = reinterpret_cast<const float4*>(Q + qrow * ldq + head * HD);
    const float4* op = reinterpret_cast<const float4*>(O + qrow * ldo + head * HD);
    const float4* gp = reinterpret_cast<const float4*>(dO + qrow * ldo + head * HD);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      const float4 t = qp[4 * i + part], oo = op[4 * i + part];
      q[i] = make_float4(t.x * sl, t.y * sl, t.z * sl, t.w * sl);
      go[i] = gp[4 * i + part];
      dq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      Di += go[i].x * oo.x + go[i].y * oo.y + go[i].z * oo.z + go[i].w * oo.w;
    }
    Di += __shfl_xor_sync(0xffffffffu, Di, 1);
    Di += __shfl_xor_sync(0xffffffffu, Di, 2);
  }
  const long long stat = ((long long)batch * gridDim.y + head) * Lq + (q_ok ? q_idx : 0);
  const float lse = LSE[stat];
  if (q_ok && part == 0) Dbuf[stat] = Di;
  for (int k0 = 0; k0 < Lk; k0 += KT) {
    __syncthreads();
    for (int i = tid; i < KT * (HD / 4); i += 128) {
      const int r = i / (HD / 4), c = i % (HD / 4);
      const bool ok = k0 + r < Lk;
      const long long row = (long long)batch * Lk + (ok ? k0 + r : 0);
      sK[r][c] = ok ? reinterpret_cast<const float4*>(K + row * ldk + head * HD)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      sV[r][c] = ok ? reinterpret_cast<const float4*>(V + row * ldv + head * HD)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < KT; ++j) {
      float a = 0.f, dp = 0.f;
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const float4 kk = sK[j][4 * i + part], vv = sV[j][4 * i + part];
        a = fmaf(q[i].x, kk.x, a); a = fmaf(q[i].y, kk.y, a); a = fmaf(q[i].z, kk.z, a); a = fmaf(q[i].w, kk.w, a);
        dp = fmaf(go[i].x, vv.x, dp); dp = fmaf(go[i].y, vv.y, dp); dp = fmaf(go[i].z, vv.z, dp); dp = fmaf(go[i].w, vv.w, dp);
      }
      a += __shfl_xor_sync(0xffffffffu, a, 1); dp += __shfl_xor_sync(0xffffffffu, dp, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2); dp += __shfl_xor_sync(0xffffffffu, dp, 2);
      const float p = (k0 + j < Lk) ? exp2f(a - lse) : 0.f;
      const float ds = p * (dp - Di);
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const float4 kk = sK[j][4 * i + part];
        dq[i].x = fmaf(ds, kk.x, dq[i].x); dq[i].y = fmaf(ds, kk.y, dq[i].y); dq[i].z = fmaf(ds, kk.z, dq[i].z); dq[i].w = fmaf(ds, kk.w, dq[i].w);
      }
    }
  }
  if (q_ok) {
    float4* row = reinterpret_cast<float4*>(dQ + ((long long)batch * Lq + q_idx) * lddq + head * HD);
#pragma unroll
    for (int i = 0; i < NF; ++i) row[4 * i + part] = make_float4(dq[i].x * scale, dq[i].y * scale, dq[i].z * scale, dq[i].w * scale);
  }
}

template <int HD>
__global__ void __launch_bounds__(128) attn_f32_bwd_dkv_kernel(const float* __restrict__ Q, long long ldq, const float* __restrict__ K, long long ldk,
                                                               const float* __restrict__ V, long long ldv, const float* __restrict__ dO,
                                                               long long ldo, const float* __restrict__ LSE, const float* __restrict__ Dbuf,
                                                               float* __restrict__ dK, long long lddk, float* __restrict__ dV, long long lddv,
                                                               int Lq, int Lk, float scale) {
  constexpr int QTILE = 32, NF = HD / 16;
  __shared__ float4 sQ[QTILE][HD / 4], sG[QTILE][HD / 4];
  __shared__ float sL[QTILE], sD[QTILE];
  const int tid = threadIdx.x, part = tid & 3, kj = tid >> 2;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int k_idx = blockIdx.x * 32 + kj;
  const bool k_ok = k_idx < Lk;
  const long long krow = (long long)batch * Lk + (k_ok ? k_idx : 0);
  const float sl = scale * 1.4426950408889634f;
  float4 k[NF], v[NF], dk[NF], dv[NF];
  {
    const float4* kp = reinterpret_cast<const float4*>(K + krow * ldk + head * HD);
    const float4* vp = reinterpret_cast<const float4*>(V + krow * ldv + head * HD);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      k[i] = kp[4 * i + part]; v[i] = vp[4 * i + part];
      dk[i] = make_float4(0.f, 0.f, 0.f, 0.f); dv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const long long stat0 = ((long long)batch * gridDim.y + head) * Lq;
  for (int q0 = 0; q0 < Lq; q0 += QTILE) {
    __syncthreads();
    for (int i = tid; i < QTILE * (HD / 4); i += 128) {
      const int r = i / (HD / 4), c = i % (HD / 4);
      const bool ok = q0 + r < Lq;
      const long long row = (long long)batch * Lq + (ok ? q0 + r : 0);
      float4 t = ok ? reinterpret_cast<const float4*>(Q + row * ldq + head * HD)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      sQ[r][c] = make_float4(t.x * sl, t.y * sl, t.z * sl, t.w * sl);
      sG[r][c] = ok ? reinterpret_cast<const float4*>(dO + row * ldo + head * HD)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid < QTILE) {
      const bool ok = q0 + tid < Lq;
      sL[tid] = ok ? LSE[stat0 + q0 + tid] : INFINITY;   // 2^(s - inf) = 0 for rows past Lq
      sD[tid] = ok ? Dbuf[stat0 + q0 + tid] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < QTILE; ++r) {
      float a = 0.f, dp = 0.f;
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const float4 qq = sQ[r][4 * i + part], gg = sG[r][4 * i + part];
        a = fmaf(qq.x, k[i].x, a); a = fmaf(qq.y, k[i].y, a); a = fmaf(qq.z, k[i].z, a); a = fmaf(qq.w, k[i].w, a);
        dp = fmaf(gg.x, v[i].x, dp); dp = fmaf(gg.y, v[i].y, dp); dp = fmaf(gg.z, v[i].z, dp); dp = fmaf(gg.w, v[i].w, dp);
      }
      a += __shfl_xor_sync(0xffffffffu, a, 1); dp += __shfl_xor_sync(0xffffffffu, dp, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2); dp += __shfl_xor_sync(0xffffffffu, dp, 2);
      const float p = exp2f(a - sL[r]);
      const float ds = p * (dp - sD[r]);
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const float4 qq = sQ[r][4 * i + part], gg = sG[r][4 * i + part];
        dv[i].x = fmaf(p, gg.x, dv[i].x); dv[i].y = fmaf(p, gg.y, dv[i].y); dv[i].z = fmaf(p, gg.z, dv[i].z); dv[i].w = fmaf(p, gg.w, dv[i].w);
        dk[i].x = fmaf(ds, qq.x, dk[i].x); dk[i].y = fmaf(ds, qq.y, dk[i].y); dk[i].z = fmaf(ds, qq.z, dk[i].z); dk[i].w = fmaf(ds, qq.w, dk[i].w);
      }
    }
  }
  if (k_ok) {
    const float un = 1.0f / 1.4426950408889634f;   // the staged queries carry scale * log2(e); dK needs scale alone
    float4* rk = reinterpret_cast<float4*>(dK + ((long long)batch * Lk + k_idx) * lddk + head * HD);
    float4* rv = reinterpret_cast<float4*>(dV + ((long long)batch * Lk + k_idx) * lddv + head * HD);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      rk[4 * i + part] = make_float4(dk[i].x * un, dk[i].y * un, dk[i].z * un, dk[i].w * un);
      rv[4 * i + part] = dv[i];
    }
  }
}

int blocks_for(long long items, int threads = 256) {
  long long b = (items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int layernorm_split(const float* x, long long ld_in, const float* w, const float* b, float eps, void* out, long long ld_out,
                    long long rows, int D, cudaStream_t st) {
  LSVS_CHECK_ARG(x && out && ld_out >= 3LL * D, "layernorm_split: bad arguments");
  LSVS_CHECK_ARG(D == 512 || D == 1024 || D == 2048, "layernorm_split: D=%d unsupported (512/1024/2048)", D);
  if (rows == 0) return LSVS_OK;
  ProfScope prof(PROF_ELEMENTWISE, st, 0, (double)rows * D * 10);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (D == 512) layernorm_split_kernel<4><<<grid, 256, 0, st>>>(x, ld_in, w, b, eps, o, ld_out, rows);
  else if (D == 1024) layernorm_split_kernel<8><<<grid, 256, 0, st>>>(x, ld_in, w, b, eps, o, ld_out, rows);
  else layernorm_split_kernel<16><<<grid, 256, 0, st>>>(x, ld_in, w, b, eps, o, ld_out, rows);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int cast_split_act(const float* x, long long ld_in, void* out, long long ld_out, long long rows, int cols, int act, cudaStream_t st) {
  LSVS_CHECK_ARG(x && out && cols % 4 == 0 && ld_in % 4 == 0 && ld_out >= 3LL * cols && ld_out % 4 == 0, "cast_split: bad arguments");
  if (rows == 0) return LSVS_OK;
  ProfScope prof(PROF_ELEMENTWISE, st, 0, (double)rows * cols * 10);
  cast_split_kernel<<<blocks_for(rows * (cols / 4)), 256, 0, st>>>(x, ld_in, reinterpret_cast<__nv_bfloat16*>(out), ld_out, rows, cols, act);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
int cast_split(const float* x, long long ld_in, void* out, long long ld_out, long long rows, int cols, bool gelu, cudaStream_t st) {
  return cast_split_act(x, ld_in, out, ld_out, rows, cols, gelu ? 1 : 0, st);
}

int pack_weight_split(const float* w, void* out, long long rows, int k_in, int k_pad, cudaStream_t st) {
  pack_split_kernel<<<blocks_for(rows * k_pad), 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16*>(out), rows, k_in, k_pad);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int patch_unfold_f32(const float* img, float* out, int frames, int H, int W, cudaStream_t st) {
  LSVS_CHECK_ARG(img && out && frames > 0 && H % 14 == 0 && W % 14 == 0, "patch_unfold_f32: image size must be a multiple of 14");
  const long long total = (long long)frames * (H / 14) * (W / 14) * 640;
  patch_unfold_f32_kernel<<<blocks_for(total), 256, 0, st>>>(img, out, H, W, H / 14, W / 14, total);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int headnorm_rope_f32(float* buf, long long ld, long long rows, int col0, int n_heads, int hd, const float* w, const float* b, float eps,
                      int rope_mode, const float2* tab, int tpf, int nsp, int gw, const int* pos_ids, int period, cudaStream_t st) {
  LSVS_CHECK_ARG(buf && w && b && (hd == 64 || hd == 128) && n_heads > 0, "headnorm_rope_f32: bad arguments");
  LSVS_CHECK_ARG(rope_mode == ROPE_NONE || tab, "headnorm_rope_f32: rope table missing");
  LSVS_CHECK_ARG(rope_mode != ROPE_2D || (tpf > 0 && gw > 0), "headnorm_rope_f32: 2-D rope needs the token grid");
  LSVS_CHECK_ARG(rope_mode != ROPE_1D || (pos_ids && period > 0), "headnorm_rope_f32: 1-D rope needs position ids");
  if (rows == 0) return LSVS_OK;
  ProfScope prof(PROF_ELEMENTWISE, st, 0, (double)rows * n_heads * hd * 8);
  const long long items = rows * n_heads;
  const unsigned grid = (unsigned)((items + 7) / 8);
  if (hd == 64) headnorm_rope_f32_kernel<64><<<grid, 256, 0, st>>>(buf, ld, rows, col0, n_heads, w, b, eps, rope_mode, tab, tpf, nsp, gw, pos_ids, period);
  else headnorm_rope_f32_kernel<128><<<grid, 256, 0, st>>>(buf, ld, rows, col0, n_heads, w, b, eps, rope_mode, tab, tpf, nsp, gw, pos_ids, period);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int attention_f32(const AttentionF32Args& a, cudaStream_t st) {
  LSVS_CHECK_ARG(a.q && a.k && a.v && (a.o || a.o32), "attention_f32: null pointer");
  LSVS_CHECK_ARG(a.batches > 0 && a.heads > 0 && a.Lq > 0 && a.Lk > 0 && a.batches <= 65535 && a.heads <= 65535, "attention_f32: bad shape");
  LSVS_CHECK_ARG(a.head_dim == 64 || a.head_dim == 128, "attention_f32: head_dim %d unsupported (64 or 128)", a.head_dim);
  LSVS_CHECK_ARG(a.ldq % 4 == 0 && a.ldk % 4 == 0 && a.ldv % 4 == 0 && a.ldo % 4 == 0 && a.section % 4 == 0, "attention_f32: strides must be multiples of 4");
  ProfScope prof(a.Lk >= 2048 ? PROF_ATTENTION_GLOBAL : PROF_ATTENTION, st, 4.0 * a.batches * (double)a.heads * a.Lq * (double)a.Lk * a.head_dim, 0);
  dim3 grid((a.Lq + 31) / 32, a.heads, a.batches);
  const float sl = a.scale * 1.4426950408889634f;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.o);
  if (a.head_dim == 64) attn_f32_kernel<64><<<grid, 128, 0, st>>>(a.q, a.ldq, a.k, a.ldk, a.v, a.ldv, o, a.ldo, a.section, a.Lq, a.Lk, sl, a.o32, a.ldo32, a.lse);
  else attn_f32_kernel<128><<<grid, 128, 0, st>>>(a.q, a.ldq, a.k, a.ldk, a.v, a.ldv, o, a.ldo, a.section, a.Lq, a.Lk, sl, a.o32, a.ldo32, a.lse);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int attention_f32_backward(const AttentionF32BwdArgs& a, cudaStream_t st) {
  LSVS_CHECK_ARG(a.q && a.k && a.v && a.o && a.d_o && a.lse && a.d_buf && a.dq && a.dk && a.dv, "attention_f32_backward: null pointer");
  LSVS_CHECK_ARG(a.batches > 0 && a.heads > 0 && a.Lq > 0 && a.Lk > 0 && a.batches <= 65535 && a.heads <= 65535, "attention_f32_backward: bad shape");
  LSVS_CHECK_ARG(a.head_dim == 64 || a.head_dim == 128, "attention_f32_backward: head_dim %d unsupported (64 or 128)", a.head_dim);
  LSVS_CHECK_ARG(a.ldq % 4 == 0 && a.ldk % 4 == 0 && a.ldv % 4 == 0 && a.ldo % 4 == 0 && a.lddq % 4 == 0 && a.lddk % 4 == 0 && a.lddv % 4 == 0,
                 "attention_f32_backward: strides must be multiples of 4");
  ProfScope prof(PROF_ATTENTION, st, 10.0 * a.batches * (double)a.heads * a.Lq * (double)a.Lk * a.head_dim, 0);
  dim3 gq((a.Lq + 31) / 32, a.heads, a.batches), gk((a.Lk + 31) / 32, a.heads, a.batches);
  if (a.head_dim == 64) {
    attn_f32_bwd_dq_kernel<64><<<gq, 128, 0, st>>>(a.q, a.ldq, a.k, a.ldk, a.v, a.ldv, a.o, a.d_o, a.ldo, a.lse, a.d_buf, a.dq, a.lddq, a.Lq, a.Lk, a.scale);
    LSVS_LAUNCH_CHECK();
    attn_f32_bwd_dkv_kernel<64><<<gk, 128, 0, st>>>(a.q, a.ldq, a.k, a.ldk, a.v, a.ldv, a.d_o, a.ldo, a.lse, a.d_buf, a.dk, a.lddk, a.dv, a.lddv, a.Lq, a.Lk, a.scale);
  } else {
    attn_f32_bwd_dq_kernel<128><<<gq, 128, 0, st>>>(a.q, a.ldq, a.k, a.ldk, a.v, a.ldv, a.o, a.d_o, a.ldo, a.lse, a.d_buf, a.dq, a.lddq, a.Lq, a.Lk, a.scale);
    LSVS_LAUNCH_CHECK();
    attn_f32_bwd_dkv_kernel<128><<<gk, 128, 0, st>>>(a.q, a.ldq, a.k, a.ldk, a.v, a.ldv, a.d_o, a.ldo, a.lse, a.d_buf, a.dk, a.lddk, a.dv, a.lddv, a.Lq, a.Lk, a.scale);
  }
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

}  // namespace lsvs
