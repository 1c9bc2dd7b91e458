// fp32 kernels for the latency-bound tail of the path: the alignment decode (alignment_head.py:427-540, which the
// reference runs with autocast disabled), GatedUpdate (gated_update.py:43-78) and the camera-head trunk.  Few rows
// (M = frames per chunk), so the work is weight streaming: one warp per output column keeps its weight row in
// flight once and applies it to every input row.
#include "small_f32.h"
#include "host_common.h"

namespace lsvs {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

constexpr int LIN_MT = 32;  // rows per launch slice (grid.y walks slices)
constexpr int LIN_NW = 4;   // output columns per warp: each x load feeds 4 weight rows

// y[m,n] = epi( sum_k in_act(x[m,k]) * W[n,k] + b[n] ),  m in [m0, m0+32), n in 4 consecutive columns per warp.
template <bool VEC>
__global__ void __launch_bounds__(256) linear_f32_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ W,
                                                         const float* __restrict__ b, float* __restrict__ y, long long ldy, int M,
                                                         int N, int K, int in_act, int out_act, const float* __restrict__ gamma,
                                                         int residual) {
  pdl_launch_dependents();
  pdl_wait();  // programmatic dependent launch (host_common.h)
  const int lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * LIN_NW;
  const int m0 = blockIdx.y * LIN_MT;
  const int mt = min(LIN_MT, M - m0);
  if (n0 >= N) return;
  float acc[LIN_MT][LIN_NW];
#pragma unroll
  for (int m = 0; m < LIN_MT; ++m)
#pragma unroll
    for (int j = 0; j < LIN_NW; ++j) acc[m][j] = 0.f;
  const float* wrow[LIN_NW];
#pragma unroll
  for (int j = 0; j < LIN_NW; ++j) wrow[j] = W + (size_t)min(n0 + j, N - 1) * K;
  if constexpr (VEC) {
    for (int k0 = lane * 4; k0 < K; k0 += 128) {
      float4 w4[LIN_NW];
#pragma unroll
      for (int j = 0; j < LIN_NW; ++j) w4[j] = __ldg(reinterpret_cast<const float4*>(wrow[j] + k0));
#pragma unroll
      for (int m = 0; m < LIN_MT; ++m) {
        if (m < mt) {
          float4 x4 = *reinterpret_cast<const float4*>(x + (size_t)(m0 + m) * ldx + k0);
          if (in_act == ACT_SILU) { x4.x = silu(x4.x); x4.y = silu(x4.y); x4.z = silu(x4.z); x4.w = silu(x4.w); }
          else if (in_act == ACT_GELU) { x4.x = gelu_erf(x4.x); x4.y = gelu_erf(x4.y); x4.z = gelu_erf(x4.z); x4.w = gelu_erf(x4.w); }
#pragma unroll
          for (int j = 0; j < LIN_NW; ++j)
            acc[m][j] = fmaf(w4[j].x, x4.x, fmaf(w4[j].y, x4.y, fmaf(w4[j].z, x4.z, fmaf(w4[j].w, x4.w, acc[m][j]))));
        }
      }
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      float w[LIN_NW];
#pragma unroll
      for (int j = 0; j < LIN_NW; ++j) w[j] = __ldg(wrow[j] + k);
#pragma unroll
      for (int m = 0; m < LIN_MT; ++m) {
        if (m < mt) {
          float xv = x[(size_t)(m0 + m) * ldx + k];
          if (in_act == ACT_SILU) xv = silu(xv);
          else if (in_act == ACT_GELU) xv = gelu_erf(xv);
#pragma unroll
          for (int j = 0; j < LIN_NW; ++j) acc[m][j] = fmaf(w[j], xv, acc[m][j]);
        }
      }
    }
  }
  float mine[LIN_NW];
#pragma unroll
  for (int j = 0; j < LIN_NW; ++j) mine[j] = 0.f;
#pragma unroll
  for (int m = 0; m < LIN_MT; ++m)
#pragma unroll
    for (int j = 0; j < LIN_NW; ++j) {
      const float s = warp_sum(acc[m][j]);
      if (lane == m) mine[j] = s;
    }
  if (lane < mt) {
#pragma unroll
    for (int j = 0; j < LIN_NW; ++j) {
      const int n = n0 + j;
      if (n >= N) break;
      float v = mine[j] + (b ? __ldg(b + n) : 0.f);
      if (out_act == ACT_GELU) v = gelu_erf(v);
      else if (out_act == ACT_SIGMOID) v = sigmoidf(v);
      float* dst = y + (size_t)(m0 + lane) * ldy + n;
      if (gamma) v *= __ldg(gamma + n);
      if (residual) v += *dst;
      *dst = v;
    }
  }
}

// Skinny fp32 GEMM for K % 4 == 0: a 256-thread block owns 8 output columns x (8*MS) rows.  The 8 warps are arranged as
// MS row groups (8 rows each) x KS slices of K; each lane streams float4s of the weight rows (coalesced) and of the x
// rows (L1/L2 resident), partial sums are reduced by shuffles, then across the K slices through shared memory.
// Many small blocks (N/8 per row tile) keep all SMs busy even for N = 512; weights are read once per row tile.
template <int MS, int KS>
__global__ void __launch_bounds__(256) linear_f32_tiled_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ W,
                                                               const float* __restrict__ b, float* __restrict__ y, long long ldy,
                                                               int M, int N, int K, int in_act, int out_act,
                                                               const float* __restrict__ gamma, int residual) {
  pdl_launch_dependents();
  pdl_wait();  // programmatic dependent launch (host_common.h)
  static_assert(MS * KS == 8, "8 warps");
  constexpr int NC = 8, RW = 8;  // columns per block, rows per warp
  __shared__ float red[KS][MS * RW][NC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ms = warp % MS, ks = warp / MS;
  const int n0 = blockIdx.x * NC;
  const int m0 = blockIdx.y * (MS * RW) + ms * RW;
  float acc[RW][NC];
#pragma unroll
  for (int r = 0; r < RW; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[r][c] = 0.f;
  const int kchunk = ((K / 4 + KS - 1) / KS) * 4;  // K slice of this warp (multiple of 4)
  const int kbeg = ks * kchunk, kend = min(K, kbeg + kchunk);
  for (int k = kbeg + lane * 4; k < kend; k += 128) {
    float4 w4[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) w4[c] = __ldg(reinterpret_cast<const float4*>(W + (size_t)min(n0 + c, N - 1) * K + k));
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      if (m0 + r < M) {
        float4 x4 = *reinterpret_cast<const float4*>(x + (size_t)(m0 + r) * ldx + k);
        if (in_act == ACT_SILU) { x4.x = silu(x4.x); x4.y = silu(x4.y); x4.z = silu(x4.z); x4.w = silu(x4.w); }
        else if (in_act == ACT_GELU) { x4.x = gelu_erf(x4.x); x4.y = gelu_erf(x4.y); x4.z = gelu_erf(x4.z); x4.w = gelu_erf(x4.w); }
#pragma unroll
        for (int c = 0; c < NC; ++c)
          acc[r][c] = fmaf(w4[c].x, x4.x, fmaf(w4[c].y, x4.y, fmaf(w4[c].z, x4.z, fmaf(w4[c].w, x4.w, acc[r][c]))));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RW; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float sres = warp_sum(acc[r][c]);
      if (lane == 0) red[ks][ms * RW + r][c] = sres;
    }
  __syncthreads();
  // 64..256 outputs per block: thread t -> (row t / 8, col t % 8)
  const int t = threadIdx.x;
  if (t < MS * RW * NC) {
    const int r = t / NC, c = t % NC;
    const int m = blockIdx.y * (MS * RW) + r, n = n0 + c;
    if (m < M && n < N) {
      float v = 0.f;
#pragma unroll
      for (int q = 0; q < KS; ++q) v += red[q][r][c];
      v += b ? __ldg(b + n) : 0.f;
      if (out_act == ACT_GELU) v = gelu_erf(v);
      else if (out_act == ACT_SIGMOID) v = sigmoidf(v);
      float* dst = y + (size_t)m * ldy + n;
      if (gamma) v *= __ldg(gamma + n);
      if (residual) v += *dst;
      *dst = v;
    }
  }
}

// One warp per (batch, head, query): optional per-head LayerNorm on q/k, optional 1-D RoPE, softmax over Nk keys.
// Lane holds elements e = lane + 32*j, so rotate-half partners (e, e + hd/2) sit in the same lane.
template <int HD>
__global__ void __launch_bounds__(128) attn_small_kernel(SmallAttnArgs a) {
  pdl_launch_dependents();
  pdl_wait();  // programmatic dependent launch (host_common.h)
  constexpr int EPL = HD / 32;
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int total = a.B * a.H * a.Nq;
  if (w >= total) return;
  const int i = w % a.Nq, h = (w / a.Nq) % a.H, bb = w / (a.Nq * a.H);

  auto load_head = [&](const float* base, long long ld, long long row, const float* nw, const float* nb, int pos, float* v) {
    const float* p = base + row * ld + h * HD;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < EPL; ++j) { v[j] = p[lane + 32 * j]; s += v[j]; }
    if (nw) {
      const float mean = warp_sum(s) * (1.0f / HD);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < EPL; ++j) { const float d = v[j] - mean; q += d * d; }
      const float rstd = rsqrtf(warp_sum(q) * (1.0f / HD) + 1e-5f);
#pragma unroll
      for (int j = 0; j < EPL; ++j) v[j] = (v[j] - mean) * rstd * __ldg(nw + lane + 32 * j) + __ldg(nb + lane + 32 * j);
    }
    if (pos >= 0) {  // 1-D RoPE over the whole head: pairs (e, e + HD/2), angle = pos * base^(-2e/HD)
#pragma unroll
      for (int j = 0; j < EPL / 2; ++j) {
        const int e = lane + 32 * j;
        const float inv_freq = 1.0f / powf(a.rope_base, (float)(2 * e) / (float)HD);
        float sn, cs;
        sincosf((float)pos * inv_freq, &sn, &cs);
        const float x1 = v[j], x2 = v[j + EPL / 2];
        v[j] = x1 * cs - x2 * sn;
        v[j + EPL / 2] = x2 * cs + x1 * sn;
      }
    }
  };

  float q[EPL], o[EPL];
  load_head(a.q, a.ldq, (long long)bb * a.Nq + i, a.qn_w, a.qn_b, a.pos_q ? a.pos_q[i] : -1, q);
#pragma unroll
  for (int j = 0; j < EPL; ++j) { q[j] *= a.scale; o[j] = 0.f; }
  float m_run = -INFINITY, l_run = 0.f;
  for (int n = 0; n < a.Nk; ++n) {
    float k[EPL];
    load_head(a.k, a.ldk, (long long)bb * a.Nk + n, a.kn_w, a.kn_b, a.pos_k ? a.pos_k[n] : -1, k);
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < EPL; ++j) d = fmaf(q[j], k[j], d);
    d = warp_sum(d);
    const float m_new = fmaxf(m_run, d);
    const float alpha = expf(m_run - m_new), p = expf(d - m_new);
    l_run = l_run * alpha + p;
    const float* vp = a.v + ((long long)bb * a.Nk + n) * a.ldv + h * HD;
#pragma unroll
    for (int j = 0; j < EPL; ++j) o[j] = o[j] * alpha + p * vp[lane + 32 * j];
    m_run = m_new;
  }
  float* op = a.out + ((long long)bb * a.Nq + i) * a.ldo + h * HD;
  const float inv = 1.0f / l_run;
#pragma unroll
  for (int j = 0; j < EPL; ++j) op[lane + 32 * j] = o[j] * inv;
}

// ---- alignment decode helpers (D = 512 per token) ------------------------------------------------
// mean over the S rows of each batch element of the row L2 norm (alignment_head.py:469).  One block per batch.
__global__ void __launch_bounds__(256) mean_row_norm_kernel(const float* __restrict__ x, int S, int D, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();  // programmatic dependent launch (host_common.h)
  __shared__ float part[8];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int s = warp; s < S; s += 8) {
    const float* r = x + ((size_t)b * S + s) * D;
    float q = 0.f;
    for (int k = lane; k < D; k += 32) q = fmaf(r[k], r[k], q);
    q = warp_sum(q);
    if (lane == 0) acc += sqrtf(q);
  }
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i];
    out[b] = t / (float)S;
  }
}

// Build the chunk-cross key/value rows [tokens ; effective memory] and the directional memory.
// first chunk: frame_init (B, NM*D) = frame_proj(tokens[:,0]); directional = (1-s)*mem + s*normalize(frame_init), s = sigmoid(alpha)
//              effective = mem * mean_norm (the un-blended parameter)      (alignment_head.py:471-479)
// later      : directional = memory_in; effective = memory_in * mean_norm    (:484-485)
// One block per (batch, memory token), plus one block per copied token row (block y >= NM).
__global__ void __launch_bounds__(128) memory_prepare_kernel(const float* __restrict__ tokens, const float* __restrict__ mem_param,
                                                             const float* __restrict__ mem_in, const float* __restrict__ frame_init,
                                                             const float* __restrict__ alpha, const float* __restrict__ mean_norm,
                                                             float* __restrict__ kv, float* __restrict__ directional, int S, int NM,
                                                             int D) {
  pdl_launch_dependents();
  pdl_wait();  // programmatic dependent launch (host_common.h)
  __shared__ float red[4];
  const int b = blockIdx.x, j = blockIdx.y;
  if (j >= NM) {  // blocks NM .. NM + S - 1: one token row each (one block copying all S rows was a 59 us latency chain)
    const int r = j - NM;
    for (int i = threadIdx.x; i < D; i += blockDim.x) kv[((size_t)b * (S + NM) + r) * D + i] = tokens[((size_t)b * S + r) * D + i];
    return;
  }
  const float mn = mean_norm[b];
  const float* src = mem_in ? mem_in + ((size_t)b * NM + j) * D : mem_param + (size_t)j * D;
  float* kvrow = kv + ((size_t)b * (S + NM) + S + j) * D;
  float* drow = directional + ((size_t)b * NM + j) * D;
  if (mem_in) {
    for (int k = threadIdx.x; k < D; k += blockDim.x) { const float v = src[k]; kvrow[k] = v * mn; drow[k] = v; }
    return;
  }
  const float* fi = frame_init + ((size_t)b * NM + j) * D;
  float q = 0.f;
  for (int k = threadIdx.x; k < D; k += blockDim.x) q = fmaf(fi[k], fi[k], q);
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  const float nrm = fmaxf(sqrtf(red[0] + red[1] + red[2] + red[3]), 1e-6f);
  const float sg = sigmoidf(alpha[0]);
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    const float v = src[k];
    kvrow[k] = v * mn;
    drow[k] = (1.0f - sg) * v + sg * (fi[k] / nrm);
  }
}

// GatedUpdate stage 1 (gated_update.py:55-60): u = ||update||; inp[b,i] = [update, mem_i*u, mean_i(mem)*u]; also mem*u.
__global__ void __launch_bounds__(128) gu_prepare_kernel(const float* __restrict__ mem, const float* __restrict__ upd,
                                                         float* __restrict__ inp, float* __restrict__ mem_scaled, int NM, int D) {
  pdl_launch_dependents();
  pdl_wait();  // programmatic dependent launch (host_common.h)
  __shared__ float red[4];
  const int b = blockIdx.x, i = blockIdx.y;
  const float* u = upd + (size_t)b * D;
  float q = 0.f;
  for (int k = threadIdx.x; k < D; k += blockDim.x) q = fmaf(u[k], u[k], q);
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  const float un = sqrtf(red[0] + red[1] + red[2] + red[3]);
  float* row = inp + ((size_t)b * NM + i) * 3 * D;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float mean = 0.f;
    for (int j = 0; j < NM; ++j) mean += mem[((size_t)b * NM + j) * D + k];
    mean /= (float)NM;
    const float mi = mem[((size_t)b * NM + i) * D + k];
    row[k] = u[k];
    row[D + k] = mi * un;
    row[2 * D + k] = mean * un;
    mem_scaled[((size_t)b * NM + i) * D + k] = mi * un;
  }
}

// stage 2: gate_in[b,i] = [delta - mem, mem_scaled]  (:66-69)
__global__ void __launch_bounds__(128) gu_gate_input_kernel(const float* __restrict__ deltas, const float* __restrict__ mem,
                                                            const float* __restrict__ mem_scaled, float* __restrict__ gate_in, int D) {
  pdl_launch_dependents();
  pdl_wait();  // programmatic dependent launch (host_common.h)
  const size_t r = blockIdx.x;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    gate_in[r * 2 * D + k] = deltas[r * D + k] - mem[r * D + k];
    gate_in[r * 2 * D + D + k] = mem_scaled[r * D + k];
  }
}

// stage 3: orthogonalise the difference against the memory row, normalise, apply gate, normalise (:72-78)
__global__ void __launch_bounds__(128) gu_finish_kernel(const float* __restrict__ gate_in, const float* __restrict__ mem,
                                                        const float* __restrict__ gate, float* __restrict__ out, int D) {
  pdl_launch_dependents();
  pdl_wait();  // programmatic dependent launch (host_common.h)
  __shared__ float red[4];
  __shared__ float bc;
  const size_t r = blockIdx.x;
  const float* diff = gate_in + r * 2 * D;
  const float* m = mem + r * D;
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) bc = red[0] + red[1] + red[2] + red[3];
    __syncthreads();
    return bc;
  };
  float p = 0.f;
  for (int k = threadIdx.x; k < D; k += blockDim.x) p = fmaf(diff[k], m[k], p);
  const float dot = block_sum(p);
  float q = 0.f;
  for (int k = threadIdx.x; k < D; k += blockDim.x) { const float o = diff[k] - dot * m[k]; q = fmaf(o, o, q); }
  const float on = fmaxf(sqrtf(block_sum(q)), 1e-12f);  // F.normalize eps
  const float g = gate[r];
  float s = 0.f;
  for (int k = threadIdx.x; k < D; k += blockDim.x) { const float v = m[k] + g * ((diff[k] - dot * m[k]) / on); s = fmaf(v, v, s); }
  const float vn = fmaxf(sqrtf(block_sum(s)), 1e-12f);
  for (int k = threadIdx.x; k < D; k += blockDim.x) out[r * D + k] = (m[k] + g * ((diff[k] - dot * m[k]) / on)) / vn;
}

// camera head: x = gate * (adaLN(tok) * (1 + scale) + shift) + tok   with (shift, scale, gate) = chunk3(mod)
__global__ void __launch_bounds__(256) modulate_kernel(const float* __restrict__ normed, const float* __restrict__ tok,
                                                       const float* __restrict__ mod, float* __restrict__ out, long long rows, int D) {
  pdl_launch_dependents();
  pdl_wait();  // programmatic dependent launch (host_common.h)
  const long long total = rows * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D;
    const int c = (int)(i % D);
    const float sh = mod[r * 3 * D + c], sc = mod[r * 3 * D + D + c], g = mod[r * 3 * D + 2 * D + c];
    out[i] = g * (normed[i] * (1.0f + sc) + sh) + tok[i];
  }
}

// out[r, c] = a[r*lda + c] (+ b[r*ldb + c]) with optional column-wise activation: cols >= relu_from get relu,
// col == exp_col gets exp (pose_enc FoV relu; chunk_sim3 scale exp alignment_head.py:538).
__global__ void __launch_bounds__(256) combine_rows_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ b,
                                                           long long ldb, float* __restrict__ out, long long ldo, long long rows, int cols,
                                                           int relu_from, int exp_col) {
  pdl_launch_dependents();
  pdl_wait();  // programmatic dependent launch (host_common.h)
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i % cols);
    float v = a[r * lda + c] + (b ? b[r * ldb + c] : 0.f);
    if (relu_from >= 0 && c >= relu_from) v = fmaxf(v, 0.f);
    if (c == exp_col) v = expf(v);
    out[r * ldo + c] = v;
  }
}

int nblk(long long items) {
  long long b = (items + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 1184 ? 1184 : b));
}

}  // namespace

int linear_f32(const float* x, long long ldx, const float* W, const float* b, float* y, long long ldy, int M, int N, int K,
               int in_act, int out_act, const float* gamma, bool residual, cudaStream_t st) {
  LSVS_CHECK_ARG(x && W && y && M > 0 && N > 0 && K > 0, "linear_f32: bad arguments");
  ProfScope prof(PROF_SMALL_F32, st, 2.0 * M * (double)N * K, (double)N * K * 4);
  dim3 grid((N + 8 * LIN_NW - 1) / (8 * LIN_NW), (M + LIN_MT - 1) / LIN_MT);
  const bool vec = (K % 4 == 0) && (ldx % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)W % 16 == 0);
  if (vec && M > 8) {
    dim3 g2((N + 7) / 8, (M + 31) / 32);
    LSVS_CUDA(launch_pdl(linear_f32_tiled_kernel<4, 2>, dim3(g2), dim3(256), 0, st, x, ldx, W, b, y, ldy, M, N, K, in_act, out_act, gamma, residual ? 1 : 0));
  } else if (vec) {
    dim3 g2((N + 7) / 8, 1);
    LSVS_CUDA(launch_pdl(linear_f32_tiled_kernel<1, 8>, dim3(g2), dim3(256), 0, st, x, ldx, W, b, y, ldy, M, N, K, in_act, out_act, gamma, residual ? 1 : 0));
  } else {
    LSVS_CUDA(launch_pdl(linear_f32_kernel<false>, dim3(grid), dim3(256), 0, st, x, ldx, W, b, y, ldy, M, N, K, in_act, out_act, gamma, residual ? 1 : 0));
  }
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int attn_small_f32(const SmallAttnArgs& a, cudaStream_t st) {
  LSVS_CHECK_ARG(a.q && a.k && a.v && a.out && a.B > 0 && a.H > 0 && a.Nq > 0 && a.Nk > 0, "attn_small: bad arguments");
  LSVS_CHECK_ARG(a.hd == 64 || a.hd == 128, "attn_small: head_dim %d unsupported", a.hd);
  ProfScope prof(PROF_SMALL_F32, st, 0, 0);
  const int total = a.B * a.H * a.Nq;
  if (a.hd == 64) LSVS_CUDA(launch_pdl(attn_small_kernel<64>, dim3((total + 3) / 4), dim3(128), 0, st, a));
  else LSVS_CUDA(launch_pdl(attn_small_kernel<128>, dim3((total + 3) / 4), dim3(128), 0, st, a));
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int mean_row_norm(const float* x, int B, int S, int D, float* out, cudaStream_t st) {
  LSVS_CUDA(launch_pdl(mean_row_norm_kernel, dim3(B), dim3(256), 0, st, x, S, D, out));
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int memory_prepare(const float* tokens, const float* mem_param, const float* mem_in, const float* frame_init, const float* alpha,
                   const float* mean_norm, float* kv, float* directional, int B, int S, int NM, int D, cudaStream_t st) {
  LSVS_CUDA(launch_pdl(memory_prepare_kernel, dim3(dim3(B, NM + S)), dim3(128), 0, st, tokens, mem_param, mem_in, frame_init, alpha, mean_norm, kv, directional, S, NM, D));
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int gu_prepare(const float* mem, const float* upd, float* inp, float* mem_scaled, int B, int NM, int D, cudaStream_t st) {
  LSVS_CUDA(launch_pdl(gu_prepare_kernel, dim3(dim3(B, NM)), dim3(128), 0, st, mem, upd, inp, mem_scaled, NM, D));
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
int gu_gate_input(const float* deltas, const float* mem, const float* mem_scaled, float* gate_in, int rows, int D, cudaStream_t st) {
  LSVS_CUDA(launch_pdl(gu_gate_input_kernel, dim3(rows), dim3(128), 0, st, deltas, mem, mem_scaled, gate_in, D));
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
int gu_finish(const float* gate_in, const float* mem, const float* gate, float* out, int rows, int D, cudaStream_t st) {
  LSVS_CUDA(launch_pdl(gu_finish_kernel, dim3(rows), dim3(128), 0, st, gate_in, mem, gate, out, D));
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
int modulate(const float* normed, const float* tok, const float* mod, float* out, long long rows, int D, cudaStream_t st) {
  LSVS_CUDA(launch_pdl(modulate_kernel, dim3(nblk(rows * D)), dim3(256), 0, st, normed, tok, mod, out, rows, D));
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}
int combine_rows(const float* a, long long lda, const float* b, long long ldb, float* out, long long ldo, long long rows, int cols,
                 int relu_from, int exp_col, cudaStream_t st) {
  LSVS_CUDA(launch_pdl(combine_rows_kernel, dim3(nblk(rows * cols)), dim3(256), 0, st, a, lda, b, ldb, out, ldo, rows, cols, relu_from, exp_col));
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

}  // namespace lsvs
