// IRLS weighted Umeyama Sim(3) between two point maps, entirely on the device with no host synchronisation.
// Replaces irls_sim3_umeyama / weighted_umeyama_sim3, aligned_vggt/models/pointAligned_wrapped_vggt.py:159-305
// (median via torch.median, boolean compaction, <= 21 x [weighted centroid + 3x3 covariance + linalg.svd + det + .item()]).
//
//   w0_i  = sqrt(conf_src_i * conf_dst_i);   keep_i = w0_i >= factor * median(w0)          (:255-262)
//   solve 0: weights w0 (kept points)                                                      (:273)
//   solve k: weights w0_i * huber(|s R x_i + t - y_i|), huber(r) = r <= delta ? 1 : delta/r (:266-289)
//   stop when |dR|, |dt|, |ds| < tol, else after max_iters re-solves                        (:292-303)
// Excluded points get weight 0 instead of being compacted (same sums).  The exact lower median is found by a 4-pass
// radix select over the float bit patterns (all weights are >= 0, so integer order == float order).
// Per solve: pass A (weighted means) and pass B (centred 3x3 covariance + source variance), both grid reductions in
// double through one atomicAdd per block and quantity, then a single-thread 3x3 SVD (one-sided Jacobi in double).
// 28 B/point/pass of HBM (L2-resident at 638 k points).
#include "small_f32.h"
#include "host_common.h"

namespace lsvs {
namespace {

struct UmeyamaState {       // lives in device memory (workspace)
  double acc[20];           // [0] W, [1..3] sum w x, [4..6] sum w y, [7..15] sum w yc xc^T, [16] sum w |xc|^2
  float R[9], t[3], s;      // current estimate
  float thresh;             // factor * median
  int done;                 // convergence flag
  int iter;
  unsigned int hist[256];
  unsigned int prefix;      // radix-select state: bits decided so far
  unsigned int rank;        // remaining rank inside the current bucket
};

__device__ __forceinline__ float combined_w(const float* cs, const float* cd, long long i) { return sqrtf(cs[i] * cd[i]); }

// ---- radix select (lower median = element of rank (M-1)/2 in ascending order, as torch.median returns) ----------
__global__ void __launch_bounds__(256) hist_kernel(const float* __restrict__ cs, const float* __restrict__ cd, long long M,
                                                   UmeyamaState* st, int pass) {
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const unsigned int prefix = st->prefix;
  const unsigned int mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int bits = __float_as_uint(combined_w(cs, cd, i));
    if ((bits & mask) == prefix) atomicAdd(&h[(bits >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void select_kernel(UmeyamaState* st, int pass, float factor) {
  if (threadIdx.x != 0) return;
  const int shift = 24 - 8 * pass;
  unsigned int rank = st->rank, cum = 0;
  int b = 0;
  for (; b < 256; ++b) {
    const unsigned int c = st->hist[b];
    if (rank < cum + c) break;
    cum += c;
  }
  st->rank = rank - cum;
  st->prefix |= (unsigned int)b << shift;
  for (int i = 0; i < 256; ++i) st->hist[i] = 0;
  if (pass == 3) st->thresh = factor * __uint_as_float(st->prefix);
}

__global__ void init_state_kernel(UmeyamaState* st, long long M) {
  if (threadIdx.x != 0) return;
  for (int i = 0; i < 20; ++i) st->acc[i] = 0.0;
  for (int i = 0; i < 256; ++i) st->hist[i] = 0;
  st->prefix = 0; st->rank = (unsigned int)((M - 1) / 2); st->done = 0; st->iter = 0; st->s = 1.f;
  for (int i = 0; i < 9; ++i) st->R[i] = (i % 4 == 0) ? 1.f : 0.f;
  st->t[0] = st->t[1] = st->t[2] = 0.f;
}

// ---- weights -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float point_weight(const UmeyamaState* st, const float* cs, const float* cd, const float* x, const float* y,
                                              long long i, float delta, bool robust) {
  const float w0 = combined_w(cs, cd, i);
  if (!(w0 >= st->thresh)) return 0.f;
  if (!robust) return w0;
  const float px = x[3 * i], py = x[3 * i + 1], pz = x[3 * i + 2];
  const float* R = st->R;
  const float s = st->s;
  const float rx = s * (R[0] * px + R[1] * py + R[2] * pz) + st->t[0] - y[3 * i];
  const float ry = s * (R[3] * px + R[4] * py + R[5] * pz) + st->t[1] - y[3 * i + 1];
  const float rz = s * (R[6] * px + R[7] * py + R[8] * pz) + st->t[2] - y[3 * i + 2];
  const float r = sqrtf(rx * rx + ry * ry + rz * rz);
  return w0 * (r <= delta ? 1.0f : delta / fmaxf(r, 1e-12f));
}

template <int N>
__device__ __forceinline__ void block_reduce_add(double* v, double* dst) {
  __shared__ double sh[8][N];
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) sh[warp][k] = v[k];
  __syncthreads();
  if (threadIdx.x < N) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
    atomicAdd(dst + threadIdx.x, t);
  }
}

__global__ void __launch_bounds__(256) moments_a_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ cs,
                                                        const float* __restrict__ cd, long long M, UmeyamaState* st, float delta) {
  if (st->done) return;
  const bool robust = st->iter > 0;
  double v[7] = {0, 0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const float w = point_weight(st, cs, cd, x, y, i, delta, robust);
    if (w != 0.f) {
      v[0] += w;
      v[1] += (double)w * x[3 * i]; v[2] += (double)w * x[3 * i + 1]; v[3] += (double)w * x[3 * i + 2];
      v[4] += (double)w * y[3 * i]; v[5] += (double)w * y[3 * i + 1]; v[6] += (double)w * y[3 * i + 2];
    }
  }
  block_reduce_add<7>(v, st->acc);
}

__global__ void __launch_bounds__(256) moments_b_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ cs,
                                                        const float* __restrict__ cd, long long M, UmeyamaState* st, float delta) {
  if (st->done) return;
  const bool robust = st->iter > 0;
  const double W = st->acc[0];
  const float mx[3] = {(float)(st->acc[1] / W), (float)(st->acc[2] / W), (float)(st->acc[3] / W)};
  const float my[3] = {(float)(st->acc[4] / W), (float)(st->acc[5] / W), (float)(st->acc[6] / W)};
  double v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const float w = point_weight(st, cs, cd, x, y, i, delta, robust);
    if (w != 0.f) {
      const float xc[3] = {x[3 * i] - mx[0], x[3 * i + 1] - mx[1], x[3 * i + 2] - mx[2]};
      const float yc[3] = {y[3 * i] - my[0], y[3 * i + 1] - my[1], y[3 * i + 2] - my[2]};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) v[3 * a + b] += (double)(w * yc[a]) * xc[b];
      v[9] += (double)w * (xc[0] * xc[0] + xc[1] * xc[1] + xc[2] * xc[2]);
    }
  }
  block_reduce_add<10>(v, st->acc + 7);
}

// ---- 3x3 SVD by one-sided Jacobi (double): A = U diag(sv) V^T, sv descending ----------------------------------
__device__ void svd3(const double A[3][3], double U[3][3], double sv[3], double V[3][3]) {
  double B[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { B[i][j] = A[i][j]; V[i][j] = (i == j); }
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 3; ++k) { alpha += B[k][p] * B[k][p]; beta += B[k][q] * B[k][q]; gamma += B[k][p] * B[k][q]; }
        off += gamma * gamma;
        if (fabs(gamma) < 1e-300) continue;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) {
          const double bp = B[k][p], bq = B[k][q];
          B[k][p] = c * bp - s * bq; B[k][q] = s * bp + c * bq;
          const double vp = V[k][p], vq = V[k][q];
          V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
        }
      }
    if (off < 1e-60) break;
  }
  double n[3];
  for (int j = 0; j < 3; ++j) n[j] = sqrt(B[0][j] * B[0][j] + B[1][j] * B[1][j] + B[2][j] * B[2][j]);
  int order[3] = {0, 1, 2};
  for (int a = 0; a < 3; ++a) for (int b = a + 1; b < 3; ++b) if (n[order[b]] > n[order[a]]) { int tmp = order[a]; order[a] = order[b]; order[b] = tmp; }
  double Vs[3][3];
  for (int j = 0; j < 3; ++j) {
    const int o = order[j];
    sv[j] = n[o];
    for (int k = 0; k < 3; ++k) { Vs[k][j] = V[k][o]; U[k][j] = n[o] > 1e-300 ? B[k][o] / n[o] : 0.0; }
  }
  // complete a (numerically) rank-deficient U to an orthonormal basis
  if (sv[2] <= 1e-14 * sv[0]) {
    if (sv[1] <= 1e-14 * sv[0]) {  // rank <= 1: pick any unit vector orthogonal to u0
      const int k = fabs(U[0][0]) < 0.9 ? 0 : 1;
      double e[3] = {0, 0, 0}; e[k] = 1;
      double d = e[0] * U[0][0] + e[1] * U[1][0] + e[2] * U[2][0];
      double w[3] = {e[0] - d * U[0][0], e[1] - d * U[1][0], e[2] - d * U[2][0]};
      const double wn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
      for (int i = 0; i < 3; ++i) U[i][1] = w[i] / wn;
    }
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = Vs[i][j];
}

__device__ double det3(const double M[3][3]) {
  return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
         M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

__global__ void solve_kernel(UmeyamaState* st, float tol, int* status) {
  if (threadIdx.x != 0 || st->done) return;
  const double W = st->acc[0];
  if (!(W >= 1e-6)) {  // reference: ValueError("Total weight too small for meaningful estimation") (:184-185)
    *status = 1;
    st->done = 1;
    return;
  }
  double mux[3], muy[3], S[3][3];
  for (int k = 0; k < 3; ++k) { mux[k] = (double)(float)(st->acc[1 + k] / W); muy[k] = (double)(float)(st->acc[4 + k] / W); }
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) S[a][b] = st->acc[7 + 3 * a + b] / W;
  const double var_x = st->acc[16] / W;
  double U[3][3], sv[3], V[3][3], UVt[3][3];
  svd3(S, U, sv, V);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) UVt[i][j] = U[i][0] * V[j][0] + U[i][1] * V[j][1] + U[i][2] * V[j][2];
  const double d = det3(UVt) < 0 ? -1.0 : 1.0;
  float R[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[3 * i + j] = (float)(U[i][0] * V[j][0] + U[i][1] * V[j][1] + d * U[i][2] * V[j][2]);
  const float s = (float)((sv[0] + sv[1] + d * sv[2]) / var_x);
  float t[3];
  for (int i = 0; i < 3; ++i) t[i] = (float)(muy[i] - (double)s * (R[3 * i] * mux[0] + R[3 * i + 1] * mux[1] + R[3 * i + 2] * mux[2]));
  if (st->iter > 0) {
    double dR = 0, dt = 0;
    for (int i = 0; i < 9; ++i) dR += (double)(R[i] - st->R[i]) * (R[i] - st->R[i]);
    for (int i = 0; i < 3; ++i) dt += (double)(t[i] - st->t[i]) * (t[i] - st->t[i]);
    if (sqrt(dR) < tol && sqrt(dt) < tol && fabs((double)s - st->s) < tol) st->done = 1;
  }
  for (int i = 0; i < 9; ++i) st->R[i] = R[i];
  for (int i = 0; i < 3; ++i) st->t[i] = t[i];
  st->s = s;
  st->iter += 1;
  for (int i = 0; i < 20; ++i) st->acc[i] = 0.0;
}

__global__ void export_kernel(const UmeyamaState* st, float* R, float* t, float* s) {
  if (threadIdx.x < 9) R[threadIdx.x] = st->R[threadIdx.x];
  if (threadIdx.x < 3) t[threadIdx.x] = st->t[threadIdx.x];
  if (threadIdx.x == 0) s[0] = st->s;
}

}  // namespace

size_t irls_umeyama_workspace_bytes() { return sizeof(UmeyamaState) + 64; }

int irls_umeyama(const float* src, const float* dst, const float* conf_src, const float* conf_dst, long long M, float factor, float delta,
                 int max_iters, float tol, float* R, float* t, float* s, int* status, void* workspace, cudaStream_t stream) {
  LSVS_CHECK_ARG(src && dst && conf_src && conf_dst && R && t && s && status && workspace, "irls_umeyama: null pointer");
  LSVS_CHECK_ARG(M > 0 && max_iters >= 0, "irls_umeyama: empty point set");
  UmeyamaState* st = reinterpret_cast<UmeyamaState*>(workspace);
  const int blocks = (int)((M + 255) / 256 < (long long)num_sms() * 4 ? (M + 255) / 256 : (long long)num_sms() * 4);
  ProfScope prof(PROF_ELEMENTWISE, stream, 0, 28.0 * (double)M * 2 * (max_iters + 1));
  LSVS_CUDA(cudaMemsetAsync(status, 0, sizeof(int), stream));
  init_state_kernel<<<1, 32, 0, stream>>>(st, M);
  count_launch();
  for (int pass = 0; pass < 4; ++pass) {
    hist_kernel<<<blocks, 256, 0, stream>>>(conf_src, conf_dst, M, st, pass);
    select_kernel<<<1, 32, 0, stream>>>(st, pass, factor);
    count_launch(2);
  }
  for (int it = 0; it <= max_iters; ++it) {
    moments_a_kernel<<<blocks, 256, 0, stream>>>(src, dst, conf_src, conf_dst, M, st, delta);
    moments_b_kernel<<<blocks, 256, 0, stream>>>(src, dst, conf_src, conf_dst, M, st, delta);
    solve_kernel<<<1, 32, 0, stream>>>(st, tol, status);
    count_launch(3);
  }
  export_kernel<<<1, 32, 0, stream>>>(st, R, t, s);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

}  // namespace lsvs

extern "C" size_t lsvs_irls_umeyama_workspace_bytes(void) { return lsvs::irls_umeyama_workspace_bytes(); }

extern "C" int lsvs_irls_umeyama(const float* src, const float* dst, const float* conf_src, const float* conf_dst, long long n_points,
                                 float conf_threshold_factor, float delta, int max_iters, float tol, float* R, float* t, float* s,
                                 int* status, void* workspace, void* stream) {
  return lsvs::irls_umeyama(src, dst, conf_src, conf_dst, n_points, conf_threshold_factor, delta, max_iters, tol, R, t, s, status, workspace,
                            (cudaStream_t)stream);
}
