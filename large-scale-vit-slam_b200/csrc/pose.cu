// Pose / Sim(3) composition of one chunk in a single launch (one block per batch element, one thread per frame).
// Follows featureAligned_vggt.py:97-143 (compose chunk Sim(3) x per-frame SE(3), re-base camera poses on frame 0,
// scale translations, mean overlap transform via Markley quaternion averaging, final 9-d pose encoding) and
// :190-196 (point-map transform).  Helpers follow data.py:12-52, geometry.py:4-37 and the upstream
// quat_to_mat / mat_to_quat / closed_form_inverse_se3 / pose_enc conventions (xyzw quaternions, w >= 0).
// The reference spends dozens of tiny ATen launches plus a cuSOLVER eigh here.
#include "small_f32.h"
#include "host_common.h"

namespace lsvs {
namespace {

struct M4 { float m[16]; };

__device__ void quat_to_R(const float* q, float* R, bool normalize) {  // xyzw
  float x = q[0], y = q[1], z = q[2], w = q[3];
  if (normalize) {
    const float n = fmaxf(sqrtf(x * x + y * y + z * z + w * w), 1e-8f);
    x /= n; y /= n; z /= n; w /= n;
  }
  const float two_s = 2.0f / (x * x + y * y + z * z + w * w);
  R[0] = 1 - two_s * (y * y + z * z); R[1] = two_s * (x * y - z * w); R[2] = two_s * (x * z + y * w);
  R[3] = two_s * (x * y + z * w); R[4] = 1 - two_s * (x * x + z * z); R[5] = two_s * (y * z - x * w);
  R[6] = two_s * (x * z - y * w); R[7] = two_s * (y * z + x * w); R[8] = 1 - two_s * (x * x + y * y);
}

__device__ M4 enc_to_mat(const float* enc, bool normalize) {  // [t(3), quat xyzw(4)] -> 4x4
  M4 o;
  float R[9];
  quat_to_R(enc + 3, R, normalize);
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) o.m[i * 4 + j] = R[i * 3 + j]; o.m[i * 4 + 3] = enc[i]; }
  o.m[12] = 0; o.m[13] = 0; o.m[14] = 0; o.m[15] = 1;
  return o;
}

__device__ M4 mul(const M4& a, const M4& b) {
  M4 o;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s = fmaf(a.m[i * 4 + k], b.m[k * 4 + j], s);
      o.m[i * 4 + j] = s;
    }
  return o;
}

__device__ M4 inv_se3(const M4& a) {
  M4 o;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o.m[i * 4 + j] = a.m[j * 4 + i];
  for (int i = 0; i < 3; ++i) o.m[i * 4 + 3] = -(o.m[i * 4] * a.m[3] + o.m[i * 4 + 1] * a.m[7] + o.m[i * 4 + 2] * a.m[11]);
  o.m[12] = 0; o.m[13] = 0; o.m[14] = 0; o.m[15] = 1;
  return o;
}

__device__ void mat_to_quat(const M4& a, float* q) {  // PyTorch3D matrix_to_quaternion, reordered to xyzw, w >= 0
  const float m00 = a.m[0], m01 = a.m[1], m02 = a.m[2], m10 = a.m[4], m11 = a.m[5], m12 = a.m[6], m20 = a.m[8], m21 = a.m[9], m22 = a.m[10];
  float qa[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22, 1.0f - m00 - m11 + m22};
  for (int i = 0; i < 4; ++i) qa[i] = qa[i] > 0 ? sqrtf(qa[i]) : 0.f;
  int best = 0;
  for (int i = 1; i < 4; ++i) if (qa[i] > qa[best]) best = i;
  float c[4];  // rijk candidate
  if (best == 0) { c[0] = qa[0] * qa[0]; c[1] = m21 - m12; c[2] = m02 - m20; c[3] = m10 - m01; }
  else if (best == 1) { c[0] = m21 - m12; c[1] = qa[1] * qa[1]; c[2] = m10 + m01; c[3] = m02 + m20; }
  else if (best == 2) { c[0] = m02 - m20; c[1] = m10 + m01; c[2] = qa[2] * qa[2]; c[3] = m12 + m21; }
  else { c[0] = m10 - m01; c[1] = m20 + m02; c[2] = m21 + m12; c[3] = qa[3] * qa[3]; }
  const float d = 2.0f * fmaxf(qa[best], 0.1f);
  float r = c[0] / d, i_ = c[1] / d, j_ = c[2] / d, k_ = c[3] / d;
  if (r < 0) { r = -r; i_ = -i_; j_ = -j_; k_ = -k_; }
  q[0] = i_; q[1] = j_; q[2] = k_; q[3] = r;
}

// largest-eigenvalue eigenvector of a symmetric 4x4 (cyclic Jacobi in double)
__device__ void top_eigvec4(double A[4][4], float* out) {
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 32; ++sweep) {
    double off = 0;
    for (int p = 0; p < 4; ++p) for (int q = p + 1; q < 4; ++q) off += A[p][q] * A[p][q];
    if (off < 1e-30) break;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 4; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 4; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i) if (A[i][i] > A[best][best]) best = i;
  double n = 0;
  for (int k = 0; k < 4; ++k) n += V[k][best] * V[k][best];
  n = sqrt(n);
  for (int k = 0; k < 4; ++k) out[k] = (float)(V[k][best] / n);
}

constexpr int MAX_OVERLAP = 128;

__device__ M4 load_gt(const float* g, int rows) {  // (rows,4) world-to-camera, rows 3 -> padded with [0 0 0 1] (poseAligned...:86-87)
  M4 o;
  for (int i = 0; i < rows * 4; ++i) o.m[i] = g[i];
  if (rows == 3) { o.m[12] = 0; o.m[13] = 0; o.m[14] = 0; o.m[15] = 1; }
  return o;
}

__global__ void __launch_bounds__(128) pose_chain_kernel(const float* __restrict__ chunk_sim3, const float* __restrict__ frame_se3,
                                                         const float* __restrict__ cam_enc, const float* __restrict__ prev_enc, int S_prev,
                                                         int overlap, int S, int H, int W, float* __restrict__ pose_out,
                                                         float* __restrict__ point_T, float* __restrict__ scale_out,
                                                         const float* __restrict__ gt, int gt_rows, int gt_mode) {
  __shared__ M4 chunk_se3, ident, mean_T, pf0;
  __shared__ float cam_enc7[MAX_OVERLAP][7];
  __shared__ float scale_sh;
  const int b = blockIdx.x;
  const float* cs = chunk_sim3 + (size_t)b * 8;
  const float* gtb = gt ? gt + (size_t)b * S * gt_rows * 4 : nullptr;
  if (threadIdx.x == 0) {
    chunk_se3 = enc_to_mat(cs, true);                                   // pose_encoding_to_extri (data.py:33-52)
    M4 e0 = enc_to_mat(cam_enc + (size_t)b * S * 9, false);              // upstream pose_encoding_to_extri_intri
    ident = inv_se3(e0);                                                 // featureAligned_vggt.py:114
    float sc = cs[7];
    if (gtb && (gt_mode & LSVS_GT_SCALE) && S > 1) {                     // poseAligned_wrapped_vggt.py:84-104 + scale_lse_solver
      const M4 centering = inv_se3(load_gt(gtb, gt_rows));               // (alignment.py:113-129): |sum x.y / sum x.x| over the
      double dots = 0, norms = 0;                                        // re-based predicted / first-frame-centred gt positions
      for (int s = 0; s < S; ++s) {
        const M4 g = mul(load_gt(gtb + (size_t)s * gt_rows * 4, gt_rows), centering);
        const M4 e = mul(enc_to_mat(cam_enc + ((size_t)b * S + s) * 9, false), ident);
        dots += (double)e.m[3] * g.m[3] + (double)e.m[7] * g.m[7] + (double)e.m[11] * g.m[11];
        norms += (double)e.m[3] * e.m[3] + (double)e.m[7] * e.m[7] + (double)e.m[11] * e.m[11];
      }
      sc *= (float)fabs(dots / norms);
    }
    scale_sh = sc;
    scale_out[b] = sc;
  }
  __syncthreads();
  const float scale = scale_sh;
  const bool gt_mean = gtb && (gt_mode & LSVS_GT_MEAN) && prev_enc;      // featureAligned_vggt.py:123-124, poseAligned...:108-109
  // re-based, scaled camera extrinsics of every frame (:116-119)
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    M4 e = mul(enc_to_mat(cam_enc + ((size_t)b * S + s) * 9, false), ident);
    e.m[3] *= scale; e.m[7] *= scale; e.m[11] *= scale;
    if (prev_enc && !gt_mean && s < overlap) {                           // :126-128
      const M4 ctx = enc_to_mat(prev_enc + ((size_t)b * S_prev + (S_prev - overlap + s)) * 9, true);
      const M4 ct = mul(inv_se3(e), ctx);
      if (overlap > 1) {                                                 // extri_to_pose_encoding (data.py:12-30)
        float q[4];
        mat_to_quat(ct, q);
        const float n = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-8f);
        cam_enc7[s][0] = ct.m[3]; cam_enc7[s][1] = ct.m[7]; cam_enc7[s][2] = ct.m[11];
        for (int k = 0; k < 4; ++k) cam_enc7[s][3 + k] = q[k] / n;
      } else {
        mean_T = ct;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!prev_enc) {
      for (int i = 0; i < 16; ++i) mean_T.m[i] = (i % 5 == 0) ? 1.f : 0.f;
    } else if (gt_mean) {
      mean_T = load_gt(gtb, gt_rows);
    } else if (overlap > 1) {                                            // averagePoseEncodings (geometry.py:4-37)
      float avg[7] = {0, 0, 0, 0, 0, 0, 0};
      double A[4][4] = {};
      for (int i = 0; i < overlap; ++i) {
        for (int k = 0; k < 3; ++k) avg[k] += cam_enc7[i][k];
        float q[4];
        float n = 0.f;
        for (int k = 0; k < 4; ++k) n += cam_enc7[i][3 + k] * cam_enc7[i][3 + k];
        n = fmaxf(sqrtf(n), 1e-8f);
        for (int k = 0; k < 4; ++k) q[k] = cam_enc7[i][3 + k] / n;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) A[r][c] += (double)(q[r] * q[c]) / overlap;
      }
      for (int k = 0; k < 3; ++k) avg[k] /= (float)overlap;
      top_eigvec4(A, avg + 3);
      mean_T = enc_to_mat(avg, true);
    }
    pf0 = mul(chunk_se3, mean_T);                                        // per_frame_se3[:,0] @ mean (:139)
    const M4 e0 = enc_to_mat(cam_enc + (size_t)b * S * 9, false);        // point_identity_alignment (:115)
    const M4 pt = prev_enc ? mul(inv_se3(pf0), e0) : e0;                 // :190-196
    for (int i = 0; i < 16; ++i) point_T[(size_t)b * 16 + i] = pt.m[i];
  }
  __syncthreads();
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float* ce = cam_enc + ((size_t)b * S + s) * 9;
    M4 e = mul(enc_to_mat(ce, false), ident);
    e.m[3] *= scale; e.m[7] *= scale; e.m[11] *= scale;
    M4 pf = (s == 0) ? pf0 : mul(mul(enc_to_mat(frame_se3 + ((size_t)b * (S - 1) + (s - 1)) * 7, true), chunk_se3), mean_T);
    const M4 al = mul(e, pf);                                            // :142
    float* o = pose_out + ((size_t)b * S + s) * 9;
    o[0] = al.m[3]; o[1] = al.m[7]; o[2] = al.m[11];
    mat_to_quat(al, o + 3);
    // extri_intri_to_pose_encoding on intrinsics rebuilt from the camera head's FoV (:109,:143)
    const float fy = (H / 2.0f) / tanf(ce[7] / 2.0f), fx = (W / 2.0f) / tanf(ce[8] / 2.0f);
    o[7] = 2.0f * atanf((H / 2.0f) / fy);
    o[8] = 2.0f * atanf((W / 2.0f) / fx);
  }
}

// pose_enc -> w2c extrinsics -> Sim(3) applied in camera-to-world space -> back to a pose encoding
// (pointAligned_wrapped_vggt.py:113-122 = pose_encoding_to_extri_intri -> apply_sim3_alignment_on_w2c, alignment.py:528-594
//  -> extri_intri_to_pose_encoding).  One thread per frame.
__global__ void pose_enc_sim3_kernel(const float* __restrict__ enc, const float* __restrict__ T, const float* __restrict__ s,
                                     float* __restrict__ out, int B, int S, int H, int W) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * S) return;
  const int b = idx / S;
  const float* e = enc + (size_t)idx * 9;
  M4 c2w = inv_se3(enc_to_mat(e, false));
  const float sc = s[b];
  c2w.m[3] *= sc; c2w.m[7] *= sc; c2w.m[11] *= sc;
  M4 Tb;
  for (int i = 0; i < 16; ++i) Tb.m[i] = T[(size_t)b * 16 + i];
  const M4 w2c = inv_se3(mul(Tb, c2w));
  float* o = out + (size_t)idx * 9;
  o[0] = w2c.m[3]; o[1] = w2c.m[7]; o[2] = w2c.m[11];
  mat_to_quat(w2c, o + 3);
  const float fy = (H / 2.0f) / tanf(e[7] / 2.0f), fx = (W / 2.0f) / tanf(e[8] / 2.0f);
  o[7] = 2.0f * atanf((H / 2.0f) / fy);
  o[8] = 2.0f * atanf((W / 2.0f) / fx);
}

}  // namespace

int pose_enc_apply_sim3(const float* enc, const float* T, const float* s, float* out, int B, int S, int H, int W, cudaStream_t st) {
  LSVS_CHECK_ARG(enc && T && s && out && B > 0 && S > 0, "pose_enc_apply_sim3: Inputs must have matching batch dimension");
  pose_enc_sim3_kernel<<<(B * S + 127) / 128, 128, 0, st>>>(enc, T, s, out, B, S, H, W);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

int pose_chain(const float* chunk_sim3, const float* frame_se3, const float* cam_enc, const float* prev_pose_enc, int S_prev,
               int overlap, int B, int S, int H, int W, float* pose_enc_out, float* point_T, float* scale_out, cudaStream_t st,
               const float* gt_poses, int gt_rows, int gt_mode) {
  LSVS_CHECK_ARG(chunk_sim3 && cam_enc && pose_enc_out && point_T && scale_out && B > 0 && S > 0, "pose_chain: bad arguments");
  LSVS_CHECK_ARG(S == 1 || frame_se3, "pose_chain: frame_se3 missing");
  const bool averaged = prev_pose_enc && !(gt_poses && (gt_mode & LSVS_GT_MEAN));
  LSVS_CHECK_ARG(!averaged || (overlap >= 1 && overlap <= MAX_OVERLAP && overlap <= S && overlap <= S_prev),
                 "pose_chain: overlap %d out of range (S=%d, previous chunk %d)", overlap, S, S_prev);
  LSVS_CHECK_ARG(!gt_poses || gt_rows == 3 || gt_rows == 4, "pose_chain: gt_poses must be (B,S,3,4) or (B,S,4,4)");
  pose_chain_kernel<<<B, 128, 0, st>>>(chunk_sim3, frame_se3, cam_enc, prev_pose_enc, S_prev, overlap, S, H, W, pose_enc_out, point_T,
                                       scale_out, gt_poses, gt_rows, gt_poses ? gt_mode : 0);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

}  // namespace lsvs
