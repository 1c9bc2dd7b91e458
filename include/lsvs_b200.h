/*
 * lsvs_b200.h — C ABI of the B200-native chunk-encoder / feature-alignment hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The reference
 * (ruppelb/Large-Scale-ViT-SLAM) is pure PyTorch, so "what its FFI would bind" are the module
 * forwards / free functions on the hot path; each entry point cites the reference
 * interface it replaces.  The Python mirror of those interfaces (large-scale-vit-slam_b200/aligned_vggt)
 * binds this header through ctypes (large-scale-vit-slam_b200/_native.py); INTEGRATION.md shows the
 * reference-side stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - tensors are dense row-major; bf16 is the raw 16-bit pattern (uint16_t);
 *   - every call returns 0 on success, a negative LSVS_E* code otherwise; lsvs_last_error() gives text;
 *   - calls only enqueue work on `stream` (no host sync) unless stated.
 */
#ifndef LSVS_B200_H_
#define LSVS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSVS_OK 0
#define LSVS_EINVAL (-1)   /* bad argument / shape (the reference raises AssertionError / ValueError) */
#define LSVS_ECUDA (-2)    /* CUDA runtime / driver error */
#define LSVS_EUNSUPPORTED (-3)

typedef uint16_t lsvs_bf16;

/* ---- library ------------------------------------------------------------------------------- */
const char* lsvs_last_error(void);
int lsvs_version(void);
/* number of kernels this library has launched since load (monotonic; for bench.py's gpu_launches) */
unsigned long long lsvs_launch_count(void);

/* ---- Sim(3) apply (memory-bound) -------------------------------------------------------------
 * replaces apply_sim3_alignment_on_point_maps  aligned_vggt/utils/alignment.py:491-526
 *          and the inlined copies              aligned_vggt/models/featureAligned_vggt.py:198-207,
 *                                              aligned_vggt/models/poseAligned_wrapped_vggt.py:180-187
 * out[b,n,:] = T[b,:3,:3] * (s[b] * pts[b,n,:]) + T[b,:3,3]      pts/out: (B, n_points, 3) fp32
 * T: (B,4,4) fp32 row-major, s: (B) fp32.  out may alias pts.  24 B/point of HBM traffic. */
int lsvs_sim3_apply_points(const float* pts, const float* T, const float* s, float* out,
                           int batch, long long n_points, void* stream);
/* replaces `depth *= chunk_scale.view(B,1,1,1,1)`  featureAligned_vggt.py:171, alignment.py:487,
 * pointAligned_wrapped_vggt.py:137-138.   out[b,i] = s[b] * x[b,i];  8 B/pixel. */
int lsvs_scale_rows(const float* x, const float* s, float* out, int batch, long long n_per_batch, void* stream);
/* replaces apply_sim3_alignment_on_c2w alignment.py:558-594 : out[b,f] = T[b] @ [R | s[b]*t]  (B,S,4,4) */
int lsvs_sim3_apply_c2w(const float* poses, const float* T, const float* s, float* out, int batch, int frames, void* stream);
/* replaces apply_sim3_alignment_on_w2c alignment.py:528-556 : extr (B,S,rows,4), rows in {3,4} -> (B,S,4,4) */
int lsvs_sim3_apply_w2c(const float* extr, int rows, const float* T, const float* s, float* out, int batch, int frames, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LSVS_B200_H_ */
