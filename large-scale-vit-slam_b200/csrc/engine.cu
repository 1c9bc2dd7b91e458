// Native runtime of the hot path: owns the packed weights (bf16 for the tensor-core GEMMs, fp32 for norms /
// biases / the fp32 decode), the activation workspace, and sequences the kernels of
//   * Aggregator.forward        (UPSTREAM vggt; call site featureAligned_vggt.py:78)
//   * AlignmentHead.forward     (aligned_vggt/heads/alignment_head.py:224-345, decode :427-540)
//   * CameraHead.forward        (UPSTREAM vggt; call site featureAligned_vggt.py:106)
// on one CUDA stream with no host synchronisation.  The Python modules in large-scale-vit-slam_b200/aligned_vggt
// hold the nn.Parameters (state_dict contract) and push them here once (lsvs_engine_set_param).
#include <cuda_bf16.h>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "attention.h"
#include "dpt.h"
#include "elementwise.h"
#include "gemm.h"
#include "host_common.h"
#include "precise.h"
#include "small_f32.h"

namespace lsvs {
namespace {

#define TRY(expr) do { int rc_ = (expr); if (rc_ != LSVS_OK) return rc_; } while (0)

struct Param {
  float* f32 = nullptr;          // engine-owned fp32 copy (always kept for small tensors / fp32 path)
  __nv_bfloat16* bf16 = nullptr; // engine-owned bf16 copy of GEMM weights (K padded to a multiple of 64)
  long long numel = 0;
  int rows = 0, cols = 0, cols_padded = 0;
  bool split = false;            // bf16 copy holds the fp32-class split [hi | hi | lo] (3 * cols_padded wide), csrc/precise.cu
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return LSVS_OK;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    LSVS_CUDA(cudaMalloc(&p, need));
    bytes = need;
    return LSVS_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct BlockW {  // UPSTREAM Block / reference CrossAttentionBlock parameters
  const float *n1w, *n1b, *n2w, *n2b, *n3w = nullptr, *n3b = nullptr;
  const __nv_bfloat16 *qkv_w = nullptr, *q_w = nullptr, *kv_w = nullptr, *proj_w, *fc1_w, *fc2_w;
  const float *qkv_b = nullptr, *q_b = nullptr, *kv_b = nullptr, *proj_b, *fc1_b, *fc2_b;
  const float *qn_w = nullptr, *qn_b = nullptr, *kn_w = nullptr, *kn_b = nullptr;
  const float *ls1 = nullptr, *ls2 = nullptr;
  bool split = false;            // weights are split for the fp32-class path (engine precision mode)
};

struct BlockWF {  // fp32 block (camera trunk, decode cross blocks)
  const float *n1w, *n1b, *n2w, *n2b, *n3w = nullptr, *n3b = nullptr;
  const float *qkv_w = nullptr, *qkv_b = nullptr, *q_w = nullptr, *q_b = nullptr, *k_w = nullptr, *k_b = nullptr, *v_w = nullptr, *v_b = nullptr;
  const float *proj_w, *proj_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
  const float *qn_w = nullptr, *qn_b = nullptr, *kn_w = nullptr, *kn_b = nullptr, *ls1 = nullptr, *ls2 = nullptr;
};

struct RopeCfg { int mode = ROPE_NONE; const float2* tab = nullptr; int tpf = 0, nsp = 0, gw = 0; const int* ids = nullptr; int period = 0; };

}  // namespace

struct Engine {
  lsvs_engine_config cfg;
  std::map<std::string, Param> params;
  bool finalized = false;
  // derived weights
  std::vector<BlockW> dino, frame, global, h_frame, h_temporal, cam_trunk;
  std::vector<BlockWF> chunk_cross, frame_cross;
  std::vector<DevBuf> fused;  // concatenated k|v weights etc.
  // position embedding of the DINO ViT interpolated to the current patch grid (set by the host side)
  DevBuf pos_embed; int pos_gh = 0, pos_gw = 0;
  // tables
  DevBuf rope2d_64, rope2d_128, rope1d_128, ids_q, ids_k;
  long long ids_q_key = -1, ids_k_key = -1;   // (S, T, first chunk) the position ids on the device were built for: uploaded once per
                                               // configuration, so a steady-state forward issues no host-to-device copy (CUDA-graph capturable)
  int rope_npos2d = 0, rope_npos1d = 0;
  // workspace
  DevBuf x, xn, qkv, att, h, tmp, im2col, yn, kvb, scratch, dpt_ws;
  DevBuf p_xs, p_ys, p_qkv, p_kv, p_att, p_hf, p_hs;   // fp32-class path: split activations / fp32 GEMM outputs
  size_t scratch_off = 0;

  ~Engine() {
    for (auto& kv : params) { if (kv.second.f32) cudaFree(kv.second.f32); if (kv.second.bf16) cudaFree(kv.second.bf16); }
    for (auto& b : fused) b.release();
    for (DevBuf* b : {&pos_embed, &rope2d_64, &rope2d_128, &rope1d_128, &ids_q, &ids_k, &x, &xn, &qkv, &att, &h, &tmp, &im2col, &yn, &kvb, &scratch, &dpt_ws,
                       &p_xs, &p_ys, &p_qkv, &p_kv, &p_att, &p_hf, &p_hs}) b->release();
  }

  const Param* find(const std::string& n) const { auto it = params.find(n); return it == params.end() ? nullptr : &it->second; }
  bool scratch_overflow = false;
  float* scratch_f32(size_t n) {  // bump allocation inside `scratch` (caller sized it; overflow is reported, not faulted)
    const size_t need = (n + 63) & ~size_t(63);
    if ((scratch_off + need) * sizeof(float) > scratch.bytes) { scratch_overflow = true; return scratch.as<float>(); }
    float* p = scratch.as<float>() + scratch_off;
    scratch_off += need;
    return p;
  }
};

namespace {

bool ends_with(const std::string& s, const char* suf) {
  const size_t n = std::strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}
bool contains(const std::string& s, const char* sub) { return s.find(sub) != std::string::npos; }

// The two large Linears of the camera head's fp32 vector path (adaLN modulation 2048 -> 6144, pose branch fc1 2048 -> 1024, 32 rows):
// as CUDA-core fp32 GEMVs they were latency-bound weight streams (130 us / 35 us per launch, 0.4 TB/s); they now run fp32-class on
// the tensor cores in EVERY precision mode — split weights [hi | hi | lo] against split activations, 2^-16 relative.
bool camera_vector_gemm(const std::string& n) {
  return n == "camera_head.poseLN_modulation.1.weight" || n == "camera_head.pose_branch.fc1.weight";
}

// GEMM weights that run on the tensor cores in bf16 (everything the reference runs under bf16 autocast at scale).
bool wants_bf16(const std::string& n) {
  if (!ends_with(n, ".weight")) return false;
  if (camera_vector_gemm(n)) return true;
  if (contains(n, "camera_head.") && !contains(n, "camera_head.trunk.")) return false;
  const bool big_block = contains(n, "camera_head.trunk.") || contains(n, "aggregator.") || contains(n, "alignment_head.frame_blocks.") || contains(n, "alignment_head.temporal_blocks.");
  if (big_block && (contains(n, "attn.qkv.") || contains(n, "attn.proj.") || contains(n, "attn.q.") || contains(n, "attn.k.") ||
                    contains(n, "attn.v.") || contains(n, "mlp.fc1.") || contains(n, "mlp.fc2.") || contains(n, "patch_embed.proj.")))
    return true;
  if (contains(n, "depth_head.") || contains(n, "point_head."))  // DPT heads: every convolution except the last 1x1 (fp32 with the activation)
    return !ends_with(n, "norm.weight") && !contains(n, "output_conv2.2.");
  return n == "alignment_head.project_in.weight";
}

// precision mode 1: what the reference itself computes in fp32 (autocast disabled, featureAligned_vggt.py:103-104) runs fp32-class —
// the camera-head trunk and the DPT heads' convolutions — and so do the alignment head's blocks + project_in; mode 2: every
// block of the path.
bool wants_split(const std::string& n, int precision) {
  if (camera_vector_gemm(n)) return true;
  if (precision <= 0 || !wants_bf16(n)) return false;
  if (precision >= 2) return true;
  return contains(n, "alignment_head.") || contains(n, "camera_head.trunk.") || contains(n, "depth_head.") || contains(n, "point_head.");
}

int need(const Engine& e, const std::string& name, const Param** out) {
  const Param* p = e.find(name);
  if (!p) return fail(LSVS_EINVAL, "engine: parameter '%s' was never set (state_dict key missing)", name.c_str());
  *out = p;
  return LSVS_OK;
}
int need_f32(const Engine& e, const std::string& name, const float** out, long long numel = -1) {
  const Param* p;
  TRY(need(e, name, &p));
  if (!p->f32) return fail(LSVS_EINVAL, "engine: parameter '%s' has no fp32 copy", name.c_str());
  if (numel >= 0 && p->numel != numel) return fail(LSVS_EINVAL, "engine: parameter '%s' has %lld elements, expected %lld", name.c_str(), p->numel, numel);
  *out = p->f32;
  return LSVS_OK;
}
int need_bf16(const Engine& e, const std::string& name, const __nv_bfloat16** out, int rows, int cols, bool* split = nullptr) {
  const Param* p;
  TRY(need(e, name, &p));
  if (!p->bf16) return fail(LSVS_EINVAL, "engine: parameter '%s' has no bf16 copy", name.c_str());
  if (p->rows != rows || p->cols != cols) return fail(LSVS_EINVAL, "engine: parameter '%s' is %dx%d, expected %dx%d", name.c_str(), p->rows, p->cols, rows, cols);
  if (split) *split = p->split;
  else if (p->split) return fail(LSVS_EINVAL, "engine: parameter '%s' is stored split (precision mode) but its consumer has no fp32-class path", name.c_str());
  *out = p->bf16;
  return LSVS_OK;
}

int load_block(const Engine& e, const std::string& pre, int D, bool qk_norm, bool layer_scale, BlockW* w) {
  const int hd_unused = 0; (void)hd_unused;
  TRY(need_f32(e, pre + "norm1.weight", &w->n1w, D)); TRY(need_f32(e, pre + "norm1.bias", &w->n1b, D));
  TRY(need_f32(e, pre + "norm2.weight", &w->n2w, D)); TRY(need_f32(e, pre + "norm2.bias", &w->n2b, D));
  bool s1, s2, s3;
  TRY(need_bf16(e, pre + "attn.qkv.weight", &w->qkv_w, 3 * D, D, &w->split)); TRY(need_f32(e, pre + "attn.qkv.bias", &w->qkv_b, 3 * D));
  TRY(need_bf16(e, pre + "attn.proj.weight", &w->proj_w, D, D, &s1)); TRY(need_f32(e, pre + "attn.proj.bias", &w->proj_b, D));
  TRY(need_bf16(e, pre + "mlp.fc1.weight", &w->fc1_w, 4 * D, D, &s2)); TRY(need_f32(e, pre + "mlp.fc1.bias", &w->fc1_b, 4 * D));
  TRY(need_bf16(e, pre + "mlp.fc2.weight", &w->fc2_w, D, 4 * D, &s3)); TRY(need_f32(e, pre + "mlp.fc2.bias", &w->fc2_b, D));
  if (s1 != w->split || s2 != w->split || s3 != w->split) return fail(LSVS_EINVAL, "engine: block '%s' mixes split and plain weights", pre.c_str());
  if (qk_norm) {
    TRY(need_f32(e, pre + "attn.q_norm.weight", &w->qn_w)); TRY(need_f32(e, pre + "attn.q_norm.bias", &w->qn_b));
    TRY(need_f32(e, pre + "attn.k_norm.weight", &w->kn_w)); TRY(need_f32(e, pre + "attn.k_norm.bias", &w->kn_b));
  }
  if (layer_scale) { TRY(need_f32(e, pre + "ls1.gamma", &w->ls1, D)); TRY(need_f32(e, pre + "ls2.gamma", &w->ls2, D)); }
  return LSVS_OK;
}

int load_block_f32(const Engine& e, const std::string& pre, int D, bool cross, bool qk_norm, BlockWF* w) {
  TRY(need_f32(e, pre + "norm1.weight", &w->n1w, D)); TRY(need_f32(e, pre + "norm1.bias", &w->n1b, D));
  TRY(need_f32(e, pre + "norm2.weight", &w->n2w, D)); TRY(need_f32(e, pre + "norm2.bias", &w->n2b, D));
  if (cross) {
    TRY(need_f32(e, pre + "norm3.weight", &w->n3w, D)); TRY(need_f32(e, pre + "norm3.bias", &w->n3b, D));
    TRY(need_f32(e, pre + "attn.q.weight", &w->q_w, (long long)D * D)); TRY(need_f32(e, pre + "attn.q.bias", &w->q_b, D));
    TRY(need_f32(e, pre + "attn.k.weight", &w->k_w, (long long)D * D)); TRY(need_f32(e, pre + "attn.k.bias", &w->k_b, D));
    TRY(need_f32(e, pre + "attn.v.weight", &w->v_w, (long long)D * D)); TRY(need_f32(e, pre + "attn.v.bias", &w->v_b, D));
  } else {
    TRY(need_f32(e, pre + "attn.qkv.weight", &w->qkv_w, 3LL * D * D)); TRY(need_f32(e, pre + "attn.qkv.bias", &w->qkv_b, 3 * D));
  }
  TRY(need_f32(e, pre + "attn.proj.weight", &w->proj_w, (long long)D * D)); TRY(need_f32(e, pre + "attn.proj.bias", &w->proj_b, D));
  TRY(need_f32(e, pre + "mlp.fc1.weight", &w->fc1_w, 4LL * D * D)); TRY(need_f32(e, pre + "mlp.fc1.bias", &w->fc1_b, 4 * D));
  TRY(need_f32(e, pre + "mlp.fc2.weight", &w->fc2_w, 4LL * D * D)); TRY(need_f32(e, pre + "mlp.fc2.bias", &w->fc2_b, D));
  if (qk_norm) {
    TRY(need_f32(e, pre + "attn.q_norm.weight", &w->qn_w)); TRY(need_f32(e, pre + "attn.q_norm.bias", &w->qn_b));
    TRY(need_f32(e, pre + "attn.k_norm.weight", &w->kn_w)); TRY(need_f32(e, pre + "attn.k_norm.bias", &w->kn_b));
  }
  TRY(need_f32(e, pre + "ls1.gamma", &w->ls1, D)); TRY(need_f32(e, pre + "ls2.gamma", &w->ls2, D));
  return LSVS_OK;
}

// ------------------------------------------------------------------------------------------------
// bf16 transformer block on the fp32 residual stream x (M rows of D=1024).
// fp32-class MLP half shared by the self- and cross-attention blocks: x += ls2 * fc2(gelu(fc1(LN(x))))  (csrc/precise.cu)
int run_mlp_precise(Engine& e, float* x, long long M, const BlockW& w, float eps, int D, float* tap, int tap_ld, cudaStream_t st) {
  TRY(e.p_xs.ensure((size_t)M * 3 * D * 2)); TRY(e.p_hf.ensure((size_t)M * 4 * D * 4)); TRY(e.p_hs.ensure((size_t)M * 12 * D * 2));
  TRY(layernorm_split(x, D, w.n2w, w.n2b, eps, e.p_xs.p, 3 * D, M, D, st));
  GemmEpilogue e1; e1.bias = w.fc1_b; e1.out = e.p_hf.p; e1.ldo = 4 * D;
  TRY(gemm_bf16(e.p_xs.p, 3 * D, w.fc1_w, 3 * D, (int)M, 4 * D, 3 * D, EPI_BIAS_F32, e1, st));
  TRY(cast_split(e.p_hf.as<float>(), 4 * D, e.p_hs.p, 12 * D, M, 4 * D, true, st));
  GemmEpilogue e2; e2.bias = w.fc2_b; e2.gamma = w.ls2; e2.resid = x; e2.ldr = D; e2.out2 = tap; e2.ld2 = tap_ld;
  return gemm_bf16(e.p_hs.p, 12 * D, w.fc2_w, 12 * D, (int)M, D, 12 * D, EPI_RESID_F32, e2, st);
}

// fp32-class transformer block: same structure as run_block below, GEMM operands split into bf16 pairs, attention in fp32
int run_block_precise(Engine& e, float* x, long long M, const BlockW& w, float eps, int heads, int hd, int attn_batches, int L,
                      const RopeCfg& rope, float* tap, int tap_ld, cudaStream_t st) {
  const int D = heads * hd;
  TRY(e.p_xs.ensure((size_t)M * 3 * D * 2)); TRY(e.p_qkv.ensure((size_t)M * 3 * D * 4)); TRY(e.p_att.ensure((size_t)M * 3 * D * 2));
  float* qkv = e.p_qkv.as<float>();
  TRY(layernorm_split(x, D, w.n1w, w.n1b, eps, e.p_xs.p, 3 * D, M, D, st));
  GemmEpilogue ep; ep.bias = w.qkv_b; ep.out = qkv; ep.ldo = 3 * D;
  TRY(gemm_bf16(e.p_xs.p, 3 * D, w.qkv_w, 3 * D, (int)M, 3 * D, 3 * D, EPI_BIAS_F32, ep, st));
  if (w.qn_w) {
    TRY(headnorm_rope_f32(qkv, 3 * D, M, 0, heads, hd, w.qn_w, w.qn_b, 1e-5f, rope.mode, rope.tab, rope.tpf, rope.nsp, rope.gw, rope.ids, rope.period, st));
    TRY(headnorm_rope_f32(qkv, 3 * D, M, D, heads, hd, w.kn_w, w.kn_b, 1e-5f, rope.mode, rope.tab, rope.tpf, rope.nsp, rope.gw, rope.ids, rope.period, st));
  }
  AttentionF32Args aa{qkv, qkv + D, qkv + 2 * D, e.p_att.p, 3 * D, 3 * D, 3 * D, 3 * D, D, attn_batches, heads, hd, L, L, 1.0f / sqrtf((float)hd)};
  TRY(attention_f32(aa, st));
  GemmEpilogue er; er.bias = w.proj_b; er.gamma = w.ls1; er.resid = x; er.ldr = D;
  TRY(gemm_bf16(e.p_att.p, 3 * D, w.proj_w, 3 * D, (int)M, D, 3 * D, EPI_RESID_F32, er, st));
  return run_mlp_precise(e, x, M, w, eps, D, tap, tap_ld, st);
}

int run_block(Engine& e, float* x, long long M, const BlockW& w, float eps, int heads, int hd, int attn_batches, int L,
              const RopeCfg& rope, float* tap, int tap_ld, cudaStream_t st) {
  if (w.split) return run_block_precise(e, x, M, w, eps, heads, hd, attn_batches, L, rope, tap, tap_ld, st);
  const int D = heads * hd;
  __nv_bfloat16 *xn = e.xn.as<__nv_bfloat16>(), *qkv = e.qkv.as<__nv_bfloat16>(), *att = e.att.as<__nv_bfloat16>(), *h = e.h.as<__nv_bfloat16>();
  TRY(layernorm(x, D, RowMap{}, w.n1w, w.n1b, eps, xn, D, RowMap{}, true, M, D, st));
  GemmEpilogue ep;
  ep.bias = w.qkv_b; ep.out = qkv; ep.ldo = 3 * D;
  int kind = EPI_BIAS_BF16;
  if (w.qn_w) {
    kind = hd == 64 ? EPI_HEADNORM64_BF16 : EPI_HEADNORM128_BF16;
    ep.qn_w = w.qn_w; ep.qn_b = w.qn_b; ep.kn_w = w.kn_w; ep.kn_b = w.kn_b; ep.n_q_cols = D; ep.n_k_cols = D; ep.ln_eps = 1e-5f;
    ep.rope_mode = rope.mode; ep.rope_tab = rope.tab; ep.tokens_per_frame = rope.tpf; ep.n_special = rope.nsp; ep.grid_w = rope.gw;
  }
  TRY(gemm_bf16(xn, D, w.qkv_w, D, (int)M, 3 * D, D, kind, ep, st));
  AttentionArgs aa{qkv, qkv + D, qkv + 2 * D, att, 3 * D, 3 * D, 3 * D, D, attn_batches, heads, hd, L, L, 1.0f / sqrtf((float)hd)};
  TRY(attention_fwd(aa, st));
  GemmEpilogue er;
  er.bias = w.proj_b; er.gamma = w.ls1; er.resid = x; er.ldr = D;
  TRY(gemm_bf16(att, D, w.proj_w, D, (int)M, D, D, EPI_RESID_F32, er, st));
  TRY(layernorm(x, D, RowMap{}, w.n2w, w.n2b, eps, xn, D, RowMap{}, true, M, D, st));
  GemmEpilogue e1;
  e1.bias = w.fc1_b; e1.out = h; e1.ldo = 4 * D;
  TRY(gemm_bf16(xn, D, w.fc1_w, D, (int)M, 4 * D, D, EPI_BIAS_GELU_BF16, e1, st));
  GemmEpilogue e2;
  e2.bias = w.fc2_b; e2.gamma = w.ls2; e2.resid = x; e2.ldr = D; e2.out2 = tap; e2.ld2 = tap_ld;
  TRY(gemm_bf16(h, 4 * D, w.fc2_w, 4 * D, (int)M, D, 4 * D, EPI_RESID_F32, e2, st));
  return LSVS_OK;
}

int ensure_tables(Engine& e, int gh, int gw, int n1d, cudaStream_t st) {
  const int n2d = (gh > gw ? gh : gw) + 2;
  if (n2d > e.rope_npos2d) {
    TRY(e.rope2d_64.ensure((size_t)n2d * 16 * 8)); TRY(e.rope2d_128.ensure((size_t)n2d * 32 * 8));
    TRY(lsvs_rope_table(e.rope2d_64.as<float>(), n2d, 16, e.cfg.rope_base, st));
    TRY(lsvs_rope_table(e.rope2d_128.as<float>(), n2d, 32, e.cfg.rope_base, st));
    e.rope_npos2d = n2d;
  }
  if (n1d > e.rope_npos1d) {
    TRY(e.rope1d_128.ensure((size_t)n1d * 64 * 8));
    TRY(lsvs_rope_table(e.rope1d_128.as<float>(), n1d, 64, e.cfg.rope_base, st));
    e.rope_npos1d = n1d;
  }
  return LSVS_OK;
}

int ensure_workspace(Engine& e, long long rows, long long patch_rows) {
  TRY(e.x.ensure((size_t)rows * 1024 * 4)); TRY(e.xn.ensure((size_t)rows * 1024 * 2)); TRY(e.qkv.ensure((size_t)rows * 3072 * 2));
  TRY(e.att.ensure((size_t)rows * 1024 * 2)); TRY(e.h.ensure((size_t)rows * 4096 * 2)); TRY(e.tmp.ensure((size_t)rows * 1024 * 4));
  if (patch_rows) TRY(e.im2col.ensure((size_t)patch_rows * 640 * 2));
  return LSVS_OK;
}

}  // namespace
}  // namespace lsvs

using lsvs::Engine;
using namespace lsvs;

extern "C" int lsvs_engine_create(const lsvs_engine_config* cfg, lsvs_engine** out) {
  LSVS_CHECK_ARG(cfg && out, "engine_create: null argument");
  LSVS_CHECK_ARG(cfg->embed_dim == 1024 && cfg->num_heads == 16 && cfg->patch_size == 14 && cfg->num_register_tokens == 4,
                 "engine_create: only the VGGT-1B geometry (dim 1024, 16 heads, patch 14, 4 registers) is built");
  LSVS_CHECK_ARG(cfg->depth >= 0 && cfg->dino_depth >= 0 && cfg->head_depth_aa >= 0, "engine_create: bad depth");
  Engine* e = new Engine();
  e->cfg = *cfg;
  *out = reinterpret_cast<lsvs_engine*>(e);
  return LSVS_OK;
}

extern "C" int lsvs_engine_destroy(lsvs_engine* h) {
  delete reinterpret_cast<Engine*>(h);
  return LSVS_OK;
}

extern "C" int lsvs_engine_set_param(lsvs_engine* h, const char* name, const float* data, long long numel, int rows, int cols, void* stream) {
  Engine& e = *reinterpret_cast<Engine*>(h);
  LSVS_CHECK_ARG(name && data && numel > 0, "engine_set_param: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const std::string n(name);
  Param& p = e.params[n];
  const bool as_bf16 = wants_bf16(n);
  if (as_bf16) {
    LSVS_CHECK_ARG(rows > 0 && cols > 0 && (long long)rows * cols == numel, "engine_set_param: '%s' needs a 2-D (rows, cols) view", name);
    const int kp = (cols + 63) / 64 * 64;
    const bool split = wants_split(n, e.cfg.precision);
    const int copies = split ? 3 : 1;
    if (p.bf16 && ((long long)p.rows * p.cols_padded * (p.split ? 3 : 1) != (long long)rows * kp * copies)) { cudaFree(p.bf16); p.bf16 = nullptr; }
    if (!p.bf16) LSVS_CUDA(cudaMalloc(&p.bf16, (size_t)rows * kp * 2 * copies));
    if (split) TRY(pack_weight_split(data, p.bf16, rows, cols, kp, st));
    else TRY(pack_weight_bf16(data, p.bf16, rows, cols, kp, st));
    p.rows = rows; p.cols = cols; p.cols_padded = kp; p.split = split;
    if (p.f32) { cudaFree(p.f32); p.f32 = nullptr; }
  } else {
    if (p.f32 && p.numel != numel) { cudaFree(p.f32); p.f32 = nullptr; }
    if (!p.f32) LSVS_CUDA(cudaMalloc(&p.f32, (size_t)numel * 4));
    LSVS_CUDA(cudaMemcpyAsync(p.f32, data, (size_t)numel * 4, cudaMemcpyDeviceToDevice, st));
    p.rows = rows; p.cols = cols; p.cols_padded = cols;
  }
  p.numel = numel;
  e.finalized = false;
  return LSVS_OK;
}

extern "C" int lsvs_engine_finalize(lsvs_engine* h, void* stream) {
  Engine& e = *reinterpret_cast<Engine*>(h);
  cudaStream_t st = (cudaStream_t)stream;
  const int D = 1024;
  e.dino.assign(e.cfg.dino_depth, BlockW{}); e.frame.assign(e.cfg.depth, BlockW{}); e.global.assign(e.cfg.depth, BlockW{});
  for (int i = 0; i < e.cfg.dino_depth; ++i) TRY(load_block(e, "aggregator.patch_embed.blocks." + std::to_string(i) + ".", D, false, true, &e.dino[i]));
  for (int i = 0; i < e.cfg.depth; ++i) {
    TRY(load_block(e, "aggregator.frame_blocks." + std::to_string(i) + ".", D, true, true, &e.frame[i]));
    TRY(load_block(e, "aggregator.global_blocks." + std::to_string(i) + ".", D, true, true, &e.global[i]));
  }
  for (auto& b : e.fused) b.release();
  e.fused.clear();
  e.h_frame.clear(); e.h_temporal.clear(); e.chunk_cross.clear(); e.frame_cross.clear(); e.cam_trunk.clear();
  if (e.cfg.with_alignment_head) {
    e.h_frame.assign(e.cfg.head_depth_aa, BlockW{}); e.h_temporal.assign(e.cfg.head_depth_aa, BlockW{});
    e.fused.reserve(2 * e.cfg.head_depth_aa);
    for (int i = 0; i < e.cfg.head_depth_aa; ++i) {
      TRY(load_block(e, "alignment_head.frame_blocks." + std::to_string(i) + ".", D, true, true, &e.h_frame[i]));
      const std::string pre = "alignment_head.temporal_blocks." + std::to_string(i) + ".";
      BlockW& w = e.h_temporal[i];
      TRY(need_f32(e, pre + "norm1.weight", &w.n1w, D)); TRY(need_f32(e, pre + "norm1.bias", &w.n1b, D));
      TRY(need_f32(e, pre + "norm2.weight", &w.n2w, D)); TRY(need_f32(e, pre + "norm2.bias", &w.n2b, D));
      TRY(need_f32(e, pre + "norm3.weight", &w.n3w, D)); TRY(need_f32(e, pre + "norm3.bias", &w.n3b, D));
      bool sk, sv, sp, s1, s2;
      TRY(need_bf16(e, pre + "attn.q.weight", &w.q_w, D, D, &w.split)); TRY(need_f32(e, pre + "attn.q.bias", &w.q_b, D));
      // k and v read the same input: fuse into one (2D, D) weight so one GEMM produces [k | v]
      const __nv_bfloat16 *kw, *vw; const float *kb, *vb;
      TRY(need_bf16(e, pre + "attn.k.weight", &kw, D, D, &sk)); TRY(need_bf16(e, pre + "attn.v.weight", &vw, D, D, &sv));
      TRY(need_f32(e, pre + "attn.k.bias", &kb, D)); TRY(need_f32(e, pre + "attn.v.bias", &vb, D));
      const size_t wbytes = (size_t)D * D * 2 * (w.split ? 3 : 1);   // split weights: rows are 3*D wide
      e.fused.emplace_back(); DevBuf& fw = e.fused.back(); TRY(fw.ensure(2 * wbytes));
      e.fused.emplace_back(); DevBuf& fb = e.fused.back(); TRY(fb.ensure((size_t)2 * D * 4));
      LSVS_CUDA(cudaMemcpyAsync(fw.p, kw, wbytes, cudaMemcpyDeviceToDevice, st));
      LSVS_CUDA(cudaMemcpyAsync(fw.as<uint8_t>() + wbytes, vw, wbytes, cudaMemcpyDeviceToDevice, st));
      LSVS_CUDA(cudaMemcpyAsync(fb.p, kb, (size_t)D * 4, cudaMemcpyDeviceToDevice, st));
      LSVS_CUDA(cudaMemcpyAsync(fb.as<float>() + D, vb, (size_t)D * 4, cudaMemcpyDeviceToDevice, st));
      w.kv_w = fw.as<__nv_bfloat16>(); w.kv_b = fb.as<float>();
      TRY(need_bf16(e, pre + "attn.proj.weight", &w.proj_w, D, D, &sp)); TRY(need_f32(e, pre + "attn.proj.bias", &w.proj_b, D));
      TRY(need_bf16(e, pre + "mlp.fc1.weight", &w.fc1_w, 4 * D, D, &s1)); TRY(need_f32(e, pre + "mlp.fc1.bias", &w.fc1_b, 4 * D));
      TRY(need_bf16(e, pre + "mlp.fc2.weight", &w.fc2_w, D, 4 * D, &s2)); TRY(need_f32(e, pre + "mlp.fc2.bias", &w.fc2_b, D));
      if (sk != w.split || sv != w.split || sp != w.split || s1 != w.split || s2 != w.split)
        return fail(LSVS_EINVAL, "engine: block '%s' mixes split and plain weights", pre.c_str());
      TRY(need_f32(e, pre + "attn.q_norm.weight", &w.qn_w, 128)); TRY(need_f32(e, pre + "attn.q_norm.bias", &w.qn_b, 128));
      TRY(need_f32(e, pre + "attn.k_norm.weight", &w.kn_w, 128)); TRY(need_f32(e, pre + "attn.k_norm.bias", &w.kn_b, 128));
      TRY(need_f32(e, pre + "ls1.gamma", &w.ls1, D)); TRY(need_f32(e, pre + "ls2.gamma", &w.ls2, D));
    }
    e.chunk_cross.assign(2, BlockWF{}); e.frame_cross.assign(2, BlockWF{});
    for (int i = 0; i < 2; ++i) {
      TRY(load_block_f32(e, "alignment_head.chunk_cross_blocks." + std::to_string(i) + ".", 512, true, true, &e.chunk_cross[i]));
      TRY(load_block_f32(e, "alignment_head.frame_cross_blocks." + std::to_string(i) + ".", 512, true, true, &e.frame_cross[i]));
    }
  }
  if (e.cfg.with_camera_head) {
    e.cam_trunk.assign(4, BlockW{});
    for (int i = 0; i < 4; ++i) TRY(load_block(e, "camera_head.trunk." + std::to_string(i) + ".", 2048, false, true, &e.cam_trunk[i]));
  }
  e.finalized = true;
  return LSVS_OK;
}

extern "C" int lsvs_engine_set_pos_embed(lsvs_engine* h, const float* pos, int gh, int gw, void* stream) {
  Engine& e = *reinterpret_cast<Engine*>(h);
  LSVS_CHECK_ARG(pos && gh > 0 && gw > 0, "engine_set_pos_embed: bad arguments");
  const size_t bytes = (size_t)(1 + gh * gw) * 1024 * 4;
  TRY(e.pos_embed.ensure(bytes));
  LSVS_CUDA(cudaMemcpyAsync(e.pos_embed.p, pos, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  e.pos_gh = gh; e.pos_gw = gw;
  return LSVS_OK;
}

// ================================================================================================ Aggregator
extern "C" int lsvs_aggregator_forward(lsvs_engine* h, const float* images, int B, int S, int H, int W, float* const* taps,
                                       const int* tap_layers, int n_taps, void* stream) {
  Engine& e = *reinterpret_cast<Engine*>(h);
  cudaStream_t st = (cudaStream_t)stream;
  LSVS_CHECK_ARG(e.finalized && e.cfg.depth > 0 && e.cfg.dino_depth > 0, "aggregator_forward: engine has no aggregator / not finalized (call lsvs_engine_finalize after setting parameters)");
  LSVS_CHECK_ARG(images && B > 0 && S > 0 && H > 0 && W > 0 && H % 14 == 0 && W % 14 == 0, "aggregator_forward: images must be (B,S,3,H,W) with H,W multiples of 14");
  LSVS_CHECK_ARG(n_taps >= 0 && (n_taps == 0 || (taps && tap_layers)), "aggregator_forward: tap list missing");
  const int gh = H / 14, gw = W / 14, Pp = gh * gw, P = Pp + 5, frames = B * S, D = 1024;
  LSVS_CHECK_ARG(e.pos_gh == gh && e.pos_gw == gw, "aggregator_forward: position embedding not set for a %dx%d patch grid", gh, gw);
  const long long M = (long long)frames * P;
  TRY(ensure_workspace(e, M, (long long)frames * Pp));
  TRY(ensure_tables(e, gh, gw, 0, st));
  float* x = e.x.as<float>();
  // --- DINOv2 ViT-L/14 patch embedding (A.2)
  const Param* pw; const float *pb, *cls, *reg, *nw, *nb, *cam, *regtok;
  TRY(need(e, "aggregator.patch_embed.patch_embed.proj.weight", &pw));
  if (pw->split) {  // fp32-class: fp32 im2col, split operands (K = 3 * 640)
    TRY(e.p_hf.ensure((size_t)frames * Pp * 640 * 4)); TRY(e.p_xs.ensure((size_t)frames * Pp * 1920 * 2));
    TRY(patch_unfold_f32(images, e.p_hf.as<float>(), frames, H, W, st));
    TRY(cast_split(e.p_hf.as<float>(), 640, e.p_xs.p, 1920, (long long)frames * Pp, 640, false, st));
  } else {
    TRY(patch_unfold(images, e.im2col.p, frames, H, W, st));
  }
  LSVS_CHECK_ARG(pw->bf16 && pw->rows == 1024 && pw->cols == 588, "aggregator: patch_embed.proj.weight must be (1024, 3*14*14)");
  TRY(need_f32(e, "aggregator.patch_embed.patch_embed.proj.bias", &pb, D));
  TRY(need_f32(e, "aggregator.patch_embed.cls_token", &cls, D)); TRY(need_f32(e, "aggregator.patch_embed.register_tokens", &reg, 4 * D));
  TRY(need_f32(e, "aggregator.patch_embed.norm.weight", &nw, D)); TRY(need_f32(e, "aggregator.patch_embed.norm.bias", &nb, D));
  TRY(need_f32(e, "aggregator.camera_token", &cam, 2 * D)); TRY(need_f32(e, "aggregator.register_token", &regtok, 8 * D));
  GemmEpilogue ep;
  ep.bias = pb; ep.out = e.tmp.p; ep.ldo = D;
  if (pw->split) TRY(gemm_bf16(e.p_xs.p, 1920, pw->bf16, 1920, frames * Pp, D, 1920, EPI_BIAS_F32, ep, st));
  else TRY(gemm_bf16(e.im2col.p, 640, pw->bf16, 640, frames * Pp, D, 640, EPI_BIAS_F32, ep, st));
  TRY(dino_assemble(e.tmp.as<float>(), cls, reg, e.pos_embed.as<float>(), x, frames, Pp, 4, D, st));
  RopeCfg none;
  for (int i = 0; i < e.cfg.dino_depth; ++i) TRY(run_block(e, x, M, e.dino[i], 1e-6f, 16, 64, frames, P, none, nullptr, 0, st));
  // final norm on the patch rows (x_norm_patchtokens), then camera / register tokens replace cls / DINO registers (A.1)
  RowMap pm{Pp, P, 5};
  TRY(layernorm(x, D, pm, nw, nb, 1e-6f, x, D, pm, false, (long long)frames * Pp, D, st));
  TRY(fill_special(cam, x, frames, S, P, 0, 1, D, st));
  TRY(fill_special(regtok, x, frames, S, P, 1, 4, D, st));
  // --- alternating frame / global attention
  RopeCfg rope; rope.mode = ROPE_2D; rope.tab = e.rope2d_64.as<float2>(); rope.tpf = P; rope.nsp = 5; rope.gw = gw;
  for (int i = 0; i < e.cfg.depth; ++i) {
    float* tap = nullptr;
    for (int t = 0; t < n_taps; ++t) if (tap_layers[t] == i) tap = taps[t];
    TRY(run_block(e, x, M, e.frame[i], 1e-5f, 16, 64, frames, P, rope, tap, 2 * D, st));
    TRY(run_block(e, x, M, e.global[i], 1e-5f, 16, 64, B, S * P, rope, tap ? tap + D : nullptr, 2 * D, st));
    // a layer requested several times (reduced-depth test configs) shares one computation
    for (int t = 0, first = -1; t < n_taps; ++t) {
      if (tap_layers[t] != i) continue;
      if (first < 0) { first = t; continue; }
      if (taps[t] != taps[first]) LSVS_CUDA(cudaMemcpyAsync(taps[t], taps[first], (size_t)M * 2 * D * 4, cudaMemcpyDeviceToDevice, st));
    }
  }
  return LSVS_OK;
}

// ================================================================================================ fp32 helpers
namespace lsvs {
namespace {

// fp32 cross-attention block on few rows (decode): x (B*Nq, D) updated in place, y (B*Nk, D).
int run_cross_block_f32(Engine& e, float* x, int B, int Nq, const float* y, int Nk, const BlockWF& w, int D, int heads,
                        const int* pos_q, const int* pos_k, cudaStream_t st) {
  const int Mq = B * Nq, Mk = B * Nk;
  float* xn = e.scratch_f32((size_t)Mq * D); float* yn = e.scratch_f32((size_t)Mk * D);
  float* q = e.scratch_f32((size_t)Mq * D); float* k = e.scratch_f32((size_t)Mk * D); float* v = e.scratch_f32((size_t)Mk * D);
  float* att = e.scratch_f32((size_t)Mq * D); float* hb = e.scratch_f32((size_t)Mq * 4 * D);
  TRY(layernorm(x, D, RowMap{}, w.n1w, w.n1b, 1e-5f, xn, D, RowMap{}, false, Mq, D, st));
  TRY(layernorm(y, D, RowMap{}, w.n3w, w.n3b, 1e-5f, yn, D, RowMap{}, false, Mk, D, st));
  TRY(linear_f32(xn, D, w.q_w, w.q_b, q, D, Mq, D, D, ACT_NONE, ACT_NONE, nullptr, false, st));
  TRY(linear_f32(yn, D, w.k_w, w.k_b, k, D, Mk, D, D, ACT_NONE, ACT_NONE, nullptr, false, st));
  TRY(linear_f32(yn, D, w.v_w, w.v_b, v, D, Mk, D, D, ACT_NONE, ACT_NONE, nullptr, false, st));
  SmallAttnArgs a{q, D, k, D, v, D, att, D, B, heads, D / heads, Nq, Nk, w.qn_w, w.qn_b, w.kn_w, w.kn_b, pos_q, pos_k, e.cfg.rope_base,
                  1.0f / sqrtf((float)(D / heads))};
  TRY(attn_small_f32(a, st));
  TRY(linear_f32(att, D, w.proj_w, w.proj_b, x, D, Mq, D, D, ACT_NONE, ACT_NONE, w.ls1, true, st));
  TRY(layernorm(x, D, RowMap{}, w.n2w, w.n2b, 1e-5f, xn, D, RowMap{}, false, Mq, D, st));
  TRY(linear_f32(xn, D, w.fc1_w, w.fc1_b, hb, 4 * D, Mq, 4 * D, D, ACT_NONE, ACT_GELU, nullptr, false, st));
  TRY(linear_f32(hb, 4 * D, w.fc2_w, w.fc2_b, x, D, Mq, D, 4 * D, ACT_NONE, ACT_NONE, w.ls2, true, st));
  return LSVS_OK;
}

int upload_ids(DevBuf& buf, const std::vector<int>& ids, cudaStream_t st) {
  TRY(buf.ensure(ids.size() * sizeof(int) + 256));
  LSVS_CUDA(cudaMemcpyAsync(buf.p, ids.data(), ids.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  return LSVS_OK;
}

}  // namespace
}  // namespace lsvs


// fp32 decode of the per-frame alignment tokens (alignment_head.py:427-540): `align_tok` row f (stride `ld_tok`)
// is the alignment token of frame f.  d_dec: device ids [0..S-1, 2S..2S+NM-1] (decode RoPE positions, :446-455).
namespace lsvs {
namespace {
int decode_forward(Engine& e, const float* x, long long ld_tok, int B, int S, const int* d_dec, const float* memory_in,
                   float* chunk_sim3, float* frame_se3, float* memory_out, cudaStream_t st) {
  const int D = 1024, DD = 512, NM = e.cfg.num_memory_tokens, frames = B * S;
  // ------------------------------------------------------------------ fp32 decode (:427-540)
  const size_t need_scratch = (size_t)B * ((size_t)(S + NM) * DD * 16 + (size_t)S * DD * 24 + (size_t)NM * DD * 16 + 8192) + (1 << 16);
  TRY(e.scratch.ensure(need_scratch * 4));
  e.scratch_off = 0;
  const float *pdw, *pdb, *dnw, *dnb, *memp = nullptr, *alpha = nullptr, *fpw = nullptr, *fpb = nullptr, *cnw, *cnb, *fnw, *fnb;
  TRY(need_f32(e, "alignment_head.project_dec.weight", &pdw, (long long)DD * D)); TRY(need_f32(e, "alignment_head.project_dec.bias", &pdb, DD));
  TRY(need_f32(e, "alignment_head.dec_norm.weight", &dnw, DD)); TRY(need_f32(e, "alignment_head.dec_norm.bias", &dnb, DD));
  if (NM > 0) {  // num_memory_tokens = 0 (alignment_head.py:211,468,504): no memory parameters, keys = the frame tokens only
    TRY(need_f32(e, "alignment_head.memory_token", &memp, (long long)NM * DD)); TRY(need_f32(e, "alignment_head.alpha", &alpha, 1));
    TRY(need_f32(e, "alignment_head.frame_proj.weight", &fpw, (long long)NM * DD * DD)); TRY(need_f32(e, "alignment_head.frame_proj.bias", &fpb, NM * DD));
  }
  TRY(need_f32(e, "alignment_head.chunk_norm.weight", &cnw, DD)); TRY(need_f32(e, "alignment_head.chunk_norm.bias", &cnb, DD));
  TRY(need_f32(e, "alignment_head.frame_norm.weight", &fnw, DD)); TRY(need_f32(e, "alignment_head.frame_norm.bias", &fnb, DD));
  float* t0 = e.scratch_f32((size_t)frames * DD); float* tok = e.scratch_f32((size_t)frames * DD);
  // per-frame alignment token = row 0 of every frame: stride P1*D
  TRY(linear_f32(x, ld_tok, pdw, pdb, t0, DD, frames, DD, D, ACT_NONE, ACT_NONE, nullptr, false, st));
  TRY(layernorm(t0, DD, RowMap{}, dnw, dnb, 1e-5f, tok, DD, RowMap{}, false, frames, DD, st));
  float* kvt = tok; float* directional = nullptr;
  if (NM > 0) {
    float* mean_norm = e.scratch_f32(B);
    TRY(mean_row_norm(tok, B, S, DD, mean_norm, st));
    float* frame_init = nullptr;
    if (!memory_in) {
      frame_init = e.scratch_f32((size_t)B * NM * DD);
      TRY(linear_f32(tok, (long long)S * DD, fpw, fpb, frame_init, (long long)NM * DD, B, NM * DD, DD, ACT_NONE, ACT_NONE, nullptr, false, st));
    }
    kvt = e.scratch_f32((size_t)B * (S + NM) * DD); directional = e.scratch_f32((size_t)B * NM * DD);
    TRY(memory_prepare(tok, memp, memory_in, frame_init, alpha, mean_norm, kvt, directional, B, S, NM, DD, st));
  }
  // chunk token: frame-0 token attends over [all frame tokens ; scaled memory]
  float* chunk_tok = e.scratch_f32((size_t)B * DD);
  LSVS_CUDA(cudaMemcpy2DAsync(chunk_tok, DD * 4, tok, (size_t)S * DD * 4, DD * 4, B, cudaMemcpyDeviceToDevice, st));
  for (int i = 0; i < 2; ++i) TRY(run_cross_block_f32(e, chunk_tok, B, 1, kvt, S + NM, e.chunk_cross[i], DD, 8, d_dec /*[0]*/, d_dec, st));
  // gated memory update (gated_update.py:43-78)
  if (NM > 0) {
    float* inp = e.scratch_f32((size_t)B * NM * 3 * DD); float* mem_scaled = e.scratch_f32((size_t)B * NM * DD);
    float* hid = e.scratch_f32((size_t)B * NM * DD); float* deltas = e.scratch_f32((size_t)B * NM * DD);
    float* gate_in = e.scratch_f32((size_t)B * NM * 2 * DD); float* ghid = e.scratch_f32((size_t)B * NM * DD); float* gate = e.scratch_f32((size_t)B * NM);
    TRY(gu_prepare(directional, chunk_tok, inp, mem_scaled, B, NM, DD, st));
    for (int i = 0; i < NM; ++i) {
      const std::string pre = "alignment_head.gated_update.delta_mlps." + std::to_string(i) + ".";
      const float *w0, *b0, *w2, *b2;
      TRY(need_f32(e, pre + "0.weight", &w0, 3LL * DD * DD)); TRY(need_f32(e, pre + "0.bias", &b0, DD));
      TRY(need_f32(e, pre + "2.weight", &w2, (long long)DD * DD)); TRY(need_f32(e, pre + "2.bias", &b2, DD));
      TRY(linear_f32(inp + (size_t)i * 3 * DD, (long long)NM * 3 * DD, w0, b0, hid + (size_t)i * DD, (long long)NM * DD, B, DD, 3 * DD, ACT_NONE, ACT_GELU, nullptr, false, st));
      TRY(linear_f32(hid + (size_t)i * DD, (long long)NM * DD, w2, b2, deltas + (size_t)i * DD, (long long)NM * DD, B, DD, DD, ACT_NONE, ACT_NONE, nullptr, false, st));
    }
    const float *g0w, *g0b, *g2w, *g2b;
    TRY(need_f32(e, "alignment_head.gated_update.gate_mlp.0.weight", &g0w, 2LL * DD * DD)); TRY(need_f32(e, "alignment_head.gated_update.gate_mlp.0.bias", &g0b, DD));
    TRY(need_f32(e, "alignment_head.gated_update.gate_mlp.2.weight", &g2w, DD)); TRY(need_f32(e, "alignment_head.gated_update.gate_mlp.2.bias", &g2b, 1));
    TRY(gu_gate_input(deltas, directional, mem_scaled, gate_in, B * NM, DD, st));
    TRY(linear_f32(gate_in, 2 * DD, g0w, g0b, ghid, DD, B * NM, DD, 2 * DD, ACT_NONE, ACT_GELU, nullptr, false, st));
    TRY(linear_f32(ghid, DD, g2w, g2b, gate, 1, B * NM, 1, DD, ACT_NONE, ACT_SIGMOID, nullptr, false, st));
    TRY(gu_finish(gate_in, directional, gate, memory_out, B * NM, DD, st));
  }
  float* chunk_n = e.scratch_f32((size_t)B * DD);
  TRY(layernorm(chunk_tok, DD, RowMap{}, cnw, cnb, 1e-5f, chunk_n, DD, RowMap{}, false, B, DD, st));
  const float *c1w, *c1b, *c2w, *c2b, *f1w, *f1b, *f2w, *f2b;
  TRY(need_f32(e, "alignment_head.chunk_sim3_decoder.fc1.weight", &c1w, 256LL * DD)); TRY(need_f32(e, "alignment_head.chunk_sim3_decoder.fc1.bias", &c1b, 256));
  TRY(need_f32(e, "alignment_head.chunk_sim3_decoder.fc2.weight", &c2w, 8LL * 256)); TRY(need_f32(e, "alignment_head.chunk_sim3_decoder.fc2.bias", &c2b, 8));
  TRY(need_f32(e, "alignment_head.frame_se3_decoder.fc1.weight", &f1w, 256LL * DD)); TRY(need_f32(e, "alignment_head.frame_se3_decoder.fc1.bias", &f1b, 256));
  TRY(need_f32(e, "alignment_head.frame_se3_decoder.fc2.weight", &f2w, 7LL * 256)); TRY(need_f32(e, "alignment_head.frame_se3_decoder.fc2.bias", &f2b, 7));
  if (S > 1) {  // per-frame tokens (frames 1..S-1) attend to the normed chunk token (:510-534)
    float* ft = e.scratch_f32((size_t)B * (S - 1) * DD);
    LSVS_CUDA(cudaMemcpy2DAsync(ft, (size_t)(S - 1) * DD * 4, tok + DD, (size_t)S * DD * 4, (size_t)(S - 1) * DD * 4, B, cudaMemcpyDeviceToDevice, st));
    for (int i = 0; i < 2; ++i) TRY(run_cross_block_f32(e, ft, B, S - 1, chunk_n, 1, e.frame_cross[i], DD, 8, d_dec + 1, d_dec, st));
    float* ftn = e.scratch_f32((size_t)B * (S - 1) * DD); float* fh = e.scratch_f32((size_t)B * (S - 1) * 256);
    TRY(layernorm(ft, DD, RowMap{}, fnw, fnb, 1e-5f, ftn, DD, RowMap{}, false, (long long)B * (S - 1), DD, st));
    TRY(linear_f32(ftn, DD, f1w, f1b, fh, 256, B * (S - 1), 256, DD, ACT_NONE, ACT_GELU, nullptr, false, st));
    TRY(linear_f32(fh, 256, f2w, f2b, frame_se3, 7, B * (S - 1), 7, 256, ACT_NONE, ACT_NONE, nullptr, false, st));
  }
  float* ch = e.scratch_f32((size_t)B * 256); float* c8 = e.scratch_f32((size_t)B * 8);
  TRY(linear_f32(chunk_n, DD, c1w, c1b, ch, 256, B, 256, DD, ACT_NONE, ACT_GELU, nullptr, false, st));
  TRY(linear_f32(ch, 256, c2w, c2b, c8, 8, B, 8, 256, ACT_NONE, ACT_NONE, nullptr, false, st));
  TRY(combine_rows(c8, 8, nullptr, 0, chunk_sim3, 8, B, 8, -1, 7, st));  // exp on the scale entry (:538)
  if (e.scratch_overflow) { e.scratch_overflow = false; return fail(LSVS_ECUDA, "alignment decode: scratch under-sized"); }
  return LSVS_OK;
}
}  // namespace
}  // namespace lsvs

// ================================================================================================ AlignmentHead
namespace lsvs {
namespace {
// prefix_out != NULL: stop after the context-free prefix (project_in, token_norm, alignment token, first frame block) and copy the
// fp32 token stream (B,S,P+1,1024) there.  prefix_in != NULL: start from such a stream instead of `tokens`.
int alignment_head_forward_impl(lsvs_engine* h, const float* tokens, const __nv_bfloat16* tokens_bf16, int B, int S, int P, int H, int W,
                                int next_overlap, const float* overlap_in, int T, const float* memory_in, float* chunk_sim3,
                                float* frame_se3, float* memory_out, float* overlap_out, void* stream, float* prefix_out = nullptr,
                                const float* prefix_in = nullptr);
}  // namespace
}  // namespace lsvs

extern "C" int lsvs_alignment_head_forward(lsvs_engine* h, const float* tokens, int B, int S, int P, int H, int W, int next_overlap,
                                           const float* overlap_in, int T, const float* memory_in, float* chunk_sim3,
                                           float* frame_se3, float* memory_out, float* overlap_out, void* stream) {
  LSVS_CHECK_ARG(tokens, "alignment_head_forward: bad arguments");
  return lsvs::alignment_head_forward_impl(h, tokens, nullptr, B, S, P, H, W, next_overlap, overlap_in, T, memory_in, chunk_sim3, frame_se3,
                                           memory_out, overlap_out, stream);
}

extern "C" int lsvs_alignment_head_forward_bf16(lsvs_engine* h, const lsvs_bf16* tokens, int B, int S, int P, int H, int W, int next_overlap,
                                                const float* overlap_in, int T, const float* memory_in, float* chunk_sim3,
                                                float* frame_se3, float* memory_out, float* overlap_out, void* stream) {
  LSVS_CHECK_ARG(tokens, "alignment_head_forward_bf16: bad arguments");
  return lsvs::alignment_head_forward_impl(h, nullptr, reinterpret_cast<const __nv_bfloat16*>(tokens), B, S, P, H, W, next_overlap, overlap_in, T,
                                           memory_in, chunk_sim3, frame_se3, memory_out, overlap_out, stream);
}

extern "C" int lsvs_alignment_head_prefix(lsvs_engine* h, const float* tokens, int B, int S, int P, int H, int W, float* prefix_out,
                                          void* stream) {
  LSVS_CHECK_ARG(tokens && prefix_out, "alignment_head_prefix: bad arguments");
  return lsvs::alignment_head_forward_impl(h, tokens, nullptr, B, S, P, H, W, 0, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, stream,
                                           prefix_out, nullptr);
}

extern "C" int lsvs_alignment_head_resume(lsvs_engine* h, const float* prefix, int B, int S, int P, int H, int W, int next_overlap,
                                          const float* overlap_in, int T, const float* memory_in, float* chunk_sim3, float* frame_se3,
                                          float* memory_out, float* overlap_out, void* stream) {
  LSVS_CHECK_ARG(prefix, "alignment_head_resume: bad arguments");
  return lsvs::alignment_head_forward_impl(h, nullptr, nullptr, B, S, P, H, W, next_overlap, overlap_in, T, memory_in, chunk_sim3, frame_se3,
                                           memory_out, overlap_out, stream, nullptr, prefix);
}

namespace lsvs {
namespace {
int alignment_head_forward_impl(lsvs_engine* h, const float* tokens, const __nv_bfloat16* tokens_bf16, int B, int S, int P, int H, int W,
                                int next_overlap, const float* overlap_in, int T, const float* memory_in, float* chunk_sim3,
                                float* frame_se3, float* memory_out, float* overlap_out, void* stream, float* prefix_out,
                                const float* prefix_in) {
  Engine& e = *reinterpret_cast<Engine*>(h);
  cudaStream_t st = (cudaStream_t)stream;
  LSVS_CHECK_ARG(e.finalized && e.cfg.with_alignment_head, "alignment_head_forward: engine has no alignment head / not finalized");
  LSVS_CHECK_ARG(B > 0 && S > 0 && P > 5, "alignment_head_forward: bad arguments");
  if (!prefix_out) {
    LSVS_CHECK_ARG(chunk_sim3 && overlap_out, "alignment_head_forward: bad arguments");
    LSVS_CHECK_ARG(memory_out || e.cfg.num_memory_tokens == 0, "alignment_head_forward: memory_out missing");
    LSVS_CHECK_ARG(S == 1 || frame_se3, "alignment_head_forward: frame_se3 output missing");
  }
  LSVS_CHECK_ARG(e.cfg.head_depth_aa >= 1, "alignment_head_forward: head has no blocks");
  const int gh = H / 14, gw = W / 14, D = 1024, DD = 512, NM = e.cfg.num_memory_tokens, P1 = P + 1, frames = B * S;
  LSVS_CHECK_ARG(gh * gw + 5 == P, "Size of tokens and image do not match (P=%d, grid %dx%d)", P, gh, gw);
  LSVS_CHECK_ARG(NM == 8 || NM == 0, "alignment_head_forward: num_memory_tokens must be 8 or 0");
  LSVS_CHECK_ARG(!overlap_in || T >= 2, "Size of tokens and overlap tokens must match");
  LSVS_CHECK_ARG(next_overlap >= 0 && next_overlap <= S, "alignment_head_forward: next_num_overlap out of range");
  const bool first = overlap_in == nullptr;
  if (first) T = S;
  LSVS_CHECK_ARG(T - 1 <= S, "alignment_head_forward: more overlap tokens than frames");
  const long long M = (long long)frames * P, Mh = (long long)frames * P1, My = (long long)B * T * P1;
  TRY(ensure_workspace(e, Mh > My ? Mh : My, 0));
  TRY(e.yn.ensure((size_t)My * D * 2)); TRY(e.kvb.ensure((size_t)My * 2 * D * 2));
  TRY(ensure_tables(e, gh, gw, 2 * S + NM + 2, st));
  // temporal position ids (alignment_head.py:279-285)
  std::vector<int> qi(S), ki(T);
  for (int j = 0; j < S; ++j) qi[j] = first ? j : j + (S - (T - 1));
  if (first) for (int j = 0; j < T; ++j) ki[j] = j;
  else { ki[0] = 0; for (int j = 1; j < T; ++j) ki[j] = S - (T - 1) + (j - 1); }
  // decode ids (:446-455): frame queries 1..S-1 vs key 0; chunk query 0 vs keys 0..S-1, then S+i+S
  std::vector<int> dec(S + (S + NM) + 1);
  for (int j = 0; j < S; ++j) dec[j] = j;                        // [0,S): 0..S-1  (dec+1 = frame query ids; dec = chunk key ids)
  for (int j = 0; j < NM; ++j) dec[S + j] = 2 * S + j;           // memory keys
  std::vector<int> all(qi); all.insert(all.end(), ki.begin(), ki.end()); all.insert(all.end(), dec.begin(), dec.end());
  const long long ids_key = (((long long)S * 4096 + T) * 2 + (first ? 1 : 0)) * 64 + NM;
  if (e.ids_q_key != ids_key) {
    e.ids_q_key = -1;
    TRY(upload_ids(e.ids_q, all, st));
    e.ids_q_key = ids_key;
  }
  const int* d_qi = e.ids_q.as<int>(); const int* d_ki = d_qi + S; const int* d_dec = d_ki + T;

  float* x = e.x.as<float>();
  __nv_bfloat16* xn = e.xn.as<__nv_bfloat16>();
  const Param* pin; const float *pin_b, *tnw, *tnb, *atok;
  TRY(need(e, "alignment_head.project_in.weight", &pin));
  LSVS_CHECK_ARG(pin->bf16 && pin->rows == D && pin->cols == 2 * D, "alignment_head.project_in.weight must be (1024, 2048)");
  TRY(need_f32(e, "alignment_head.project_in.bias", &pin_b, D));
  TRY(need_f32(e, "alignment_head.token_norm.weight", &tnw, D)); TRY(need_f32(e, "alignment_head.token_norm.bias", &tnb, D));
  TRY(need_f32(e, "alignment_head.per_frame_alignment_token", &atok, 2 * D));
  // project_in + token_norm, written behind the per-frame alignment token (:242-270)
  GemmEpilogue ep; ep.bias = pin_b; ep.out = e.tmp.p; ep.ldo = D;
  if (prefix_in) {   // the owner of the chunk already ran the context-free prefix: its token stream is the starting point
    LSVS_CUDA(cudaMemcpyAsync(x, prefix_in, (size_t)Mh * D * 4, cudaMemcpyDeviceToDevice, st));
  } else if (tokens_bf16) {  // tokens that already are bf16 (the chunk scheduler ships them that way): they are the GEMM operand as they stand
    LSVS_CHECK_ARG(!pin->split, "alignment_head_forward_bf16: a precision-mode engine needs fp32 tokens");
    TRY(gemm_bf16(tokens_bf16, 2 * D, pin->bf16, 2 * D, (int)M, D, 2 * D, EPI_BIAS_F32, ep, st));
  } else if (pin->split) {
    TRY(e.p_hs.ensure((size_t)M * 6 * D * 2));
    TRY(cast_split(tokens, 2 * D, e.p_hs.p, 6 * D, M, 2 * D, false, st));
    TRY(gemm_bf16(e.p_hs.p, 6 * D, pin->bf16, 6 * D, (int)M, D, 6 * D, EPI_BIAS_F32, ep, st));
  } else {
    TRY(cast_rows_bf16(tokens, 2 * D, e.h.p, 2 * D, M, 2 * D, st));
    TRY(gemm_bf16(e.h.p, 2 * D, pin->bf16, 2 * D, (int)M, D, 2 * D, EPI_BIAS_F32, ep, st));
  }
  if (!prefix_in) {
    TRY(layernorm(e.tmp.as<float>(), D, RowMap{}, tnw, tnb, 1e-5f, x, D, RowMap{P, P1, 1}, false, M, D, st));
    TRY(fill_special(atok, x, frames, S, P1, 0, 1, D, st));
  }

  RopeCfg r2; r2.mode = ROPE_2D; r2.tab = e.rope2d_128.as<float2>(); r2.tpf = P1; r2.nsp = 6; r2.gw = gw;
  for (int i = 0; i < e.cfg.head_depth_aa; ++i) {
    if (!(prefix_in && i == 0)) TRY(run_block(e, x, Mh, e.h_frame[i], 1e-5f, 8, 128, frames, P1, r2, nullptr, 0, st));
    if (prefix_out && i == 0) {   // everything up to here needs no context: it can run on the rank that encoded the chunk
      LSVS_CUDA(cudaMemcpyAsync(prefix_out, x, (size_t)Mh * D * 4, cudaMemcpyDeviceToDevice, st));
      return LSVS_OK;
    }
    // temporal cross block on the RAW (B*P1, S, C) view: groups of S consecutive flat rows (:372-377)
    const BlockW& w = e.h_temporal[i];
    if (w.split) {  // fp32-class cross block (csrc/precise.cu): split GEMM operands, fp32 q/k LayerNorm + 1-D RoPE + attention
      TRY(e.p_xs.ensure((size_t)Mh * 3 * D * 2)); TRY(e.p_ys.ensure((size_t)My * 3 * D * 2)); TRY(e.p_qkv.ensure((size_t)Mh * D * 4));
      TRY(e.p_kv.ensure((size_t)My * 2 * D * 4)); TRY(e.p_att.ensure((size_t)Mh * 3 * D * 2));
      float *qf = e.p_qkv.as<float>(), *kvf = e.p_kv.as<float>();
      TRY(layernorm_split(x, D, w.n1w, w.n1b, 1e-5f, e.p_xs.p, 3 * D, Mh, D, st));
      TRY(layernorm_split(first ? x : overlap_in, D, w.n3w, w.n3b, 1e-5f, e.p_ys.p, 3 * D, My, D, st));
      GemmEpilogue eq; eq.bias = w.q_b; eq.out = qf; eq.ldo = D;
      TRY(gemm_bf16(e.p_xs.p, 3 * D, w.q_w, 3 * D, (int)Mh, D, 3 * D, EPI_BIAS_F32, eq, st));
      GemmEpilogue ek; ek.bias = w.kv_b; ek.out = kvf; ek.ldo = 2 * D;
      TRY(gemm_bf16(e.p_ys.p, 3 * D, w.kv_w, 3 * D, (int)My, 2 * D, 3 * D, EPI_BIAS_F32, ek, st));
      TRY(headnorm_rope_f32(qf, D, Mh, 0, 8, 128, w.qn_w, w.qn_b, 1e-5f, ROPE_1D, e.rope1d_128.as<float2>(), 0, 0, 0, d_qi, S, st));
      TRY(headnorm_rope_f32(kvf, 2 * D, My, 0, 8, 128, w.kn_w, w.kn_b, 1e-5f, ROPE_1D, e.rope1d_128.as<float2>(), 0, 0, 0, d_ki, T, st));
      AttentionF32Args aa{qf, kvf, kvf + D, e.p_att.p, D, 2 * D, 2 * D, 3 * D, D, B * P1, 8, 128, S, T, 1.0f / sqrtf(128.f)};
      TRY(attention_f32(aa, st));
      GemmEpilogue er; er.bias = w.proj_b; er.gamma = w.ls1; er.resid = x; er.ldr = D;
      TRY(gemm_bf16(e.p_att.p, 3 * D, w.proj_w, 3 * D, (int)Mh, D, 3 * D, EPI_RESID_F32, er, st));
      TRY(run_mlp_precise(e, x, Mh, w, 1e-5f, D, nullptr, 0, st));
      continue;
    }
    __nv_bfloat16 *yn = e.yn.as<__nv_bfloat16>(), *q = e.qkv.as<__nv_bfloat16>(), *kv = e.kvb.as<__nv_bfloat16>(), *att = e.att.as<__nv_bfloat16>(), *hb = e.h.as<__nv_bfloat16>();
    TRY(layernorm(x, D, RowMap{}, w.n1w, w.n1b, 1e-5f, xn, D, RowMap{}, true, Mh, D, st));
    TRY(layernorm(first ? x : overlap_in, D, RowMap{}, w.n3w, w.n3b, 1e-5f, yn, D, RowMap{}, true, My, D, st));
    GemmEpilogue eq; eq.bias = w.q_b; eq.out = q; eq.ldo = D; eq.qn_w = w.qn_w; eq.qn_b = w.qn_b; eq.n_q_cols = D; eq.n_k_cols = 0;
    eq.rope_mode = ROPE_1D; eq.rope_tab = e.rope1d_128.as<float2>(); eq.pos_ids = d_qi; eq.pos_period = S;
    TRY(gemm_bf16(xn, D, w.q_w, D, (int)Mh, D, D, EPI_HEADNORM128_BF16, eq, st));
    GemmEpilogue ek; ek.bias = w.kv_b; ek.out = kv; ek.ldo = 2 * D; ek.kn_w = w.kn_w; ek.kn_b = w.kn_b; ek.n_q_cols = 0; ek.n_k_cols = D;
    ek.rope_mode = ROPE_1D; ek.rope_tab = e.rope1d_128.as<float2>(); ek.pos_ids = d_ki; ek.pos_period = T;
    TRY(gemm_bf16(yn, D, w.kv_w, D, (int)My, 2 * D, D, EPI_HEADNORM128_BF16, ek, st));
    AttentionArgs aa{q, kv, kv + D, att, D, 2 * D, 2 * D, D, B * P1, 8, 128, S, T, 1.0f / sqrtf(128.f)};
    TRY(attention_fwd(aa, st));
    GemmEpilogue er; er.bias = w.proj_b; er.gamma = w.ls1; er.resid = x; er.ldr = D;
    TRY(gemm_bf16(att, D, w.proj_w, D, (int)Mh, D, D, EPI_RESID_F32, er, st));
    TRY(layernorm(x, D, RowMap{}, w.n2w, w.n2b, 1e-5f, xn, D, RowMap{}, true, Mh, D, st));
    GemmEpilogue e1; e1.bias = w.fc1_b; e1.out = hb; e1.ldo = 4 * D;
    TRY(gemm_bf16(xn, D, w.fc1_w, D, (int)Mh, 4 * D, D, EPI_BIAS_GELU_BF16, e1, st));
    GemmEpilogue e2; e2.bias = w.fc2_b; e2.gamma = w.ls2; e2.resid = x; e2.ldr = D;
    TRY(gemm_bf16(hb, 4 * D, w.fc2_w, 4 * D, (int)Mh, D, 4 * D, EPI_RESID_F32, e2, st));
  }
  // processed overlap tokens for the next chunk: frame 0 and the last `next_overlap` frames (:343)
  for (int b = 0; b < B; ++b) {
    const size_t frame_bytes = (size_t)P1 * D * 4;
    float* dst = overlap_out + (size_t)b * (1 + next_overlap) * P1 * D;
    const float* src = x + (size_t)b * S * P1 * D;
    LSVS_CUDA(cudaMemcpyAsync(dst, src, frame_bytes, cudaMemcpyDeviceToDevice, st));
    if (next_overlap > 0)
      LSVS_CUDA(cudaMemcpyAsync(dst + (size_t)P1 * D, src + (size_t)(S - next_overlap) * P1 * D, frame_bytes * next_overlap, cudaMemcpyDeviceToDevice, st));
  }

  return decode_forward(e, x, (long long)P1 * D, B, S, d_dec, memory_in, chunk_sim3, frame_se3, memory_out, st);
}
}  // namespace
}  // namespace lsvs

// ================================================================================================ CameraHead
extern "C" int lsvs_camera_head_forward(lsvs_engine* h, const float* tokens_last, int B, int S, int P, int num_iterations,
                                        float* pose_enc, float* pose_enc_iters, void* stream) {
  Engine& e = *reinterpret_cast<Engine*>(h);
  cudaStream_t st = (cudaStream_t)stream;
  LSVS_CHECK_ARG(e.finalized && e.cfg.with_camera_head, "camera_head_forward: engine has no camera head / not finalized");
  LSVS_CHECK_ARG(tokens_last && pose_enc && B > 0 && S > 0 && P > 0 && num_iterations > 0, "camera_head_forward: bad arguments");
  ConstWeights const_w(true);   // every W below is an engine-owned weight: the few-row GEMMs prefetch it under their predecessor
  const int C = 2048, frames = B * S;
  TRY(ensure_workspace(e, 2 * (long long)frames + 256, 0));
  const size_t need_scratch = (size_t)frames * (size_t)(C * 8 + 3 * C + 3 * C + 4 * C + 4096) + (1 << 14);
  TRY(e.scratch.ensure(need_scratch * 4));
  e.scratch_off = 0;
  const float *tnw, *tnb, *trw, *trb, *empty, *epw, *epb, *mb, *b1b, *b2w, *b2b;
  TRY(need_f32(e, "camera_head.token_norm.weight", &tnw, C)); TRY(need_f32(e, "camera_head.token_norm.bias", &tnb, C));
  TRY(need_f32(e, "camera_head.trunk_norm.weight", &trw, C)); TRY(need_f32(e, "camera_head.trunk_norm.bias", &trb, C));
  TRY(need_f32(e, "camera_head.empty_pose_tokens", &empty, 9));
  TRY(need_f32(e, "camera_head.embed_pose.weight", &epw, 9LL * C)); TRY(need_f32(e, "camera_head.embed_pose.bias", &epb, C));
  const __nv_bfloat16 *mws, *b1ws;   // split [hi | hi | lo] (camera_vector_gemm)
  bool sp1 = false, sp2 = false;
  TRY(need_bf16(e, "camera_head.poseLN_modulation.1.weight", &mws, 3 * C, C, &sp1)); TRY(need_f32(e, "camera_head.poseLN_modulation.1.bias", &mb, 3 * C));
  TRY(need_bf16(e, "camera_head.pose_branch.fc1.weight", &b1ws, C / 2, C, &sp2)); TRY(need_f32(e, "camera_head.pose_branch.fc1.bias", &b1b, C / 2));
  LSVS_CHECK_ARG(sp1 && sp2, "camera_head_forward: vector-path weights are not stored split");
  TRY(e.p_xs.ensure((size_t)frames * 3 * C * 2));
  void* xs = e.p_xs.p;
  TRY(need_f32(e, "camera_head.pose_branch.fc2.weight", &b2w, 9LL * (C / 2))); TRY(need_f32(e, "camera_head.pose_branch.fc2.bias", &b2b, 9));
  float* tok = e.scratch_f32((size_t)frames * C); float* normed = e.scratch_f32((size_t)frames * C);
  float* emb = e.scratch_f32((size_t)frames * C); float* mod = e.scratch_f32((size_t)frames * 3 * C);
  float* xx = e.scratch_f32((size_t)frames * C); float* xn = e.scratch_f32((size_t)frames * C);
  float* qkv = e.scratch_f32((size_t)frames * 3 * C); float* att = e.scratch_f32((size_t)frames * C);
  float* hb = e.scratch_f32((size_t)frames * 4 * C); float* bh = e.scratch_f32((size_t)frames * (C / 2));
  float* pred = e.scratch_f32((size_t)frames * 9); float* delta = e.scratch_f32((size_t)frames * 9); float* pin = e.scratch_f32((size_t)frames * 9);
  // camera token of the last tapped layer = row 0 of every frame
  TRY(layernorm(tokens_last, (long long)P * C, RowMap{}, tnw, tnb, 1e-5f, tok, C, RowMap{}, false, frames, C, st));
  TRY(layernorm(tok, C, RowMap{}, nullptr, nullptr, 1e-6f, normed, C, RowMap{}, false, frames, C, st));  // adaln_norm (no affine)
  for (int it = 0; it < num_iterations; ++it) {
    if (it == 0) TRY(combine_rows(empty, 0, nullptr, 0, pin, 9, frames, 9, -1, -1, st));  // broadcast the empty pose token
    else TRY(combine_rows(pred, 9, nullptr, 0, pin, 9, frames, 9, -1, -1, st));
    TRY(linear_f32(pin, 9, epw, epb, emb, C, frames, C, 9, ACT_NONE, ACT_NONE, nullptr, false, st));
    {   // mod = Linear(SiLU(emb)): fp32-class on the tensor cores (weight stream 75 MB)
      TRY(cast_split_act(emb, C, xs, 3LL * C, frames, C, 2, st));
      GemmEpilogue em; em.bias = mb; em.out = mod; em.ldo = 3 * C;
      TRY(gemm_bf16(xs, 3 * C, mws, 3 * C, frames, 3 * C, 3 * C, EPI_BIAS_F32, em, st));
    }
    TRY(modulate(normed, tok, mod, xx, frames, C, st));
    // trunk: bf16 tensor-core blocks on the fp32 residual stream (16 heads x 128, sequence = the S frames of a chunk)
    for (int i = 0; i < 4; ++i) TRY(run_block(e, xx, frames, e.cam_trunk[i], 1e-5f, 16, 128, B, S, RopeCfg{}, nullptr, 0, st));
    TRY(layernorm(xx, C, RowMap{}, trw, trb, 1e-5f, xn, C, RowMap{}, false, frames, C, st));
    {   // pose branch: fc1 fp32-class on the tensor cores, its GELU applied on the way into fc2
      TRY(cast_split_act(xn, C, xs, 3LL * C, frames, C, 0, st));
      GemmEpilogue e1; e1.bias = b1b; e1.out = bh; e1.ldo = C / 2;
      TRY(gemm_bf16(xs, 3 * C, b1ws, 3 * C, frames, C / 2, 3 * C, EPI_BIAS_F32, e1, st));
    }
    TRY(linear_f32(bh, C / 2, b2w, b2b, delta, 9, frames, 9, C / 2, ACT_GELU, ACT_NONE, nullptr, false, st));
    if (it == 0) TRY(combine_rows(delta, 9, nullptr, 0, pred, 9, frames, 9, -1, -1, st));
    else TRY(combine_rows(pred, 9, delta, 9, pred, 9, frames, 9, -1, -1, st));
    // UPSTREAM returns the activated encoding of every refinement iteration (the reference only reads the last, :109)
    if (pose_enc_iters) TRY(combine_rows(pred, 9, nullptr, 0, pose_enc_iters + (size_t)it * frames * 9, 9, frames, 9, 7, -1, st));
  }
  TRY(combine_rows(pred, 9, nullptr, 0, pose_enc, 9, frames, 9, 7, -1, st));  // activate_pose: T, quat linear; FoV relu
  if (e.scratch_overflow) { e.scratch_overflow = false; return fail(LSVS_ECUDA, "camera_head_forward: scratch under-sized"); }
  return LSVS_OK;
}

// ---------------------------------------------------------------------------------------------------- DPT head
// UPSTREAM vggt/heads/dpt_head.py DPTHead.forward (depth_head / point_head, featureAligned_vggt.py:166,183).
// Per frame chunk: LayerNorm of the 4 tapped token maps -> 1x1 projections (+uv position embedding) -> resize
// (transposed conv x4 / x2, identity, 3x3 stride-2 conv) -> layer_rn 3x3 convs -> 4 refinement stages (residual conv units,
// bilinear x2, 1x1 out_conv) -> output_conv1 -> bilinear to the image size (+ position embedding) -> output_conv2 ->
// activation.  The 1x1 out_conv of a fusion block is applied BEFORE its bilinear upsample (both are linear and the
// interpolation weights sum to one, so the result is the same and the GEMM runs on a quarter of the pixels).
extern "C" int lsvs_dpt_head_forward(lsvs_engine* h, const char* prefix, const float* const* taps, int frames, int P, int H, int W,
                                     int output_dim, int activation, float* pred, float* conf, int frames_chunk, void* stream) {
  Engine& e = *reinterpret_cast<Engine*>(h);
  cudaStream_t st = (cudaStream_t)stream;
  LSVS_CHECK_ARG(prefix && taps && pred && conf && frames > 0, "dpt_head_forward: bad arguments");
  FewRowsKernel no_fewrows(false);   // per-frame results must not depend on the frames per pass (coarse levels cross M = 128)
  LSVS_CHECK_ARG(output_dim >= 2 && output_dim <= 4 && (activation == 0 || activation == 1), "dpt_head_forward: output_dim 2..4, activation 0 (exp) or 1 (inv_log)");
  const int ph = H / 14, pw = W / 14, Pp = ph * pw, C = 2048;
  LSVS_CHECK_ARG(ph > 0 && pw > 0 && P == Pp + 5, "dpt_head_forward: token count %d does not match the %dx%d patch grid (+5)", P, ph, pw);
  for (int l = 0; l < 4; ++l) LSVS_CHECK_ARG(taps[l], "dpt_head_forward: tap %d is null", l);
  const std::string pre(prefix);
  const int oc[4] = {256, 512, 1024, 1024};
  struct Grid { int h, w; long long rows(int F) const { return (long long)F * (h + 2) * (w + 2); } };
  const Grid g1{4 * ph, 4 * pw}, g2{2 * ph, 2 * pw}, g3{ph, pw}, g4{(ph - 1) / 2 + 1, (pw - 1) / 2 + 1}, g5{8 * ph, 8 * pw}, gF{14 * ph, 14 * pw};
  const Grid lvl[4] = {g1, g2, g3, g4};
  // frames per pass: the caller's choice, else as many as keep the workspace under ~6 GB (larger passes fill the GPU better
  // on the coarse levels: 32 frames at 154x518 run 1.35x faster than 4 x 8)
  int F0 = frames_chunk > 0 ? frames_chunk : frames;
  if (frames_chunk <= 0) {
    const double per_frame = 2.0 * ((double)(gF.h + 2) * (gF.w + 2) * 192 + (double)(g5.h + 2) * (g5.w + 2) * 384 +
                                    (double)(g1.h + 2) * (g1.w + 2) * 2304 + (double)Pp * 16384);
    const int fit = (int)(6.0e9 / per_frame);
    F0 = fit < 1 ? 1 : fit;
  }
  if (F0 > frames) F0 = frames;
  const int F0_plain = F0;
  // ---- weights
  const float *nw, *nb;
  TRY(need_f32(e, pre + "norm.weight", &nw, C)); TRY(need_f32(e, pre + "norm.bias", &nb, C));
  const __nv_bfloat16 *pw_[4], *rnw[4], *ct0w, *ct1w, *c3w, *oc1w, *oc2w;
  const float *pb_[4], *ct0b, *ct1b, *c3b, *oc1b, *oc2b, *finw, *finb;
  // fp32-class mode (precision >= 1): every weight below is stored split [hi | hi | lo]; the flags must agree
  int n_w = 0, n_split = 0;
  auto need_w = [&](const std::string& name, const __nv_bfloat16** out, int rows, int cols) -> int {
    bool sp = false;
    TRY(need_bf16(e, name, out, rows, cols, &sp));
    ++n_w; n_split += sp ? 1 : 0;
    return LSVS_OK;
  };
  for (int l = 0; l < 4; ++l) {
    TRY(need_w(pre + "projects." + std::to_string(l) + ".weight", &pw_[l], oc[l], C));
    TRY(need_f32(e, pre + "projects." + std::to_string(l) + ".bias", &pb_[l], oc[l]));
    TRY(need_w(pre + "scratch.layer" + std::to_string(l + 1) + "_rn.weight", &rnw[l], 256, 9 * oc[l]));
  }
  TRY(need_w(pre + "resize_layers.0.weight", &ct0w, 16 * 256, 256)); TRY(need_f32(e, pre + "resize_layers.0.bias", &ct0b, 256));
  TRY(need_w(pre + "resize_layers.1.weight", &ct1w, 4 * 512, 512)); TRY(need_f32(e, pre + "resize_layers.1.bias", &ct1b, 512));
  TRY(need_w(pre + "resize_layers.3.weight", &c3w, 1024, 9 * 1024)); TRY(need_f32(e, pre + "resize_layers.3.bias", &c3b, 1024));
  TRY(need_w(pre + "scratch.output_conv1.weight", &oc1w, 128, 9 * 256)); TRY(need_f32(e, pre + "scratch.output_conv1.bias", &oc1b, 128));
  TRY(need_w(pre + "scratch.output_conv2.0.weight", &oc2w, 64, 9 * 128)); TRY(need_f32(e, pre + "scratch.output_conv2.0.bias", &oc2b, 64));
  TRY(need_f32(e, pre + "scratch.output_conv2.2.weight", &finw, 32LL * output_dim)); TRY(need_f32(e, pre + "scratch.output_conv2.2.bias", &finb, output_dim));
  struct Rcu { const __nv_bfloat16 *w1, *w2; const float *b1, *b2; };
  struct Fuse { Rcu u1, u2; const __nv_bfloat16* ow; const float* ob; };
  Fuse fu[4];
  for (int r = 0; r < 4; ++r) {
    const std::string b = pre + "scratch.refinenet" + std::to_string(r + 1) + ".";
    for (int u = (r == 3 ? 2 : 1); u <= 2; ++u) {
      Rcu& q = u == 1 ? fu[r].u1 : fu[r].u2;
      const std::string c = b + "resConfUnit" + std::to_string(u) + ".";
      TRY(need_w(c + "conv1.weight", &q.w1, 256, 9 * 256)); TRY(need_f32(e, c + "conv1.bias", &q.b1, 256));
      TRY(need_w(c + "conv2.weight", &q.w2, 256, 9 * 256)); TRY(need_f32(e, c + "conv2.bias", &q.b2, 256));
    }
    TRY(need_w(b + "out_conv.weight", &fu[r].ow, 256, 256)); TRY(need_f32(e, b + "out_conv.bias", &fu[r].ob, 256));
  }
  LSVS_CHECK_ARG(n_split == 0 || n_split == n_w, "dpt_head_forward: '%s' mixes split and plain weights", prefix);
  const bool precise = n_split > 0;   // fp32 activations, split-bf16 GEMM operands (3x the reduction length)
  // ---- workspace (bf16 elements; fp32 in the fp32-class mode), bump-allocated
  if (precise && frames_chunk <= 0) {   // the split patch rows of the full-resolution convolution are 27 * 128 * 2 bytes per pixel
    const double per_frame = (double)(gF.h + 2) * (gF.w + 2) * (27.0 * 128 * 2 + 4.0 * 192) + (double)(g5.h + 2) * (g5.w + 2) * 4.0 * 384 +
                             (double)(g1.h + 2) * (g1.w + 2) * 4.0 * 2304 + (double)Pp * 4.0 * 16384;
    const int fit = (int)(6.0e9 / per_frame);
    F0 = fit < 1 ? 1 : (fit < F0_plain ? fit : F0_plain);
  }
  const size_t esz = precise ? 4 : 2;
  const long long rows0 = (long long)F0 * Pp, rows4 = (long long)F0 * g4.h * g4.w;
  size_t off = 0;
  auto take = [&](long long elems) { const size_t o = off; off += ((size_t)elems * esz + 255) & ~size_t(255); return o; };
  const size_t o_tok = take(rows0 * C), o_proj = take(rows0 * 1024), o_ct = take(rows0 * 4096 > rows4 * 1024 ? rows0 * 4096 : rows4 * 1024),
               o_col = take(rows4 * 9 * 1024);
  size_t o_x[4], o_rn[4];
  for (int l = 0; l < 4; ++l) { o_x[l] = take(lvl[l].rows(F0) * oc[l]); o_rn[l] = take(lvl[l].rows(F0) * 256); }
  const long long rmax = g1.rows(F0);
  const size_t o_t = take(rmax * 256), o_a2 = take(rmax * 256), o_o = take(rmax * 256), o_oc = take(rmax * 256);
  const size_t o_u3 = take(g3.rows(F0) * 256), o_u2 = take(g2.rows(F0) * 256), o_u1 = take(g1.rows(F0) * 256), o_u5 = take(g5.rows(F0) * 256);
  const size_t o_o1 = take(g5.rows(F0) * 128), o_uf = take(gF.rows(F0) * 128), o_o2 = take(gF.rows(F0) * 64);
  // uv position-embedding tables (fp32): per projection width on the patch grid, and 128 channels at full resolution
  size_t o_tab[5];
  for (int l = 0; l < 4; ++l) o_tab[l] = take((long long)(ph + pw) * oc[l]);          // (w + h) * C/2 floats = (w + h) * C bf16-sized slots
  o_tab[4] = take((long long)(gF.h + gF.w) * 128);
  size_t o_split = 0;   // fp32-class mode: split-bf16 GEMM operand of the largest convolution / projection
  if (precise) {
    long long mx = rows0 * 3 * C;
    auto upd = [&](long long v) { if (v > mx) mx = v; };
    for (int l = 0; l < 4; ++l) upd(lvl[l].rows(F0) * 27 * oc[l]);
    upd(g1.rows(F0) * 27 * 256); upd(g5.rows(F0) * 27 * 256); upd(gF.rows(F0) * 27 * 128); upd(rows4 * 27 * 1024); upd(rows0 * 3 * 1024);
    o_split = off; off += ((size_t)mx * 2 + 255) & ~size_t(255);
  }
  LSVS_CHECK_ARG(gF.rows(F0) < (1ll << 31), "dpt_head_forward: frame chunk too large");
  TRY(e.dpt_ws.ensure(off));
  uint8_t* ws = e.dpt_ws.as<uint8_t>();
  auto B16 = [&](size_t o) { return reinterpret_cast<__nv_bfloat16*>(ws + o); };
  const float aspect = (float)W / (float)H;
  auto F32 = [&](size_t o) { return reinterpret_cast<float*>(ws + o); };
  float *tabU[5], *tabV[5];
  for (int l = 0; l < 5; ++l) {
    const int tw = l < 4 ? pw : gF.w, th = l < 4 ? ph : gF.h, tc = l < 4 ? oc[l] : 128;
    tabU[l] = F32(o_tab[l]); tabV[l] = tabU[l] + (size_t)tw * (tc / 2);
    TRY(dpt_uv_tables(tabU[l], tabV[l], th, tw, tc, aspect, 0.1f, st));
  }

  auto conv = [&](const __nv_bfloat16* x, const Grid& g, int F, int Cin, const __nv_bfloat16* w, const float* b, int OC, int taps_,
                  bool relu, const __nv_bfloat16* r1, const __nv_bfloat16* r2, __nv_bfloat16* out) -> int {
    GemmEpilogue ep;
    ep.bias = b; ep.out = out; ep.ldo = OC;
    ep.conv_taps = taps_ == 9 ? 9 : 0; ep.conv_c = Cin; ep.conv_hp = g.h + 2; ep.conv_wp = g.w + 2; ep.conv_relu = relu ? 1 : 0; ep.conv_mask = 1;
    ep.res1 = r1; ep.res2 = r2;
    return gemm_bf16(x, Cin, w, taps_ * Cin, (int)g.rows(F), OC, taps_ * Cin, EPI_CONV_BF16, ep, st);
  };
  // ResidualConvUnit on a = relu(x) (the fusion blocks use an in-place ReLU, so the skip path carries relu(x)):
  // out = [relu](conv2(relu(conv1(a))) + a + extra)
  auto rcu = [&](const Rcu& q, const __nv_bfloat16* a, const Grid& g, int F, const __nv_bfloat16* extra, bool relu_out, __nv_bfloat16* out) -> int {
    TRY(conv(a, g, F, 256, q.w1, q.b1, 256, 9, true, nullptr, nullptr, B16(o_t)));
    return conv(B16(o_t), g, F, 256, q.w2, q.b2, 256, 9, relu_out, a, extra, out);
  };

  if (precise) {
    // ---- fp32-class flow: the same graph with fp32 activations (the reference runs these heads with autocast disabled,
    // featureAligned_vggt.py:103).  A convolution = split patch rows (dpt_split_im2col3 / cast_split) x split weights on the
    // tcgen05 GEMM (fp32 out) + its tail (residuals, ReLU, border mask) as one element-wise pass.
    void* S = ws + o_split;
    auto gemm_f32 = [&](long long rows, int K3, const __nv_bfloat16* w, const float* b, int OC, float* out) -> int {
      GemmEpilogue ep;
      ep.bias = b; ep.out = out; ep.ldo = OC;
      return gemm_bf16(S, K3, w, K3, (int)rows, OC, K3, EPI_BIAS_F32, ep, st);
    };
    auto convp = [&](const float* x, const Grid& g, int F, int Cin, const __nv_bfloat16* w, const float* b, int OC, int taps_, bool relu,
                     const float* r1, const float* r2, float* out) -> int {
      const long long rows = g.rows(F);
      if (taps_ == 9) TRY(dpt_split_im2col3(x, S, F, g.h + 2, g.w + 2, Cin, st));
      else TRY(cast_split(x, Cin, S, 3LL * Cin, rows, Cin, false, st));
      TRY(gemm_f32(rows, 3 * taps_ * Cin, w, b, OC, out));
      return dpt_post_f32(out, r1, r2, relu ? 1 : 0, F, g.h + 2, g.w + 2, OC, st);
    };
    auto rcup = [&](const Rcu& q, const float* a, const Grid& g, int F, const float* extra, bool relu_out, float* out) -> int {
      TRY(convp(a, g, F, 256, q.w1, q.b1, 256, 9, true, nullptr, nullptr, F32(o_t)));
      return convp(F32(o_t), g, F, 256, q.w2, q.b2, 256, 9, relu_out, a, extra, out);
    };
    for (int f0 = 0; f0 < frames; f0 += F0) {
      const int F = frames - f0 < F0 ? frames - f0 : F0;
      const long long r0 = (long long)F * Pp;
      for (int l = 0; l < 4; ++l) {
        for (int f = 0; f < F; ++f)   // patch tokens of one frame (the 5 special tokens are skipped)
          TRY(layernorm_split(taps[l] + ((size_t)(f0 + f) * P + 5) * C, C, nw, nb, 1e-5f, reinterpret_cast<__nv_bfloat16*>(S) + (size_t)f * Pp * 3 * C, 3LL * C, Pp, C, st));
        TRY(gemm_f32(r0, 3 * C, pw_[l], pb_[l], oc[l], F32(o_proj)));
        TRY(dpt_add_pos_embed(F32(o_proj), F, ph, pw, oc[l], aspect, 0.1f, tabU[l], tabV[l], st, true));
        if (l == 0 || l == 1) {
          const int k = l == 0 ? 4 : 2;
          TRY(cast_split(F32(o_proj), oc[l], S, 3LL * oc[l], r0, oc[l], false, st));
          TRY(gemm_f32(r0, 3 * oc[l], l == 0 ? ct0w : ct1w, nullptr, k * k * oc[l], F32(o_ct)));
          TRY(dpt_convt_shuffle(F32(o_ct), l == 0 ? ct0b : ct1b, F32(o_x[l]), F, ph, pw, oc[l], k, st, true));
        } else if (l == 2) {
          TRY(dpt_pad(F32(o_proj), F32(o_x[l]), F, ph, pw, oc[l], st, true));
        } else {
          TRY(dpt_im2col_s2(F32(o_proj), F32(o_col), F, ph, pw, oc[l], st, true));
          const long long r4 = (long long)F * g4.h * g4.w;
          TRY(cast_split(F32(o_col), 9 * 1024, S, 27LL * 1024, r4, 9 * 1024, false, st));
          TRY(gemm_f32(r4, 27 * 1024, c3w, c3b, 1024, F32(o_ct)));
          TRY(dpt_pad(F32(o_ct), F32(o_x[l]), F, g4.h, g4.w, 1024, st, true));
        }
        TRY(convp(F32(o_x[l]), lvl[l], F, oc[l], rnw[l], nullptr, 256, 9, true, nullptr, nullptr, F32(o_rn[l])));
      }
      TRY(rcup(fu[3].u2, F32(o_rn[3]), g4, F, nullptr, false, F32(o_o)));
      TRY(convp(F32(o_o), g4, F, 256, fu[3].ow, fu[3].ob, 256, 1, false, nullptr, nullptr, F32(o_oc)));
      TRY(dpt_bilinear(F32(o_oc), F32(o_u3), F, g4.h, g4.w, g3.h, g3.w, 256, aspect, 0.f, nullptr, nullptr, st, true));
      const size_t o_upp[4] = {o_u1, o_u2, o_u3, 0};
      for (int r = 2; r >= 0; --r) {
        const Grid& g = lvl[r];
        TRY(rcup(fu[r].u1, F32(o_rn[r]), g, F, F32(o_upp[r]), true, F32(o_a2)));
        TRY(rcup(fu[r].u2, F32(o_a2), g, F, nullptr, false, F32(o_o)));
        TRY(convp(F32(o_o), g, F, 256, fu[r].ow, fu[r].ob, 256, 1, false, nullptr, nullptr, F32(o_oc)));
        const Grid& gn = r == 0 ? g5 : lvl[r - 1];
        TRY(dpt_bilinear(F32(o_oc), F32(r == 0 ? o_u5 : o_upp[r - 1]), F, g.h, g.w, gn.h, gn.w, 256, aspect, 0.f, nullptr, nullptr, st, true));
      }
      TRY(convp(F32(o_u5), g5, F, 256, oc1w, oc1b, 128, 9, false, nullptr, nullptr, F32(o_o1)));
      TRY(dpt_bilinear(F32(o_o1), F32(o_uf), F, g5.h, g5.w, gF.h, gF.w, 128, aspect, 0.1f, tabU[4], tabV[4], st, true));
      TRY(convp(F32(o_uf), gF, F, 128, oc2w, oc2b, 64, 9, true, nullptr, nullptr, F32(o_o2)));
      TRY(dpt_final(F32(o_o2), 64, finw, finb, output_dim, activation, pred + (size_t)f0 * gF.h * gF.w * (output_dim - 1),
                    conf + (size_t)f0 * gF.h * gF.w, F, gF.h, gF.w, st, true));
    }
    return LSVS_OK;
  }

  for (int f0 = 0; f0 < frames; f0 += F0) {
    const int F = frames - f0 < F0 ? frames - f0 : F0;
    const long long r0 = (long long)F * Pp;
    for (int l = 0; l < 4; ++l) {
      const float* tap = taps[l] + (size_t)f0 * P * C;
      TRY(layernorm(tap, C, RowMap{Pp, P, 5}, nw, nb, 1e-5f, B16(o_tok), C, RowMap{}, true, r0, C, st));
      GemmEpilogue ep;
      ep.bias = pb_[l]; ep.out = B16(o_proj); ep.ldo = oc[l];
      TRY(gemm_bf16(B16(o_tok), C, pw_[l], C, (int)r0, oc[l], C, EPI_BIAS_BF16, ep, st));
      TRY(dpt_add_pos_embed(B16(o_proj), F, ph, pw, oc[l], aspect, 0.1f, tabU[l], tabV[l], st));
      if (l == 0 || l == 1) {
        const int k = l == 0 ? 4 : 2;
        GemmEpilogue ec;
        ec.out = B16(o_ct); ec.ldo = k * k * oc[l];
        TRY(gemm_bf16(B16(o_proj), oc[l], l == 0 ? ct0w : ct1w, oc[l], (int)r0, k * k * oc[l], oc[l], EPI_BIAS_BF16, ec, st));
        TRY(dpt_convt_shuffle(B16(o_ct), l == 0 ? ct0b : ct1b, B16(o_x[l]), F, ph, pw, oc[l], k, st));
      } else if (l == 2) {
        TRY(dpt_pad(B16(o_proj), B16(o_x[l]), F, ph, pw, oc[l], st));
      } else {
        TRY(dpt_im2col_s2(B16(o_proj), B16(o_col), F, ph, pw, oc[l], st));
        GemmEpilogue ec;
        ec.bias = c3b; ec.out = B16(o_ct); ec.ldo = 1024;
        TRY(gemm_bf16(B16(o_col), 9 * 1024, c3w, 9 * 1024, F * g4.h * g4.w, 1024, 9 * 1024, EPI_BIAS_BF16, ec, st));
        TRY(dpt_pad(B16(o_ct), B16(o_x[l]), F, g4.h, g4.w, 1024, st));
      }
      // layer_rn: 3x3, no bias; its only consumer is a ResidualConvUnit whose in-place ReLU rewrites it, so store relu(.)
      TRY(conv(B16(o_x[l]), lvl[l], F, oc[l], rnw[l], nullptr, 256, 9, true, nullptr, nullptr, B16(o_rn[l])));
    }
    // refinenet4 (no residual input): out_conv(upsample(RCU2(rn4)))
    TRY(rcu(fu[3].u2, B16(o_rn[3]), g4, F, nullptr, false, B16(o_o)));
    TRY(conv(B16(o_o), g4, F, 256, fu[3].ow, fu[3].ob, 256, 1, false, nullptr, nullptr, B16(o_oc)));
    TRY(dpt_bilinear(B16(o_oc), B16(o_u3), F, g4.h, g4.w, g3.h, g3.w, 256, aspect, 0.f, nullptr, nullptr, st));
    const size_t o_up[4] = {o_u1, o_u2, o_u3, 0};
    for (int r = 2; r >= 0; --r) {
      const Grid& g = lvl[r];
      // output = x0 + RCU1(rn_r), then RCU2's in-place ReLU: a2 = relu(conv2(relu(conv1(rn))) + rn + x0)
      TRY(rcu(fu[r].u1, B16(o_rn[r]), g, F, B16(o_up[r]), true, B16(o_a2)));
      TRY(rcu(fu[r].u2, B16(o_a2), g, F, nullptr, false, B16(o_o)));
      TRY(conv(B16(o_o), g, F, 256, fu[r].ow, fu[r].ob, 256, 1, false, nullptr, nullptr, B16(o_oc)));
      const Grid& gn = r == 0 ? g5 : lvl[r - 1];
      TRY(dpt_bilinear(B16(o_oc), B16(r == 0 ? o_u5 : o_up[r - 1]), F, g.h, g.w, gn.h, gn.w, 256, aspect, 0.f, nullptr, nullptr, st));
    }
    TRY(conv(B16(o_u5), g5, F, 256, oc1w, oc1b, 128, 9, false, nullptr, nullptr, B16(o_o1)));
    TRY(dpt_bilinear(B16(o_o1), B16(o_uf), F, g5.h, g5.w, gF.h, gF.w, 128, aspect, 0.1f, tabU[4], tabV[4], st));
    TRY(conv(B16(o_uf), gF, F, 128, oc2w, oc2b, 64, 9, true, nullptr, nullptr, B16(o_o2)));
    TRY(dpt_final(B16(o_o2), 64, finw, finb, output_dim, activation, pred + (size_t)f0 * gF.h * gF.w * (output_dim - 1),
                  conf + (size_t)f0 * gF.h * gF.w, F, gF.h, gF.w, st));
  }
  return LSVS_OK;
}

extern "C" int lsvs_pose_chain(const float* chunk_sim3, const float* frame_se3, const float* cam_enc, const float* prev_pose_enc,
                               int S_prev, int overlap, int B, int S, int H, int W, float* pose_enc_out, float* point_T,
                               float* scale_out, void* stream) {
  return lsvs::pose_chain(chunk_sim3, frame_se3, cam_enc, prev_pose_enc, S_prev, overlap, B, S, H, W, pose_enc_out, point_T, scale_out,
                          (cudaStream_t)stream);
}

extern "C" int lsvs_pose_chain_gt(const float* chunk_sim3, const float* frame_se3, const float* cam_enc, const float* prev_pose_enc,
                                  int S_prev, int overlap, int B, int S, int H, int W, const float* gt_poses, int gt_rows, int gt_mode,
                                  float* pose_enc_out, float* point_T, float* scale_out, void* stream) {
  LSVS_CHECK_ARG(gt_poses != nullptr, "pose_chain_gt: gt_poses is null");
  return lsvs::pose_chain(chunk_sim3, frame_se3, cam_enc, prev_pose_enc, S_prev, overlap, B, S, H, W, pose_enc_out, point_T, scale_out,
                          (cudaStream_t)stream, gt_poses, gt_rows, gt_mode);
}

extern "C" int lsvs_alignment_decode_forward(lsvs_engine* h, const float* align_tokens, int B, int S, const float* memory_in,
                                             float* chunk_sim3, float* frame_se3, float* memory_out, void* stream) {
  Engine& e = *reinterpret_cast<Engine*>(h);
  cudaStream_t st = (cudaStream_t)stream;
  LSVS_CHECK_ARG(e.finalized && e.cfg.with_alignment_head, "alignment_decode_forward: engine has no alignment head / not finalized");
  LSVS_CHECK_ARG(align_tokens && chunk_sim3 && B > 0 && S > 0 && (S == 1 || frame_se3), "alignment_decode_forward: bad arguments");
  const int NM = e.cfg.num_memory_tokens;
  LSVS_CHECK_ARG(NM == 8 || NM == 0, "alignment_decode_forward: num_memory_tokens must be 8 or 0");
  LSVS_CHECK_ARG(memory_out || NM == 0, "alignment_decode_forward: memory_out missing");
  std::vector<int> dec(S + NM);
  for (int j = 0; j < S; ++j) dec[j] = j;
  for (int j = 0; j < NM; ++j) dec[S + j] = 2 * S + j;
  if (e.ids_k_key != (long long)S * 64 + NM) {
    e.ids_k_key = -1;
    TRY(upload_ids(e.ids_k, dec, st));
    e.ids_k_key = (long long)S * 64 + NM;
  }
  return decode_forward(e, align_tokens, 1024, B, S, e.ids_k.as<int>(), memory_in, chunk_sim3, frame_se3, memory_out, st);
}

extern "C" int lsvs_pose_enc_apply_sim3(const float* pose_enc, const float* T, const float* s, float* out, int B, int S, int H, int W,
                                        void* stream) {
  return lsvs::pose_enc_apply_sim3(pose_enc, T, s, out, B, S, H, W, (cudaStream_t)stream);
}
