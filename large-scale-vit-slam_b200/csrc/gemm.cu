// bf16 GEMM  C[M,N] = A[M,K] . W[N,K]^T  (both operands K-major, i.e. nn.Linear layout) on the 5th-gen tensor cores.
// Two persistent, warp-specialised kernels (TMA producer warp / single-thread tcgen05.mma issuer / epilogue warps), fp32
// accumulators double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1:
//   * gemm_bf16_tcgen05_2cta  CTA pairs (cta_group::2): 256x256 tiles, each CTA loads its 128 A rows and half of the W
//                             tile; 5-6 stage 32 KB ring; used when M > 256 and N % 256 == 0 (all large GEMMs)
//   * gemm_bf16_tcgen05<BN>   one CTA, 128 x BN tiles (BN 64 / 128 / 256), 4-stage ring; small M or N
// Fused epilogues (one thread owns one output row of the tile):
//   bias | bias+GELU(erf) | bias -> fp32 | bias+LayerScale+fp32 residual (TMA reduce-add) | bias + per-head LayerNorm + RoPE
//   (q,k) | convolution on a padded NHWC grid (bias + residuals + ReLU + border mask; the producer shifts the A rows per tap).
// bf16 outputs of the pair kernel are staged in 128B-swizzled shared memory and leave as bulk TMA stores.
// Tiles are walked n-fastest so that an A row-panel stays in L2 across its N tiles.  Programmatic dependent launch: the
// prologue (barriers, tensor-memory allocation) overlaps the previous kernel's tail.
//
// Replaces the library calls at: UPSTREAM Attention.qkv/proj, Mlp.fc1/fc2,
// alignment_head.py:242 (project_in), cross_attention.py:55-57,76 (q/k/v/proj), UPSTREAM DPTHead convolutions, plus the
// elementwise q_norm/k_norm/RoPE/LayerScale/residual passes that the reference runs as separate ATen kernels.
#include <cstdlib>

#include "gemm.h"
#include "host_common.h"
#include "ptx.cuh"
#include "tensormap.h"

namespace lsvs {
// Kernel-selection switch for A/B measurements: 0 auto, 1 force the single-CTA kernel, 2 pair kernel without TMA store / reduce-add,
// 3 pair kernel without B loads (WRONG results; halves L2->SM bytes: qkv 75.7 -> 70.6 us, i.e. the pair kernel is not L2-bound).
// It only exists in measurement builds (LSVS_NVCC_DEFINES="-DLSVS_MEASURE", entry point lsvs_debug_gemm_mode in capi_gemm.cu);
// the product library compiles the switch away.
#ifdef LSVS_MEASURE
int g_gemm_mode = 0;
#else
constexpr int g_gemm_mode = 0;
#endif
// Split-K over the TMA reduce-add epilogue (short chunks only: fewer residual-GEMM tiles than CTA pairs) adds the K slices'
// partial sums into the fp32 residual with L2 atomics in arrival order, so those outputs can differ in the last bit from run to
// run.  LSVS_DETERMINISTIC=1 in the environment keeps one slice per tile (bit-reproducible, slower on 4-8 frame chunks).
static bool deterministic() {
  static const bool on = [] { const char* v = getenv("LSVS_DETERMINISTIC"); return v && atoi(v) != 0; }();
  return on;
}
static thread_local bool g_const_weights = false;
ConstWeights::ConstWeights(bool on) : prev(g_const_weights) { g_const_weights = on; }
ConstWeights::~ConstWeights() { g_const_weights = prev; }
static thread_local bool g_fewrows_kernel = true;
FewRowsKernel::FewRowsKernel(bool on) : prev(g_fewrows_kernel) { g_fewrows_kernel = on; }
FewRowsKernel::~FewRowsKernel() { g_fewrows_kernel = prev; }
namespace {

constexpr int BM = 128;
constexpr int BK = 64;           // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int NUM_THREADS = 192;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int PARAM_OFFSET = BAR_OFFSET + 256;
  static constexpr int TOTAL = PARAM_OFFSET + 2048 + 1024;  // barriers + epilogue params + slack for 1024-byte alignment
};

// Exact-erf GELU, branch-free and almost MUFU-free: erf as a clamped rational polynomial x P(x^2) / Q(x^2) (the
// single-precision form used by Eigen; |error| < 5e-7 measured against scipy, i.e. four orders of magnitude below the
// bf16 rounding of the output): 12 FMA + 1 MUFU.RCP.  libdevice erff costs ~2x the instructions, and a formulation with
// two MUFU ops per element (rcp + ex2) makes the fc1 epilogue MUFU-bound at exactly the main-loop rate (measured: 176 us
// vs 115 us).
__device__ __forceinline__ float gelu_erf(float x) {
  const float xc = fminf(fmaxf(x * 0.70710678118654752f, -4.0f), 4.0f);
  const float x2 = xc * xc;
  float p = fmaf(x2, -2.72614225801306e-10f, 2.77068142495902e-08f);
  p = fmaf(x2, p, -2.10102402082508e-06f);
  p = fmaf(x2, p, -5.69250639462346e-05f);
  p = fmaf(x2, p, -7.34990630326855e-04f);
  p = fmaf(x2, p, -2.95459980854025e-03f);
  p = fmaf(x2, p, -1.60960333262415e-02f);
  float q = fmaf(x2, -1.45660718464996e-05f, -2.13374055278905e-04f);
  q = fmaf(x2, q, -1.68282697438203e-03f);
  q = fmaf(x2, q, -7.37332916720468e-03f);
  q = fmaf(x2, q, -1.42647390514189e-02f);
  float rq;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rq) : "f"(q));
  const float hx = 0.5f * x;
  return fmaf(hx, xc * p * rq, hx);
}

// two elements at a time on the packed fp32x2 pipe (FFMA2 / FMUL2): halves the issue slots of the polynomial — the fc1
// epilogue is issue-bound (8 or 16 epilogue warps make no difference: 128 x 256 elements x ~21 instructions per tile and SM)
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_dup(float c) { return f2_pack(c, c); }
__device__ __forceinline__ void gelu_erf2(float& a, float& b) {
  const float ac = fminf(fmaxf(a * 0.70710678118654752f, -4.0f), 4.0f), bc = fminf(fmaxf(b * 0.70710678118654752f, -4.0f), 4.0f);
  const unsigned long long xc = f2_pack(ac, bc), x2 = f2_mul(xc, xc);
  unsigned long long p = f2_fma(x2, f2_dup(-2.72614225801306e-10f), f2_dup(2.77068142495902e-08f));
  p = f2_fma(x2, p, f2_dup(-2.10102402082508e-06f));
  p = f2_fma(x2, p, f2_dup(-5.69250639462346e-05f));
  p = f2_fma(x2, p, f2_dup(-7.34990630326855e-04f));
  p = f2_fma(x2, p, f2_dup(-2.95459980854025e-03f));
  p = f2_fma(x2, p, f2_dup(-1.60960333262415e-02f));
  unsigned long long q = f2_fma(x2, f2_dup(-1.45660718464996e-05f), f2_dup(-2.13374055278905e-04f));
  q = f2_fma(x2, q, f2_dup(-1.68282697438203e-03f));
  q = f2_fma(x2, q, f2_dup(-7.37332916720468e-03f));
  q = f2_fma(x2, q, f2_dup(-1.42647390514189e-02f));
  float q0, q1, r0, r1;
  f2_unpack(q, q0, q1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(q0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(q1));
  const unsigned long long hx = f2_mul(f2_pack(a, b), f2_dup(0.5f));
  const unsigned long long t = f2_mul(f2_mul(xc, p), f2_pack(r0, r1));
  f2_unpack(f2_fma(hx, t, hx), a, b);
}

// ---- epilogue helpers: `v` holds NC consecutive fp32 accumulator columns of one output row ----------
template <int NC>
__device__ __forceinline__ void store_bf16_row(__nv_bfloat16* dst, const float* v) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < NC / 8; ++i) {
    uint4 u;
    u.x = ptx::pack_bf16(v[8 * i + 0], v[8 * i + 1]);
    u.y = ptx::pack_bf16(v[8 * i + 2], v[8 * i + 3]);
    u.z = ptx::pack_bf16(v[8 * i + 4], v[8 * i + 5]);
    u.w = ptx::pack_bf16(v[8 * i + 6], v[8 * i + 7]);
    d4[i] = u;
  }
}

template <int NC>
__device__ __forceinline__ void load_acc(uint32_t taddr, float* v) {
#pragma unroll
  for (int c = 0; c < NC; c += 32) ptx::tmem_ld_32x32b_x32(taddr + c, reinterpret_cast<uint32_t*>(v + c));
  ptx::tmem_ld_wait();
}

// bf16 output through shared memory + TMA store (pair kernel): a lane holding one output row writes 16-byte pieces of 32
// different rows per store instruction (32 sectors each); staged in a per-warp 32 x 64 tile (128B-swizzled, conflict-free)
// the same data leaves the SM as one bulk tensor store (measured: qkv GEMM 75.6 us, 61.9 us with the per-lane stores removed).
struct TmaOut { const CUtensorMap* tm; uint8_t* tile; int m_base; };
template <int NC>
__device__ __forceinline__ void store_bf16_tma(const TmaOut& to, const float* v, int n) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < NC / 64; ++t) {
    if (lane == 0) ptx::tma_store_wait_read<0>();  // the previous store has finished reading the tile
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* f = v + t * 64 + 8 * i;
      *reinterpret_cast<uint4*>(to.tile + lane * 128 + ((i ^ (lane & 7)) << 4)) =
          make_uint4(ptx::pack_bf16(f[0], f[1]), ptx::pack_bf16(f[2], f[3]), ptx::pack_bf16(f[4], f[5]), ptx::pack_bf16(f[6], f[7]));
    }
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_2d(to.tm, to.tile, n + t * 64, to.m_base);
      ptx::tma_store_commit();
    }
  }
}

// Token position of global row m for the 2-D RoPE: frames of `tpf` tokens, the first `nsp` of each are
// special (position 0,0), patches are row-major on a grid `gw` wide and get (y+1, x+1).
__device__ __forceinline__ void pos2d(const GemmEpilogue& e, int m, int& py, int& px) {
  const int t = m % e.tokens_per_frame;
  if (t < e.n_special) { py = 0; px = 0; return; }
  const int pp = t - e.n_special;
  py = pp / e.grid_w + 1;
  px = pp % e.grid_w + 1;
}

// Rotate pairs (j, j+R/2) of the R-wide region v[0..R) by angle table row `tab` (cos,sin per frequency).
// q/k LayerNorm parameters staged once per CTA in shared memory (broadcast LDS.128 instead of one LDG per element)
constexpr int ROPE_SP = 64;   // positions per frequency in the shared-memory copy of the 2-D RoPE table
struct EpiSmem { float qn_w[128], qn_b[128], kn_w[128], kn_b[128]; };
struct EpiSmemRope { EpiSmem p; float2 rope[32 * 64]; };   // + the 2-D RoPE table, transposed: [frequency (<= 32)][position (<= 64)]

__device__ __forceinline__ void stage_epi_params(EpiSmem* sp, const GemmEpilogue& e, int hd) {
  const int t = threadIdx.x;
  if (t < hd) {
    sp->qn_w[t] = e.qn_w ? __ldg(e.qn_w + t) : 1.f; sp->qn_b[t] = e.qn_b ? __ldg(e.qn_b + t) : 0.f;
    sp->kn_w[t] = e.kn_w ? __ldg(e.kn_w + t) : 1.f; sp->kn_b[t] = e.kn_b ? __ldg(e.kn_b + t) : 0.f;
  }
}
// positions a 2-D RoPE launch can ask for: rows 0 .. max(last grid row, grid width) of the table
__device__ __forceinline__ int rope2d_positions(const GemmEpilogue& e) {
  const int rows = (e.tokens_per_frame - e.n_special + e.grid_w - 1) / e.grid_w;
  return (rows > e.grid_w ? rows : e.grid_w) + 1;
}
// returns the shared-memory table, or nullptr when the launch has no 2-D RoPE / more positions than the copy holds.
// (The table is engine set-up data written long before this launch: reading it ahead of griddepcontrol.wait is safe.)
__device__ __forceinline__ const float2* stage_rope_table(EpiSmemRope* sp, const GemmEpilogue& e, int hd) {
  if (e.rope_mode != ROPE_2D) return nullptr;
  const int npos = rope2d_positions(e), nf = hd / 4;
  if (npos > ROPE_SP) return nullptr;
  for (int i = threadIdx.x; i < npos * nf; i += blockDim.x) {
    const int pos = i / nf, f = i - pos * nf;
    sp->rope[f * ROPE_SP + pos] = __ldg(e.rope_tab + i);
  }
  return sp->rope;
}

// Rotate pairs (j, j+R/2) of the R-wide region v[0..R) by the angle table row `tab4` ((cos,sin) per frequency, two
// frequencies per 16-byte load).
template <int R>
__device__ __forceinline__ void rope_region(float* v, const float4* __restrict__ tab4) {
#pragma unroll
  for (int j = 0; j < R / 2; j += 2) {
    const float4 cs = __ldg(tab4 + j / 2);
    const float a0 = v[j], b0 = v[j + R / 2], a1 = v[j + 1], b1 = v[j + 1 + R / 2];
    v[j] = a0 * cs.x - b0 * cs.y;
    v[j + R / 2] = b0 * cs.x + a0 * cs.y;
    v[j + 1] = a1 * cs.z - b1 * cs.w;
    v[j + 1 + R / 2] = b1 * cs.z + a1 * cs.w;
  }
}

// The same rotation from the TRANSPOSED shared-memory copy of the table ([frequency][position], ROPE_SP positions per frequency):
// the 32 rows of a warp are consecutive tokens, i.e. consecutive x positions, so a lane-per-row read of the global
// [position][frequency] table touches 32 different 128-byte lines per instruction — measured: 10.8 of the 81 us of the
// qkv + LayerNorm + RoPE GEMM (tools/ub_gemm_epi4.py) — while the transposed copy is read 8 consecutive bytes per lane.
template <int R>
__device__ __forceinline__ void rope_region_s(float* v, const float2* tab_t, int pos) {
#pragma unroll
  for (int j = 0; j < R / 2; ++j) {
    const float2 cs = tab_t[j * ROPE_SP + pos];
    const float a0 = v[j], b0 = v[j + R / 2];
    v[j] = a0 * cs.x - b0 * cs.y;
    v[j + R / 2] = b0 * cs.x + a0 * cs.y;
  }
}

// bias + LayerNorm over one head (HD columns, all in this thread) + RoPE, in place.  bias: global (16-byte loads),
// w / b: shared-memory copies of the head LayerNorm parameters.  The epilogue warps are latency-bound (two per SM
// sub-partition, ncu: 22 % issue slots, half of the arithmetic stalled on the long scoreboard), so the reductions run on four
// independent accumulators and the token position (three integer divisions) is computed once per row by the caller.
template <int HD>
__device__ __forceinline__ void head_norm_rope(float* v, const float* __restrict__ bias, const float* w, const float* b,
                                               const GemmEpilogue& e, int py, int px, const float2* rope_s = nullptr) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < HD; i += 4) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + i));
    v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
    s0 += v[i]; s1 += v[i + 1]; s2 += v[i + 2]; s3 += v[i + 3];
  }
  const float mean = ((s0 + s1) + (s2 + s3)) * (1.0f / HD);
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
  for (int i = 0; i < HD; i += 4) {
    const float d0 = v[i] - mean, d1 = v[i + 1] - mean, d2 = v[i + 2] - mean, d3 = v[i + 3] - mean;
    q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
  }
  const float rstd = rsqrtf(((q0 + q1) + (q2 + q3)) * (1.0f / HD) + e.ln_eps);
  const float nmr = -mean * rstd;
#pragma unroll
  for (int i = 0; i < HD; i += 4) {
    const float4 ww = *reinterpret_cast<const float4*>(w + i), bb = *reinterpret_cast<const float4*>(b + i);
    v[i] = fmaf(fmaf(v[i], rstd, nmr), ww.x, bb.x);
    v[i + 1] = fmaf(fmaf(v[i + 1], rstd, nmr), ww.y, bb.y);
    v[i + 2] = fmaf(fmaf(v[i + 2], rstd, nmr), ww.z, bb.z);
    v[i + 3] = fmaf(fmaf(v[i + 3], rstd, nmr), ww.w, bb.w);
  }
  if (e.rope_mode == ROPE_2D && rope_s != nullptr) {
    rope_region_s<HD / 2>(v, rope_s, py);
    rope_region_s<HD / 2>(v + HD / 2, rope_s, px);
  } else if (e.rope_mode == ROPE_2D) {
    rope_region<HD / 2>(v, reinterpret_cast<const float4*>(e.rope_tab + (size_t)py * (HD / 4)));
    rope_region<HD / 2>(v + HD / 2, reinterpret_cast<const float4*>(e.rope_tab + (size_t)px * (HD / 4)));
  } else if (e.rope_mode == ROPE_1D) {
    rope_region<HD>(v, reinterpret_cast<const float4*>(e.rope_tab + (size_t)py * (HD / 2)));   // py carries the 1-D position
  }
}

// EPI_CONV_BF16: k-block -> (channel offset within the tap, A row shifted by the tap offset on the padded grid)
__device__ __forceinline__ void conv_tap_coords(const GemmEpilogue& e, int& a_k, int& a_row) {
  if (e.conv_taps == 9) {
    const int tap = a_k / e.conv_c;
    a_k -= tap * e.conv_c;
    a_row += (tap / 3 - 1) * e.conv_wp + (tap % 3 - 1);
  }
}

template <int BN, int EPI, bool TMA = false>
__device__ __forceinline__ void epilogue_row(const GemmEpilogue& e, const EpiSmem* sp, uint32_t taddr, int m, int n0, bool row_ok, int c_begin = 0, int c_end = BN,
                                             const TmaOut* to = nullptr, const float2* rope_s = nullptr) {
  if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_F32) {
    constexpr int CH = TMA ? 64 : 32;
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += CH) {
      float v[CH];
      __syncwarp();
      load_acc<CH>(taddr + c, v);
      if (row_ok || TMA) {   // TMA: rows past M are computed (garbage) and clipped by the tensor map
        const int n = n0 + c;
#pragma unroll
        for (int i = 0; i < CH; i += 4) {
          const float4 bb = e.bias ? __ldg(reinterpret_cast<const float4*>(e.bias + n + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float x0 = v[i] + bb.x, x1 = v[i + 1] + bb.y, x2 = v[i + 2] + bb.z, x3 = v[i + 3] + bb.w;
          if constexpr (EPI == EPI_BIAS_GELU_BF16) { gelu_erf2(x0, x1); gelu_erf2(x2, x3); }
          v[i] = x0; v[i + 1] = x1; v[i + 2] = x2; v[i + 3] = x3;
        }
        if constexpr (EPI == EPI_BIAS_F32) {
          float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (size_t)m * e.ldo + n);
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else if constexpr (TMA) {
          store_bf16_tma<CH>(*to, v, n);
        } else {
          store_bf16_row<32>(reinterpret_cast<__nv_bfloat16*>(e.out) + (size_t)m * e.ldo + n, v);
        }
      }
    }
  } else if constexpr (EPI == EPI_CONV_BF16) {
    bool border = false;
    if (e.conv_mask) {
      const int rem = m % (e.conv_hp * e.conv_wp);
      const int yp = rem / e.conv_wp, xp = rem - yp * e.conv_wp;
      border = (yp == 0) | (yp == e.conv_hp - 1) | (xp == 0) | (xp == e.conv_wp - 1);
    }
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      float v[32];
      __syncwarp();
      load_acc<32>(taddr + c, v);
      if (row_ok) {
        const int n = n0 + c;
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.out) + (size_t)m * e.ldo + n;
        if (border) {
#pragma unroll
          for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(dst)[i] = make_uint4(0u, 0u, 0u, 0u);
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bb = e.bias ? __ldg(reinterpret_cast<const float4*>(e.bias + n + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
          }
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const void* rp = r == 0 ? e.res1 : e.res2;
            if (rp) {
              const uint4* r4 = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(rp) + (size_t)m * e.ldo + n);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = r4[i];
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { v[8 * i + 2 * j] += ptx::bf16_lo(w[j]); v[8 * i + 2 * j + 1] += ptx::bf16_hi(w[j]); }
              }
            }
          }
          if (e.conv_relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          store_bf16_row<32>(dst, v);
        }
      }
    }
  } else if constexpr (EPI == EPI_RESID_F32) {
    // residual rows are prefetched one 32-column chunk ahead: the fp32 read-modify-write is latency-bound
    float4 rnext[8];
    float4* r4 = row_ok ? reinterpret_cast<float4*>(e.resid + (size_t)m * e.ldr + n0 + c_begin) : nullptr;
    if (row_ok) {
#pragma unroll
      for (int i = 0; i < 8; ++i) rnext[i] = r4[i];
    }
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      float v[32];
      float4 rcur[8];
      __syncwarp();
      ptx::tmem_ld_32x32b_x32(taddr + c, reinterpret_cast<uint32_t*>(v));
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rcur[i] = rnext[i];
        if (c + 32 < c_end) {
#pragma unroll
          for (int i = 0; i < 8; ++i) rnext[i] = r4[8 + i];
        }
      }
      ptx::tmem_ld_wait();
      if (row_ok) {
        const int n = n0 + c;
        float4* o2 = e.out2 ? reinterpret_cast<float4*>(e.out2 + (size_t)m * e.ld2 + n) : nullptr;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 r = rcur[i];
          const float4 g = e.gamma ? __ldg(reinterpret_cast<const float4*>(e.gamma + n) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
          const float4 bb = e.bias ? __ldg(reinterpret_cast<const float4*>(e.bias + n) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          r.x += g.x * (v[4 * i + 0] + bb.x);
          r.y += g.y * (v[4 * i + 1] + bb.y);
          r.z += g.z * (v[4 * i + 2] + bb.z);
          r.w += g.w * (v[4 * i + 3] + bb.w);
          r4[i] = r;
          if (o2) o2[i] = r;
        }
        r4 += 8;
      }
    }
  } else {  // EPI_QKV_NORM_ROPE_64 / _128 : columns [0,n_q) q heads, [n_q, n_q+n_k) k heads, remainder plain (+bias)
    constexpr int HD = (EPI == EPI_HEADNORM64_BF16) ? 64 : 128;
    int py = 0, px = 0;   // RoPE position of this row: once per tile, not once per head
    if (e.rope_mode == ROPE_2D) pos2d(e, m, py, px);
    else if (e.rope_mode == ROPE_1D) py = __ldg(e.pos_ids + (m % e.pos_period));
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += HD) {
      float v[HD];
      __syncwarp();
      load_acc<HD>(taddr + c, v);
      if (!row_ok && !TMA) continue;  // reconverges at the __syncwarp above / after the loop (TMA: computed and clipped)
      const int n = n0 + c;
      if (n < e.n_q_cols) {
        head_norm_rope<HD>(v, e.bias + n, sp->qn_w, sp->qn_b, e, py, px, rope_s);
      } else if (n < e.n_q_cols + e.n_k_cols) {
        head_norm_rope<HD>(v, e.bias + n, sp->kn_w, sp->kn_b, e, py, px, rope_s);
      } else {
#pragma unroll
        for (int i = 0; i < HD; i += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(e.bias + n + i));
          v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
        }
      }
      if constexpr (TMA) store_bf16_tma<HD>(*to, v, n);
      else store_bf16_row<HD>(reinterpret_cast<__nv_bfloat16*>(e.out) + (size_t)m * e.ldo + n, v);
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                  GemmEpilogue epi) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  EpiSmem* sp = reinterpret_cast<EpiSmem*>(smem + L::PARAM_OFFSET);
  if constexpr (EPI == EPI_HEADNORM64_BF16 || EPI == EPI_HEADNORM128_BF16) stage_epi_params(sp, epi, EPI == EPI_HEADNORM64_BF16 ? 64 : 128);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_n = N / BN;
  const int n_tiles_m = (M + BM - 1) / BM;
  const int n_tiles = n_tiles_m * n_tiles_n;
  const int n_kb = K / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
#pragma unroll
    for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(full_bar + i, 1); ptx::mbar_init(empty_bar + i, 1); }
#pragma unroll
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(tmem_full + i, 1); ptx::mbar_init(tmem_empty + i, 4); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 2 * BN);
    ptx::tmem_relinquish();
  }
  pdl_launch_dependents();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its results are visible from here on

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles_n) * BM, n0 = (tile % n_tiles_n) * BN;
      for (int kb = 0; kb < n_kb; ++kb) {
        ptx::mbar_wait(empty_bar + stage, phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          ptx::mbar_expect_tx(full_bar + stage, L::STAGE_BYTES);
          int a_k = kb * BK, a_row = m0;
          if constexpr (EPI == EPI_CONV_BF16) conv_tap_coords(epi, a_k, a_row);
          ptx::tma_load_2d(sa, &tmA, full_bar + stage, a_k, a_row);
          ptx::tma_load_2d(sa + L::A_BYTES, &tmB, full_bar + stage, kb * BK, n0);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      ptx::mbar_wait(tmem_empty + acc, acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < n_kb; ++kb) {
        ptx::mbar_wait(full_bar + stage, phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t adesc = ptx::umma_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = ptx::umma_desc_sw128(sa + L::A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)  // advance 32 bytes (encoded >> 4) inside the swizzle row per K step
            ptx::umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          ptx::umma_commit(empty_bar + stage);  // frees the smem slot when these MMAs retire
          if (kb == n_kb - 1) ptx::umma_commit(tmem_full + acc);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5 -> TMEM lane quarter warp%4)
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles_n) * BM, n0 = (tile % n_tiles_n) * BN;
      ptx::mbar_wait(tmem_full + acc, acc_phase);
      ptx::tc_fence_after();
      const int m = m0 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
      epilogue_row<BN, EPI>(epi, sp, taddr, m, n0, m < M);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty + acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 2 * BN);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs computes one 256 x 256 tile.  Each CTA stages its own 128 rows
// of A and HALF of the B tile (128 of the 256 weight rows); the leader's single MMA thread issues tcgen05.mma
// .cta_group::2 (M = 256), which reads both halves of B from the two CTAs' shared memory and writes each CTA's 128
// accumulator rows into that CTA's TMEM.  Per CTA and k-block that is 32 KB of L2 traffic instead of 48 KB for the same
// 128x256x64 MACs, and the stage shrinks so that 6 stages fit: both attack what bounds the 1-CTA kernel (L2->SM bytes in
// flight).  Barriers: TMA of both CTAs credits the leader's `full`; `empty` / `tmem_full` are signalled in both CTAs by a
// multicast commit; the peer's epilogue warps arrive remotely on the leader's `tmem_empty`.
constexpr int BN2 = 256;
template <int EPI>
struct Smem2 {
  // the fp32-residual epilogue stages 32x32 accumulator blocks in shared memory for TMA reduce-add and gives up one
  // pipeline stage for the staging tiles
  static constexpr bool TRANSPOSE = (EPI == EPI_RESID_F32);
  // bf16 outputs are staged the same way (32 x 64 bf16 tile per epilogue warp) for a plain TMA store
  static constexpr bool TMA_BF16 = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_HEADNORM64_BF16 || EPI == EPI_HEADNORM128_BF16);
  static constexpr bool TILES = TRANSPOSE || TMA_BF16;
  static constexpr int STAGES = TILES ? 5 : 6;
  static constexpr int A_BYTES = BM * BK * 2;          // 16 KB: this CTA's 128 rows
  static constexpr int B_BYTES = (BN2 / 2) * BK * 2;   // 16 KB: this CTA's half of the weight tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int PARAM_OFFSET = BAR_OFFSET + 256;
  static constexpr bool HEADNORM = (EPI == EPI_HEADNORM64_BF16 || EPI == EPI_HEADNORM128_BF16);
  static constexpr int PARAM_BYTES = TRANSPOSE ? 0 : (HEADNORM ? (int)sizeof(EpiSmemRope) : 2048);   // q/k LayerNorm parameters + transposed RoPE table (head-norm epilogues only)
  static constexpr int TILE_OFFSET = (PARAM_OFFSET + PARAM_BYTES + 1023) / 1024 * 1024;  // 128B-swizzled tiles: 1024-byte aligned
  static constexpr int TILE_BYTES = 32 * 32 * 4;       // one 32 x 32 fp32 block (or 32 x 64 bf16)
  static constexpr int TILE_BUFS = TRANSPOSE ? 2 : 1;  // the residual epilogue keeps two tiles per warp (load / store double buffer)
  static constexpr int TOTAL = TILE_OFFSET + (TILES ? 8 * TILE_BUFS * TILE_BYTES : 0) + 1024;
  static_assert(TOTAL <= 232448, "shared memory budget");
};

// resid[m, n] += gamma[n] * (acc[m, n] + bias[n]) without reading the residual in the SM: each warp scales its 32 x 32
// accumulator block, writes it (128B-swizzled, conflict-free) into a 4 KB shared-memory tile and fires one TMA
// reduce-add (cp.reduce.async.bulk.tensor ... .add) at the fp32 residual stream; the add happens in L2, rows past M are
// clipped by the tensor map.  The row-per-lane read-modify-write it replaces spent its time in long-scoreboard stalls
// (profiles/: proj GEMM 57 us vs 29 us with a plain bf16 epilogue).
__device__ __forceinline__ void epilogue_resid_tma(const GemmEpilogue& e, const CUtensorMap* tmR, uint8_t* tile, uint32_t taddr,
                                                   int m_base, int n0, int c_begin, int c_end, bool with_bias = true) {
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int c = c_begin; c < c_end; c += 32) {
    float v[32];
    __syncwarp();
    load_acc<32>(taddr + c, v);
    const int n = n0 + c;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 g = e.gamma ? __ldg(reinterpret_cast<const float4*>(e.gamma + n) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 bb = (e.bias && with_bias) ? __ldg(reinterpret_cast<const float4*>(e.bias + n) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      v[4 * i + 0] = g.x * (v[4 * i + 0] + bb.x);
      v[4 * i + 1] = g.y * (v[4 * i + 1] + bb.y);
      v[4 * i + 2] = g.z * (v[4 * i + 2] + bb.z);
      v[4 * i + 3] = g.w * (v[4 * i + 3] + bb.w);
    }
    if (lane == 0) ptx::tma_store_wait_read<0>();  // the previous reduce has finished reading the tile
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(tile + lane * 128 + ((i ^ (lane & 7)) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_reduce_add_2d(tmR, tile, n, m_base);
      ptx::tma_store_commit();
    }
  }
}

// The same update without shared memory (LSVS_GEMM_RESID_RED=1, A/B): the warp transposes its 32 x 32 accumulator block in registers
// (5 butterfly stages of 16 shuffles) so that a lane owns one COLUMN, and issues one red.global.add.f32 per row — 128 contiguous
// bytes per warp instruction, added in L2 like the TMA reduction, but the 2 x 128 KB per tile of staging traffic (st.shared + the
// TMA engine's read) never touch the shared-memory pipe that the main loop saturates.
__device__ __forceinline__ void epilogue_resid_red(const GemmEpilogue& e, uint32_t taddr, int m_base, int M, int n0, int c_begin, int c_end,
                                                   bool with_bias = true) {
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int c = c_begin; c < c_end; c += 32) {
    float v[32];
    __syncwarp();
    load_acc<32>(taddr + c, v);
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if ((j & s) == 0) {
          const float send = (lane & s) ? v[j] : v[j | s];
          const float recv = __shfl_xor_sync(0xffffffffu, send, s);
          if (lane & s) v[j] = recv; else v[j | s] = recv;
        }
      }
    }
    // now v[j] = element (row j, column lane) of the block
    const int n = n0 + c + lane;
    const float g = e.gamma ? __ldg(e.gamma + n) : 1.f;
    const float bb = (e.bias && with_bias) ? __ldg(e.bias + n) : 0.f;
    float* dst = e.resid + (size_t)m_base * e.ldr + n;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (m_base + j < M) {
        const float x = g * (v[j] + bb);
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + (size_t)j * e.ldr), "f"(x) : "memory");
      }
    }
  }
}

template <int EPI, int EW>  // EW epilogue warps per CTA (4 or 8)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
gemm_bf16_tcgen05_2cta(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmR, int M, int N, int K, GemmEpilogue epi, int use_tma_reduce) {
  using L = Smem2<EPI>;
  constexpr int STAGES2 = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES2;
  uint64_t* tmem_full = empty_bar + STAGES2;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* resid_bar = tmem_empty + 3;   // 2 per epilogue warp: residual tiles landed (load-add-store epilogue)
  static_assert(!L::TRANSPOSE || (STAGES2 * 2 + 4 + 1 + 2 * EW) * 8 <= 256, "barrier area");
  EpiSmem* sp = reinterpret_cast<EpiSmem*>(smem + L::PARAM_OFFSET);
  const float2* rope_s = nullptr;
  if constexpr (EPI == EPI_HEADNORM64_BF16 || EPI == EPI_HEADNORM128_BF16) {
    stage_epi_params(sp, epi, EPI == EPI_HEADNORM64_BF16 ? 64 : 128);
    rope_s = stage_rope_table(reinterpret_cast<EpiSmemRope*>(sp), epi, EPI == EPI_HEADNORM64_BF16 ? 64 : 128);
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = ptx::cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_tiles_n = N / BN2;
  const int n_tiles_m = (M + 2 * BM - 1) / (2 * BM);
  // split-K (reduce-add residual epilogue only, few tiles): work item = (tile, K slice); every slice adds its partial
  // gamma * acc into the residual (L2 atomics of the TMA reduce), slice 0 also carries the bias
  const int splits = (use_tma_reduce >> 8) > 1 ? (use_tma_reduce >> 8) : 1;
  const int n_tiles = n_tiles_m * n_tiles_n * splits;   // work items
  const int n_kb = (K / BK) / splits;                   // k-blocks per work item

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
#pragma unroll
    for (int i = 0; i < STAGES2; ++i) { ptx::mbar_init(full_bar + i, 2); ptx::mbar_init(empty_bar + i, 1); }
#pragma unroll
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(tmem_full + i, 1); ptx::mbar_init(tmem_empty + i, 2 * EW); }
    if constexpr (L::TRANSPOSE) {
#pragma unroll
      for (int i = 0; i < 2 * EW; ++i) ptx::mbar_init(resid_bar + i, 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(tmem_slot, 2 * BN2);
    ptx::tmem_relinquish_2sm();
  }
  pdl_launch_dependents();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its results are visible from here on

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int item = pair; item < n_tiles; item += n_pairs) {
      const int tile = item / splits, kb0 = (item - tile * splits) * n_kb;
      const int m0 = (tile / n_tiles_n) * (2 * BM) + (int)cta * BM;
      const int n0 = (tile % n_tiles_n) * BN2 + (int)cta * (BN2 / 2);
      for (int kb = kb0; kb < kb0 + n_kb; ++kb) {
        ptx::mbar_wait(empty_bar + stage, phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
#ifdef LSVS_MEASURE
          const bool skip_b = (use_tma_reduce & 2) != 0;  // lsvs_debug_gemm_mode 3: halves the L2->SM bytes, wrong results
#else
          constexpr bool skip_b = false;
#endif
          if (cta == 0) ptx::mbar_expect_tx(full_bar + stage, skip_b ? 2 * L::A_BYTES : 2 * L::STAGE_BYTES);  // bytes of both CTAs land here
          else ptx::mbar_arrive_remote(full_bar + stage, 0);
          int a_k = kb * BK, a_row = m0;
          if constexpr (EPI == EPI_CONV_BF16) conv_tap_coords(epi, a_k, a_row);
          ptx::tma_load_2d_2sm(sa, &tmA, full_bar + stage, a_k, a_row);
          if (!skip_b) ptx::tma_load_2d_2sm(sa + L::A_BYTES, &tmB, full_bar + stage, kb * BK, n0);
        }
        __syncwarp();
        if (++stage == STAGES2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (cta == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, BN2, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        ptx::mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN2;
        for (int kb = 0; kb < n_kb; ++kb) {
          ptx::mbar_wait(full_bar + stage, phase);
          ptx::tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = ptx::smem_u32(smem + stage * L::STAGE_BYTES);
            const uint64_t adesc = ptx::umma_desc_sw128(sa, 16, 1024);
            const uint64_t bdesc = ptx::umma_desc_sw128(sa + L::A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              ptx::umma_bf16_ss_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            ptx::umma_commit_2sm(empty_bar + stage, 3);
            if (kb == n_kb - 1) ptx::umma_commit_2sm(tmem_full + acc, 3);
          }
          __syncwarp();
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (EW warps: lane quarter = warp % 4, column half = (warp-2)/4)
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;
    constexpr int COLS = BN2 / (EW / 4);
    int acc = 0;
    uint32_t acc_phase = 0;
    if (L::TRANSPOSE && (use_tma_reduce & 16)) {
      // ---- alternative residual epilogue without L2 reductions (LSVS_GEMM_RESID_LOADSTORE=1): resid tile -> shared memory (TMA
      // load, issued one 32-column chunk ahead), add gamma * (acc + bias) in place, TMA store back.  Measured slower than the
      // reduce-add form on B200 (see launch2); kept for A/B runs.
      // One split only (every element is written once per GEMM, so the prefetched tile of the NEXT output tile is never stale).
      uint8_t* tiles = smem + L::TILE_OFFSET + (warp - 2) * 2 * L::TILE_BYTES;
      uint64_t* lbar = resid_bar + 2 * (warp - 2);
      constexpr int NCH = COLS / 32;
      auto chunk_coords = [&](int item, int c, int& n, int& mb) {
        mb = (item / n_tiles_n) * (2 * BM) + (int)cta * BM + quarter * 32;
        n = (item % n_tiles_n) * BN2 + part * COLS + 32 * c;
      };
      uint32_t g = 0;   // chunks processed by this warp: buffer g & 1, barrier phase (g >> 1) & 1
      if (pair < n_tiles && lane == 0) {
        int n, mb;
        chunk_coords(pair, 0, n, mb);
        ptx::mbar_expect_tx(lbar, L::TILE_BYTES);
        ptx::tma_load_2d(tiles, &tmR, lbar, n, mb);
      }
      for (int item = pair; item < n_tiles; item += n_pairs) {
        ptx::mbar_wait(tmem_full + acc, acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN2;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c, ++g) {
          int n, mb;
          chunk_coords(item, c, n, mb);
          if (lane == 0) {   // prefetch the next chunk's residual tile into the other buffer once its last store has read it
            const bool next_here = c + 1 < NCH, next_item = item + n_pairs < n_tiles;
            if (next_here || next_item) {
              int nn, nmb;
              chunk_coords(next_here ? item : item + n_pairs, next_here ? c + 1 : 0, nn, nmb);
              ptx::tma_store_wait_read<0>();
              ptx::mbar_expect_tx(lbar + ((g + 1) & 1), L::TILE_BYTES);
              ptx::tma_load_2d(tiles + ((g + 1) & 1) * L::TILE_BYTES, &tmR, lbar + ((g + 1) & 1), nn, nmb);
            }
          }
          float v[32];
          __syncwarp();
          load_acc<32>(taddr + part * COLS + 32 * c, v);
          ptx::mbar_wait(lbar + (g & 1), (g >> 1) & 1);
          uint8_t* tile = tiles + (g & 1) * L::TILE_BYTES;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4* slot = reinterpret_cast<float4*>(tile + lane * 128 + ((i ^ (lane & 7)) << 4));
            float4 r = *slot;
            const float4 gm = epi.gamma ? __ldg(reinterpret_cast<const float4*>(epi.gamma + n) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
            const float4 bb = epi.bias ? __ldg(reinterpret_cast<const float4*>(epi.bias + n) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            r.x += gm.x * (v[4 * i + 0] + bb.x); r.y += gm.y * (v[4 * i + 1] + bb.y);
            r.z += gm.z * (v[4 * i + 2] + bb.z); r.w += gm.w * (v[4 * i + 3] + bb.w);
            *slot = r;
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmR, tile, n, mb);
            ptx::tma_store_commit();
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (cta == 0) ptx::mbar_arrive(tmem_empty + acc);
          else ptx::mbar_arrive_remote(tmem_empty + acc, 0);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    } else
    for (int item = pair; item < n_tiles; item += n_pairs) {
      const int tile = item / splits, slice = item - tile * splits;
      const int m0 = (tile / n_tiles_n) * (2 * BM) + (int)cta * BM;
      const int n0 = (tile % n_tiles_n) * BN2;
      ptx::mbar_wait(tmem_full + acc, acc_phase);
      ptx::tc_fence_after();
      const int m = m0 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN2;
      if (L::TRANSPOSE && (use_tma_reduce & 32)) {
        epilogue_resid_red(epi, taddr, m0 + quarter * 32, M, n0, part * COLS, (part + 1) * COLS, slice == 0);
      } else if (L::TRANSPOSE && (use_tma_reduce & 1)) {
        epilogue_resid_tma(epi, &tmR, smem + L::TILE_OFFSET + (warp - 2) * L::TILE_BUFS * L::TILE_BYTES, taddr, m0 + quarter * 32, n0, part * COLS, (part + 1) * COLS, slice == 0);
      } else if (L::TMA_BF16 && (use_tma_reduce & 4)) {
        const TmaOut to{&tmR, smem + L::TILE_OFFSET + (warp - 2) * L::TILE_BYTES, m0 + quarter * 32};
        epilogue_row<BN2, EPI, L::TMA_BF16>(epi, sp, taddr, m, n0, m < M, part * COLS, (part + 1) * COLS, &to, rope_s);
      } else {
        epilogue_row<BN2, EPI>(epi, sp, taddr, m, n0, m < M, part * COLS, (part + 1) * COLS, nullptr, rope_s);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta == 0) ptx::mbar_arrive(tmem_empty + acc);
        else ptx::mbar_arrive_remote(tmem_empty + acc, 0);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  if (((L::TRANSPOSE && (use_tma_reduce & 1) && !(use_tma_reduce & 32)) || (L::TMA_BF16 && (use_tma_reduce & 4))) && warp >= 2 && lane == 0) ptx::tma_store_wait<0>();  // reductions issued by this lane are complete
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 1) ptx::tmem_dealloc_2sm(tmem_base, 2 * BN2);
}

template <int EPI>
int launch2(const CUtensorMap* tmA, const CUtensorMap* tmB, int M, int N, int K, const GemmEpilogue& e, cudaStream_t st) {
  // epilogue warps: 4 (255 registers each) for the 128-wide LayerNorm+RoPE epilogue, 8 otherwise (16 measured: no gain; the
  // 64-wide LayerNorm+RoPE epilogue fits 168 registers since its output is staged for the TMA store: 81.9 -> 77.7 us)
  constexpr int EW = (EPI == EPI_HEADNORM128_BF16) ? 4 : 8;
  auto kern = gemm_bf16_tcgen05_2cta<EPI, EW>;
  static bool configured = false;
  if (!configured) {
    LSVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem2<EPI>::TOTAL));
    configured = true;
  }
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * (N / BN2);
  const int max_pairs = num_sms() / 2;
  const CUtensorMap* tmR = tmA;
  int use_red = 0;
  if (EPI == EPI_RESID_F32 && e.out2 == nullptr && g_gemm_mode != 2 && ((uintptr_t)e.resid % 16 == 0) && (e.ldr % 4 == 0)) {
    tmR = tmap_2d_f32_box32(e.resid, (uint64_t)N, (uint64_t)M, (uint64_t)e.ldr * 4, 32);
    if (!tmR) return LSVS_ECUDA;
    use_red = 1;
  }
  if (Smem2<EPI>::TMA_BF16 && g_gemm_mode != 2 && ((uintptr_t)e.out % 16 == 0) && (e.ldo % 8 == 0)) {
    tmR = tmap_2d_bf16(e.out, (uint64_t)N, (uint64_t)M, (uint64_t)e.ldo * 2, 64, 32);
    if (!tmR) return LSVS_ECUDA;
    use_red |= 4;
  }
  if (g_gemm_mode == 3) use_red |= 2;
  int splits = 1;
  if ((use_red & 1) && g_gemm_mode == 0 && !deterministic()) {  // few tiles (short chunks): slice K while all slices still fit in one round of CTA pairs
    const int n_kb = K / BK;
    while (splits < 8 && tiles * splits * 2 <= max_pairs && n_kb % (2 * splits) == 0 && n_kb / (2 * splits) >= 4) splits *= 2;
  }
  // LSVS_GEMM_RESID_LOADSTORE=1 (A/B runs): load-add-store epilogue instead of the TMA reduce-add.  Measured on B200: slower
  // (projection 34.3 vs 33.1 us, fc2 89.5 vs 85.5 us, 398 vs 404 frames/s in the pipeline — profiles/r2_gemm_resid_ab.md), so the
  // L2 reduction units are not what holds the residual epilogue back; the reduce-add stays the default.
  static const bool resid_loadstore = [] { const char* v = getenv("LSVS_GEMM_RESID_LOADSTORE"); return v && atoi(v) != 0; }();
  if ((use_red & 1) && splits == 1 && resid_loadstore) use_red |= 16;
  // LSVS_GEMM_RESID_RED=1 (A/B runs): register transpose + red.global.add.f32, no shared memory (epilogue_resid_red).  Measured slower
  // too (projection 35.3 vs 33.6 us, step 56.9 vs 56.1 ms): the K = 1024 residual GEMM is HBM-bound by the 54 MB fp32 residual stream
  // (135 MB per launch once the stream has left L2), not by its epilogue.
  static const bool resid_red = [] { const char* v = getenv("LSVS_GEMM_RESID_RED"); return v && atoi(v) != 0; }();
  if ((use_red & 1) && !(use_red & 16) && resid_red) use_red |= 32;
  use_red |= splits << 8;
  const int items = tiles * splits;
  const int pairs = items < max_pairs ? items : max_pairs;
  LSVS_CUDA(launch_pdl(kern, dim3(2 * pairs), dim3(64 + 32 * EW), Smem2<EPI>::TOTAL, st, *tmA, *tmB, *tmR, M, N, K, e, use_red));
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Few rows (M <= 128; the camera-head trunk runs on the S frames of a chunk: 32 rows, D = 2048): the GEMM is WEIGHT STREAMING —
// 2 bytes of W per 2 M flops — and bounded by HBM, not by the tensor cores (UPSTREAM CameraHead trunk: 4 iterations x 4 blocks
// = 1.6 GB of bf16 weights per chunk).  The operands are swapped so that W is the 128-row side of the MMA,
//     D^T[128 weight rows, MT rows of A] (+)= W_tile[128, 64] . A_tile[MT, 64]^T          (MT = 32 / 64 / 128 >= M)
// so every byte staged for the tensor core is payload (the 128 x 64 tiles of the kernel above pad the A side to 128 rows — two
// thirds of the shared-memory fill are zeros — and N / 64 tiles keep only 32..128 SMs loading).  Work item = (128 weight rows,
// K slice); the slices of one tile form a thread-block CLUSTER (up to 16 CTAs, two CTAs per SM): each CTA streams its K range
// through a 5-stage TMA ring, leaves its fp32 partial tile in its own shared memory, and after one cluster barrier CTA r sums
// rows [r, r+1) * MT / slices of all partials over distributed shared memory IN SLICE ORDER (bit-reproducible: no atomics, no
// workspace) and runs the fused epilogue on them; the transposed tile makes those global stores coalesced (lanes = consecutive n).
// Inside a ConstWeights scope (the engine's forward passes: W are model weights packed at load time) the W loads of the first ring
// stages are issued BEFORE griddepcontrol.wait; a caller of the bare C ABI may have produced W with the previous kernel, so there
// they wait like every other load.
template <int MT>
struct SmemFewRows {
  static constexpr int STAGES = MT == 32 ? 5 : (MT == 64 ? 4 : 3);
  static constexpr int W_BYTES = 128 * BK * 2;          // 16 KB of weights per k-block
  static constexpr int X_BYTES = MT * BK * 2;
  static constexpr int STAGE_BYTES = W_BYTES + X_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
  static_assert(STAGES * STAGE_BYTES >= MT * 128 * 4, "the partial tile reuses the ring");
  static_assert(2 * TOTAL <= 232448, "two CTAs per SM");
};

__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t cta) {
  float v;
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %1, %2;\n\tld.shared::cluster.f32 %0, [ra];\n\t}" : "=f"(v) : "r"(local_addr), "r"(cta) : "memory");
  return v;
}

template <int MT, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 2)
gemm_fewrows_tcgen05(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, int M, int N, int K,
                     GemmEpilogue epi, int w_is_constant) {
  using L = SmemFewRows<MT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + L::STAGES;
  uint64_t* acc_bar = empty_bar + L::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x, slices = gridDim.x;   // cluster = the K slices of one tile: cluster rank == blockIdx.x
  const int n0 = blockIdx.y * 128;
  const int n_kb_all = K / BK;
  const int kb0 = (int)((long long)slice * n_kb_all / slices);
  const int n_kb = (int)((long long)(slice + 1) * n_kb_all / slices) - kb0;   // >= 1 (host: slices <= K / BK)
  const int pre = n_kb < L::STAGES ? n_kb : L::STAGES;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmX);
#pragma unroll
    for (int i = 0; i < L::STAGES; ++i) { ptx::mbar_init(full_bar + i, 1); ptx::mbar_init(empty_bar + i, 1); }
    ptx::mbar_init(acc_bar, 1);
    ptx::fence_mbar_init();
    if (w_is_constant) {
      for (int i = 0; i < pre; ++i) {   // weights first: they do not depend on the predecessor in the stream
        ptx::mbar_expect_tx(full_bar + i, L::STAGE_BYTES);
        ptx::tma_load_2d(smem + i * L::STAGE_BYTES, &tmW, full_bar + i, (kb0 + i) * BK, n0);
      }
    }
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, MT);
    ptx::tmem_relinquish();
  }
  pdl_launch_dependents();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < pre; ++i) {
        if (!w_is_constant) {
          ptx::mbar_expect_tx(full_bar + i, L::STAGE_BYTES);
          ptx::tma_load_2d(smem + i * L::STAGE_BYTES, &tmW, full_bar + i, (kb0 + i) * BK, n0);
        }
        ptx::tma_load_2d(smem + i * L::STAGE_BYTES + L::W_BYTES, &tmX, full_bar + i, (kb0 + i) * BK, 0);
      }
    }
    __syncwarp();
    for (int i = pre; i < n_kb; ++i) {
      const int stage = i % L::STAGES;
      ptx::mbar_wait(empty_bar + stage, ((i / L::STAGES) & 1) ^ 1);
      if (lane == 0) {
        uint8_t* s = smem + stage * L::STAGE_BYTES;
        ptx::mbar_expect_tx(full_bar + stage, L::STAGE_BYTES);
        ptx::tma_load_2d(s, &tmW, full_bar + stage, (kb0 + i) * BK, n0);
        ptx::tma_load_2d(s + L::W_BYTES, &tmX, full_bar + stage, (kb0 + i) * BK, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, MT, 0, 0);
    for (int i = 0; i < n_kb; ++i) {
      const int stage = i % L::STAGES;
      ptx::mbar_wait(full_bar + stage, (i / L::STAGES) & 1);
      ptx::tc_fence_after();
      if (lane == 0) {
        const uint32_t s = ptx::smem_u32(smem + stage * L::STAGE_BYTES);
        const uint64_t adesc = ptx::umma_desc_sw128(s, 16, 1024);
        const uint64_t bdesc = ptx::umma_desc_sw128(s + L::W_BYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) ptx::umma_bf16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0);
        ptx::umma_commit(empty_bar + stage);
        if (i == n_kb - 1) ptx::umma_commit(acc_bar);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ partial tile -> shared memory (ring storage: every MMA has retired)
    const int quarter = warp & 3;
    ptx::mbar_wait(acc_bar, 0);
    ptx::tc_fence_after();
    float* part = reinterpret_cast<float*>(smem) + quarter * 32 + lane;   // [MT][128]: lanes = consecutive weight rows
#pragma unroll
    for (int c = 0; c < MT; c += 32) {
      float v[32];
      load_acc<32>(tmem_base + ((uint32_t)(quarter * 32) << 16) + c, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) part[(c + j) * 128] = v[j];
    }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();   // every slice's partial is in its CTA's shared memory
  if (warp >= 2) {
    // ------------------------------------------------------------ slice-ordered sum over the cluster + fused epilogue
    const int nl = (warp & 3) * 32 + lane, n = n0 + nl;
    const int m_begin = (int)((long long)slice * MT / slices), m_end_t = (int)((long long)(slice + 1) * MT / slices);
    const int m_end = m_end_t < M ? m_end_t : M;
    const float bias = epi.bias ? __ldg(epi.bias + n) : 0.f;
    const float gamma = (EPI == EPI_RESID_F32 && epi.gamma) ? __ldg(epi.gamma + n) : 1.f;
    const uint32_t base = ptx::smem_u32(reinterpret_cast<float*>(smem) + nl);
    constexpr int MAXS = 16, RG = 4;   // four rows at a time: up to 64 distributed-shared-memory loads in flight per thread
    for (int m = m_begin; m < m_end; m += RG) {
      float rv[RG], p[RG][MAXS];
      if constexpr (EPI == EPI_RESID_F32) {
#pragma unroll
        for (int j = 0; j < RG; ++j) rv[j] = (m + j < m_end) ? epi.resid[(size_t)(m + j) * epi.ldr + n] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < RG; ++j)
#pragma unroll
        for (int s = 0; s < MAXS; ++s) p[j][s] = (m + j < m_end && s < slices) ? ld_dsmem_f32(base + (uint32_t)(m + j) * 512u, (uint32_t)s) : 0.f;
#pragma unroll
      for (int j = 0; j < RG; ++j) {
        if (m + j >= m_end) break;
        float acc = p[j][0];
#pragma unroll
        for (int s = 1; s < MAXS; ++s) acc += p[j][s];   // slice order; absent slices add +0
        float x = acc + bias;
        if constexpr (EPI == EPI_BIAS_GELU_BF16) x = gelu_erf(x);
        const size_t row = (size_t)(m + j);
        if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16) {
          reinterpret_cast<__nv_bfloat16*>(epi.out)[row * epi.ldo + n] = __float2bfloat16_rn(x);
        } else if constexpr (EPI == EPI_BIAS_F32) {
          reinterpret_cast<float*>(epi.out)[row * epi.ldo + n] = x;
        } else {
          const float y = rv[j] + gamma * x;
          epi.resid[row * epi.ldr + n] = y;
          if (epi.out2) epi.out2[row * epi.ld2 + n] = y;
        }
      }
    }
  }
  ptx::cluster_sync();   // no CTA leaves while a peer still reads its partial
  if (warp == 1) ptx::tmem_dealloc(tmem_base, MT);
}

template <int MT, int EPI>
int launch_fewrows(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmEpilogue& e, cudaStream_t st) {
  using L = SmemFewRows<MT>;
  const CUtensorMap* tmW = tmap_2d_bf16(W, K, N, (uint64_t)ldw * 2, BK, 128);
  const CUtensorMap* tmX = tmap_2d_bf16(A, K, M, (uint64_t)lda * 2, BK, MT);
  if (!tmW || !tmX) return LSVS_ECUDA;
  auto kern = gemm_fewrows_tcgen05<MT, EPI>;
  // K slices (= cluster size, a power of two <= 16; 16 is a non-portable cluster size, used only if the occupancy query accepts it):
  // as many as keep every CTA resident at two per SM with at least 4 k-blocks (64 KB of W) each
  static int max_slices = 0;
  if (!max_slices) {
    LSVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    const char* v = getenv("LSVS_GEMM_FEWROWS_SLICES");
    int want = v ? atoi(v) : 8;   // measured: clusters of 16 run the K = 8192 shape at 17.3 us instead of 11.1 (few clusters of that size are co-resident)
    want = want < 1 ? 1 : (want > 16 ? 16 : want);
    if (want > 8) {
      int n_clusters = 0;
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(16, 1); q.blockDim = dim3(NUM_THREADS); q.dynamicSmemBytes = L::TOTAL;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension; qa[0].val.clusterDim.x = 16; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
          cudaOccupancyMaxActiveClusters(&n_clusters, kern, &q) != cudaSuccess || n_clusters < 1) {
        cudaGetLastError();
        want = 8;
      }
    }
    max_slices = want;
  }
  const int tiles = N / 128, n_kb = K / BK, slots = 2 * num_sms();
  int slices = 1;
  while (2 * slices <= max_slices && tiles * 2 * slices <= slots && n_kb / (2 * slices) >= 4) slices *= 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)slices, (unsigned)tiles); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = L::TOTAL; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = (unsigned)slices; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 2;
  LSVS_CUDA(cudaLaunchKernelEx(&cfg, kern, *tmW, *tmX, M, N, K, e, g_const_weights ? 1 : 0));
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

template <int EPI>
int launch_fewrows_mt(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmEpilogue& e, cudaStream_t st) {
  if (M <= 32) return launch_fewrows<32, EPI>(A, lda, W, ldw, M, N, K, e, st);
  if (M <= 64) return launch_fewrows<64, EPI>(A, lda, W, ldw, M, N, K, e, st);
  return launch_fewrows<128, EPI>(A, lda, W, ldw, M, N, K, e, st);
}

template <int BN, int EPI>
int launch(const CUtensorMap* tmA, const CUtensorMap* tmB, int M, int N, int K, const GemmEpilogue& e, cudaStream_t st) {
  using L = SmemLayout<BN>;
  auto kern = gemm_bf16_tcgen05<BN, EPI>;
  static bool configured = false;
  if (!configured) {
    LSVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int tiles = ((M + BM - 1) / BM) * (N / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  LSVS_CUDA(launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), L::TOTAL, st, *tmA, *tmB, M, N, K, e));
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

}  // namespace

int gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int epi_kind, const GemmEpilogue& e,
              cudaStream_t st) {
  LSVS_CHECK_ARG(A && W && M > 0 && N > 0 && K > 0, "gemm: null operand or empty shape (M=%d N=%d K=%d)", M, N, K);
  LSVS_CHECK_ARG(K % BK == 0, "gemm: K=%d must be a multiple of %d", K, BK);
  const bool conv = (epi_kind == EPI_CONV_BF16);
  int a_cols = K;   // width of the A matrix: K, or the channels of one tap for the shifted-row convolution
  if (conv) {
    LSVS_CHECK_ARG(e.out && e.ldo >= N, "gemm: conv epilogue needs out with ldo >= N");
    LSVS_CHECK_ARG(e.conv_taps == 0 || e.conv_taps == 9, "gemm: conv_taps must be 0 or 9");
    if (e.conv_taps == 9) {
      LSVS_CHECK_ARG(e.conv_c > 0 && e.conv_c % BK == 0 && K == 9 * e.conv_c && e.conv_wp > 2, "gemm: bad 3x3 convolution geometry (C=%d, K=%d)", e.conv_c, K);
      a_cols = e.conv_c;
    }
    LSVS_CHECK_ARG(!e.conv_mask || (e.conv_hp > 2 && e.conv_wp > 2 && M % (e.conv_hp * e.conv_wp) == 0), "gemm: border mask needs M = frames * hp * wp");
  }
  LSVS_CHECK_ARG(lda >= a_cols && ldw >= K && lda % 8 == 0 && ldw % 8 == 0, "gemm: leading dimensions must be >= K and 16-byte aligned");
  const bool bn256 = (N % 256 == 0);
  const bool bn64 = (conv || epi_kind == EPI_BIAS_F32) && N == 64;   // full-resolution DPT output convolution (32 channels padded to 64)
  LSVS_CHECK_ARG(bn256 || bn64 || N % 128 == 0, "gemm: N=%d must be a multiple of 128", N);
  bool pair = bn256 && M > 2 * BM && g_gemm_mode != 1;  // CTA pairs (256x256 tiles) once there are enough rows
  bool wide = bn256;                                     // single-CTA kernel: 128x256 tiles, else 128x128
  if (pair && g_gemm_mode == 0 && !(epi_kind == EPI_RESID_F32 && e.out2 == nullptr)) {  // (residual GEMMs: pair kernel + split-K instead)
    // few rows (short chunks: S = 4..8 frames): the 256x256 pair tiles leave SMs idle or waste a whole round; 128x128 single-CTA
    // tiles do a quarter of the work on half the SMs at ~0.6x the per-tile efficiency
    static const double penalty = [] { const char* v = getenv("LSVS_GEMM_NARROW_PENALTY"); return v ? atof(v) : 1.6; }();   // 1.25 / 1.6 / 2.0 measured at 4- and 5-frame chunks: 10.33 / 9.97 / 10.14 and 10.85 / 10.49 / 10.65 ms (profiles/r2d_narrow_penalty.log)
    const int sms = num_sms();
    const long long tiles_pair = (long long)((M + 2 * BM - 1) / (2 * BM)) * (N / BN2);
    const long long tiles_128 = (long long)((M + BM - 1) / BM) * (N / 128);
    const long long rounds_pair = (tiles_pair + sms / 2 - 1) / (sms / 2), rounds_128 = (tiles_128 + sms - 1) / sms;
    if ((double)rounds_128 * 0.5 * penalty < (double)rounds_pair) { pair = false; wide = false; }
  }
  // few rows (camera-head trunk, M = frames of one chunk): weight streaming with swapped operands, K slices over a cluster
  static const bool fewrows = [] { const char* v = getenv("LSVS_GEMM_FEWROWS"); return !(v && v[0] == '0'); }();
  if (fewrows && g_fewrows_kernel && M <= BM && N % 128 == 0 && (epi_kind == EPI_BIAS_BF16 || epi_kind == EPI_BIAS_GELU_BF16 || epi_kind == EPI_BIAS_F32 || epi_kind == EPI_RESID_F32)) {
    LSVS_CHECK_ARG(epi_kind == EPI_RESID_F32 ? (e.resid != nullptr && e.ldr >= N) : (e.out != nullptr && e.ldo >= N), "gemm: missing output / residual");
    ProfScope prof(PROF_GEMM, st, 2.0 * M * (double)N * K, 0);
    switch (epi_kind) {
      case EPI_BIAS_BF16: return launch_fewrows_mt<EPI_BIAS_BF16>(A, lda, W, ldw, M, N, K, e, st);
      case EPI_BIAS_GELU_BF16: return launch_fewrows_mt<EPI_BIAS_GELU_BF16>(A, lda, W, ldw, M, N, K, e, st);
      case EPI_BIAS_F32: return launch_fewrows_mt<EPI_BIAS_F32>(A, lda, W, ldw, M, N, K, e, st);
      default: return launch_fewrows_mt<EPI_RESID_F32>(A, lda, W, ldw, M, N, K, e, st);
    }
  }
  const int BN = wide ? 256 : 128;
  const CUtensorMap* tmA = tmap_2d_bf16(A, a_cols, M, (uint64_t)lda * 2, BK, BM);
  const CUtensorMap* tmB = tmap_2d_bf16(W, K, N, (uint64_t)ldw * 2, BK, pair ? BN2 / 2 : BN);
  if (!tmA || !tmB) return LSVS_ECUDA;
  ProfScope prof(PROF_GEMM, st, 2.0 * M * (double)N * K, 0);
  if (epi_kind == EPI_HEADNORM64_BF16 || epi_kind == EPI_HEADNORM128_BF16) {
    LSVS_CHECK_ARG(e.bias && e.out, "gemm: head-norm epilogue needs bias and out");
    LSVS_CHECK_ARG((e.n_q_cols == 0 || (e.qn_w && e.qn_b)) && (e.n_k_cols == 0 || (e.kn_w && e.kn_b)), "gemm: missing q/k norm weights");
    LSVS_CHECK_ARG(e.rope_mode == ROPE_NONE || e.rope_tab, "gemm: rope table missing");
    LSVS_CHECK_ARG(e.rope_mode != ROPE_2D || (e.tokens_per_frame > 0 && e.grid_w > 0), "gemm: 2-D rope needs the token grid");
    LSVS_CHECK_ARG(e.rope_mode != ROPE_1D || (e.pos_ids && e.pos_period > 0), "gemm: 1-D rope needs position ids");
  }
#define LSVS_GEMM_CASE(KIND)                                                            \
  case KIND:                                                                            \
    if (pair) return launch2<KIND>(tmA, tmB, M, N, K, e, st);                           \
    return wide ? launch<256, KIND>(tmA, tmB, M, N, K, e, st) : launch<128, KIND>(tmA, tmB, M, N, K, e, st);
  // few rows, N / 64 tiles (the round-2 path; LSVS_GEMM_FEWROWS=0 for A/B runs against gemm_fewrows_tcgen05)
  if (M <= BM && N % 64 == 0 && N / 64 >= 16) {
    const CUtensorMap* tmB64 = tmap_2d_bf16(W, K, N, (uint64_t)ldw * 2, BK, 64);
    if (!tmB64) return LSVS_ECUDA;
    switch (epi_kind) {
      case EPI_BIAS_BF16: return launch<64, EPI_BIAS_BF16>(tmA, tmB64, M, N, K, e, st);
      case EPI_BIAS_GELU_BF16: return launch<64, EPI_BIAS_GELU_BF16>(tmA, tmB64, M, N, K, e, st);
      case EPI_BIAS_F32: return launch<64, EPI_BIAS_F32>(tmA, tmB64, M, N, K, e, st);
      case EPI_RESID_F32: return launch<64, EPI_RESID_F32>(tmA, tmB64, M, N, K, e, st);
      default: break;
    }
  }
  if (bn64 && epi_kind == EPI_BIAS_F32) {   // the same convolution in the fp32-class mode (split operands, fp32 out)
    const CUtensorMap* tmB64 = tmap_2d_bf16(W, K, N, (uint64_t)ldw * 2, BK, 64);
    if (!tmB64) return LSVS_ECUDA;
    return launch<64, EPI_BIAS_F32>(tmA, tmB64, M, N, K, e, st);
  }
  switch (epi_kind) {
    LSVS_GEMM_CASE(EPI_BIAS_BF16)
    LSVS_GEMM_CASE(EPI_BIAS_GELU_BF16)
    LSVS_GEMM_CASE(EPI_BIAS_F32)
    LSVS_GEMM_CASE(EPI_RESID_F32)
    LSVS_GEMM_CASE(EPI_HEADNORM64_BF16)
    LSVS_GEMM_CASE(EPI_HEADNORM128_BF16)
    case EPI_CONV_BF16:
      if (bn64) {
        const CUtensorMap* tmB64 = tmap_2d_bf16(W, K, N, (uint64_t)ldw * 2, BK, 64);
        if (!tmB64) return LSVS_ECUDA;
        return launch<64, EPI_CONV_BF16>(tmA, tmB64, M, N, K, e, st);
      }
      if (pair) return launch2<EPI_CONV_BF16>(tmA, tmB, M, N, K, e, st);
      return wide ? launch<256, EPI_CONV_BF16>(tmA, tmB, M, N, K, e, st) : launch<128, EPI_CONV_BF16>(tmA, tmB, M, N, K, e, st);
    default:
      return fail(LSVS_EINVAL, "gemm: unknown epilogue %d", epi_kind);
  }
#undef LSVS_GEMM_CASE
  (void)BN;
}

}  // namespace lsvs
