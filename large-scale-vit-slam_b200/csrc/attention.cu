// Fused flash-style attention  O = softmax(Q K^T * scale) V  on tcgen05 / TMEM / TMA (no mask; ragged tails).
//
// One CTA owns up to two 128-row query tiles (A, B) of one (batch, head) and streams the K/V blocks once for
// both.  Warp roles (10 warps):
//   warp 8      TMA producer: Q tiles once, then K and V blocks through two independent 3-stage rings
//   warp 9      tcgen05.mma issuer: S_t = Q_t K^T (SS, fp32 in TMEM), O_t += P_t V (SS, P from shared memory)
//   warps 0-3   softmax for tile A, warps 4-7 softmax for tile B (registers moved to them with setmaxnreg): one thread per query row (TMEM lane), so the
//               row max / row sum need no shuffles.  exp2 with the softmax scale folded in; running max and sum
//               in registers; O stays in TMEM and is rescaled in place (tcgen05.ld/st) only when a row max moved.
// The two tiles ping-pong on the tensor core: while the softmax warps of one tile work, the MMAs of the other
// run.  S_t(i+1) is issued as soon as the softmax warps have pulled S_t(i) into registers (s_free barrier).
//
// Replaces F.scaled_dot_product_attention behind UPSTREAM Attention (Aggregator frame/global/DINO blocks,
// alignment-head frame blocks alignment_head.py:363, camera-head trunk); q_norm/k_norm/RoPE are already applied
// by the QKV GEMM epilogue (csrc/gemm.cu).
#include "attention.h"
#include "host_common.h"
#include "ptx.cuh"
#include "tensormap.h"

namespace lsvs {
namespace {

constexpr int QT = 128;            // query rows per tile (UMMA M)
// threads = (NT + 1) warpgroups: one softmax warpgroup per query tile, the last one holds the TMA warp, the MMA warp and 2 idle warps
constexpr int KV_STAGES = 2;

template <int HD>
struct Cfg {
  // Two 128-row query tiles per CTA; 128-key blocks at head dim 64, 64-key blocks at head dim 128 (shared memory).
  // (Measured alternative at head dim 64: four tiles x 64-key blocks, i.e. 16 softmax warps: 455 vs 666 TFLOP/s —
  //  the per-block barrier / fence overhead doubles per key and outweighs the extra latency hiding.)
  static constexpr int NT = 2;                              // query tiles per CTA
  static constexpr int BKV = (HD == 64) ? 128 : 64;         // keys per block (UMMA N of S, K of PV)
  static constexpr int NTHREADS = (NT + 1) * 128;
  static constexpr int MAXNREG = 168;                       // launch-time registers / thread (65536 / NTHREADS, multiple of 8)
  static constexpr int KB = HD / 64;                        // 64-element (128 B) column blocks of the head dim
  static constexpr int Q_TILE_BYTES = QT * HD * 2;
  static constexpr int K_TILE_BYTES = BKV * HD * 2;
  static constexpr int V_TILE_BYTES = BKV * HD * 2;
  static constexpr int P_TILE_BYTES = QT * BKV * 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + NT * Q_TILE_BYTES;
  static constexpr int OFF_V = OFF_K + KV_STAGES * K_TILE_BYTES;
  static constexpr int OFF_P = OFF_V + KV_STAGES * V_TILE_BYTES;
  static constexpr int OFF_BAR = OFF_P + 2 * NT * P_TILE_BYTES;   // P double-buffered per query tile
  static constexpr int SMEM = OFF_BAR + 512 + 1024;
  static constexpr int S_COL = 0;                            // TMEM columns: S_A, S_B, O_A, O_B
  static constexpr int O_COL = NT * BKV;
  static constexpr int TMEM_COLS = 512;
  static_assert(O_COL + NT * HD <= 512, "TMEM budget");
  static_assert(OFF_BAR + 1024 + 1024 <= 232448, "shared memory budget");
};

struct Bars {
  uint64_t q_full;
  uint64_t k_full[KV_STAGES], k_empty[KV_STAGES], v_full[KV_STAGES], v_empty[KV_STAGES];
  uint64_t s_full[4], s_free[4], p_ready[4][2], pv_done[4][2];  // [tile][P buffer = iteration parity]
  uint32_t tmem_slot;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100): halves the issue slots of the scale-and-shift and of the row sums
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float f2_lo(unsigned long long v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo; }
__device__ __forceinline__ float f2_hi(unsigned long long v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return hi; }
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// exp2 on the FMA/ALU pipes for a share of the elements (the MUFU pipe is the bottleneck at head dim 64):
// round-to-nearest split x = n + f via the 1.5*2^23 magic add, degree-3 minimax 2^f on [-0.5, 0.5] (max rel. error
// 7.5e-5, far below the bf16 rounding of P), exponent patched in with an integer add.  Two elements per call.
__device__ __forceinline__ void ex2_poly2(unsigned long long x2, float& r0, float& r1) {
  const float MAGIC = 12582912.0f;  // 1.5 * 2^23
  const unsigned long long xc = f2_pack(fmaxf(f2_lo(x2), -125.0f), fmaxf(f2_hi(x2), -125.0f));
  const unsigned long long xf = f2_add(xc, f2_pack(MAGIC, MAGIC));
  const unsigned long long fi = f2_add(xf, f2_pack(-MAGIC, -MAGIC));
  const unsigned long long fr = f2_fma(fi, f2_pack(-1.0f, -1.0f), xc);
  unsigned long long p = f2_fma(f2_pack(0.055171654f, 0.055171654f), fr, f2_pack(0.24261113f, 0.24261113f));
  p = f2_fma(p, fr, f2_pack(0.69326097f, 0.69326097f));
  p = f2_fma(p, fr, f2_pack(0.99992806f, 0.99992806f));
  r0 = __int_as_float(__float_as_int(f2_lo(p)) + (__float_as_int(f2_lo(xf)) << 23));
  r1 = __int_as_float(__float_as_int(f2_hi(p)) + (__float_as_int(f2_hi(xf)) << 23));
}

// register re-balancing between the service warpgroup and the softmax warpgroups (setmaxnreg, warpgroup-wide)
template <int HD> __device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
}
template <int HD> __device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
}

#ifndef LSVS_ATTN_POLY_PAIRS
#define LSVS_ATTN_POLY_PAIRS 0  // of every 4 element pairs, this many use the polynomial path (0 disables)
#endif

#ifdef LSVS_DEBUG_HANG
__device__ int g_dbg_iter[64];
#define DBG_ITER(i) do { if (lane == 0 && blockIdx.x < 2 && ptx::g_lsvs_hang[0] == 0) g_dbg_iter[blockIdx.x * 32 + warp] = (i); } while (0)
#else
#define DBG_ITER(i) do {} while (0)
#endif

template <int HD>
__global__ void __maxnreg__(Cfg<HD>::MAXNREG)
attention_fwd_tcgen05(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ O, int ldo, int Lq, int Lk,
                      float scale_log2e) {
  using C = Cfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Bars* bars = reinterpret_cast<Bars*>(smem + C::OFF_BAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NT = C::NT;
  constexpr int W_TMA = 4 * NT, W_MMA = 4 * NT + 1;  // warp ids of the two service warps
  const int q0 = blockIdx.x * (NT * QT);      // first query row (within the sequence) of this CTA
  const int head = blockIdx.y, batch = blockIdx.z;
  const int n_tiles = min(NT, (Lq - q0 + QT - 1) / QT);  // only tiles with at least one valid row
  const int n_kv = (Lk + C::BKV - 1) / C::BKV;
  const int q_row0 = batch * Lq + q0;          // global row of tile A's first query
  const int kv_row0 = batch * Lk;
  const int col0 = head * HD;

  if (warp == W_TMA && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(&bars->q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      ptx::mbar_init(&bars->k_full[i], 1); ptx::mbar_init(&bars->k_empty[i], 1);
      ptx::mbar_init(&bars->v_full[i], 1); ptx::mbar_init(&bars->v_empty[i], 1);
    }
    for (int t = 0; t < NT; ++t) {
      ptx::mbar_init(&bars->s_full[t], 1); ptx::mbar_init(&bars->s_free[t], 4);
      ptx::mbar_init(&bars->p_ready[t][0], 4); ptx::mbar_init(&bars->p_ready[t][1], 4);
      ptx::mbar_init(&bars->pv_done[t][0], 1); ptx::mbar_init(&bars->pv_done[t][1], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == W_MMA) { ptx::tmem_alloc(&bars->tmem_slot, C::TMEM_COLS); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == W_TMA) {
    // ============================================================ TMA producer
    reg_dec<HD>();
    if (lane == 0) {
      ptx::mbar_expect_tx(&bars->q_full, n_tiles * C::Q_TILE_BYTES);
      for (int t = 0; t < n_tiles; ++t)
        for (int kb = 0; kb < C::KB; ++kb)
          ptx::tma_load_2d(smem + C::OFF_Q + t * C::Q_TILE_BYTES + kb * (QT * 128), &tmQ, &bars->q_full, col0 + kb * 64,
                           q_row0 + t * QT);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < n_kv; ++i) {
      DBG_ITER(i);
      // K(i) first (needed early for S), then V(i)
      ptx::mbar_wait(&bars->k_empty[stage], phase ^ 1);
      if (lane == 0) {
        ptx::mbar_expect_tx(&bars->k_full[stage], C::K_TILE_BYTES);
        for (int kb = 0; kb < C::KB; ++kb)
          ptx::tma_load_2d(smem + C::OFF_K + stage * C::K_TILE_BYTES + kb * (C::BKV * 128), &tmK, &bars->k_full[stage],
                           col0 + kb * 64, kv_row0 + i * C::BKV);
      }
      ptx::mbar_wait(&bars->v_empty[stage], phase ^ 1);
      if (lane == 0) {
        ptx::mbar_expect_tx(&bars->v_full[stage], C::V_TILE_BYTES);
        for (int kb = 0; kb < C::KB; ++kb)
          ptx::tma_load_2d(smem + C::OFF_V + stage * C::V_TILE_BYTES + kb * (C::BKV * 128), &tmV, &bars->v_full[stage],
                           col0 + kb * 64, kv_row0 + i * C::BKV);
      }
      __syncwarp();
      if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == W_MMA) {
    // ============================================================ MMA issuer
    reg_dec<HD>();
    constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(QT, C::BKV, 0, 0);  // S = Q K^T : both K-major
    constexpr uint32_t idesc_o = ptx::umma_idesc_bf16(QT, HD, 0, 1);      // O = P V   : A K-major, B (V) MN-major
    const uint32_t sQ = ptx::smem_u32(smem + C::OFF_Q), sK = ptx::smem_u32(smem + C::OFF_K);
    const uint32_t sV = ptx::smem_u32(smem + C::OFF_V), sP = ptx::smem_u32(smem + C::OFF_P);

    auto issue_S = [&](int t, int stage) {
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        const uint64_t a = ptx::umma_desc_sw128(sQ + t * C::Q_TILE_BYTES + (k / 4) * (QT * 128) + (k % 4) * 32, 16, 1024);
        const uint64_t b = ptx::umma_desc_sw128(sK + stage * C::K_TILE_BYTES + (k / 4) * (C::BKV * 128) + (k % 4) * 32, 16, 1024);
        ptx::umma_bf16_ss(tmem + C::S_COL + t * C::BKV, a, b, idesc_s, k != 0);
      }
    };
    auto issue_PV = [&](int t, int stage, int pbuf, bool accumulate) {
#pragma unroll
      for (int k = 0; k < C::BKV / 16; ++k) {
        const uint64_t a = ptx::umma_desc_sw128(sP + (t * 2 + pbuf) * C::P_TILE_BYTES + (k / 4) * (QT * 128) + (k % 4) * 32, 16, 1024);
        // V block: rows = keys (K dim), 64-wide column blocks (N dim) C::BKV*128 bytes apart; 16 keys per step
        const uint64_t b = ptx::umma_desc_sw128(sV + stage * C::V_TILE_BYTES + k * (16 * 128), C::BKV * 128, 1024);
        ptx::umma_bf16_ss(tmem + C::O_COL + t * HD, a, b, idesc_o, (accumulate || k != 0) ? 1u : 0u);
      }
    };

    ptx::mbar_wait(&bars->q_full, 0);
    ptx::mbar_wait(&bars->k_full[0], 0);
    ptx::tc_fence_after();
    if (lane == 0) {
      for (int t = 0; t < n_tiles; ++t) { issue_S(t, 0); ptx::umma_commit(&bars->s_full[t]); }
      ptx::umma_commit(&bars->k_empty[0]);  // K(0) free once S_A(0), S_B(0) retire
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < n_kv; ++i) {
      DBG_ITER(i);
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == KV_STAGES) { nstage = 0; nphase ^= 1; }
      if (i + 1 < n_kv) {
        // S_t(i+1) as soon as the softmax warps hold S_t(i) in registers
        ptx::mbar_wait(&bars->k_full[nstage], nphase);
        for (int t = 0; t < n_tiles; ++t) {
          ptx::mbar_wait(&bars->s_free[t], i & 1);
          ptx::tc_fence_after();
          if (lane == 0) { issue_S(t, nstage); ptx::umma_commit(&bars->s_full[t]); }
          __syncwarp();
        }
        if (lane == 0) ptx::umma_commit(&bars->k_empty[nstage]);  // K(i+1) free once both S MMAs retire
        __syncwarp();
      }
      ptx::mbar_wait(&bars->v_full[stage], phase);
      for (int t = 0; t < n_tiles; ++t) {
        ptx::mbar_wait(&bars->p_ready[t][i & 1], (i >> 1) & 1);
        ptx::tc_fence_after();
        if (lane == 0) { issue_PV(t, stage, i & 1, i > 0); ptx::umma_commit(&bars->pv_done[t][i & 1]); }
        __syncwarp();
      }
      if (lane == 0) ptx::umma_commit(&bars->v_empty[stage]);
      __syncwarp();
      stage = nstage;
      phase = nphase;
    }
  } else if (warp > W_MMA) {
    reg_dec<HD>();  // idle warps of the service warpgroup (setmaxnreg is warpgroup-wide)
  } else {
    // ============================================================ softmax / correction / epilogue
    reg_inc<HD>();
    const int t = warp >> 2;                   // query tile of this softmax warpgroup
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;         // query row within the tile == TMEM lane
    if (t < n_tiles) {
      const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
      const uint32_t tS = tmem + lane_addr + C::S_COL + t * C::BKV;
      const uint32_t tO = tmem + lane_addr + C::O_COL + t * HD;
      uint8_t* sP0 = smem + C::OFF_P + (t * 2) * C::P_TILE_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
      // m_ref is the reference maximum used inside exp2; it trails the true running maximum by at most 8 (log2
      // units), so P <= 2^8 and O / l stay exact after the final division, while the TMEM rescale of O (and the wait
      // for the previous PV product it needs) only happens on the rare block where a row's maximum jumps by more.
      float m_ref = -INFINITY, l_run = 0.f;
      for (int i = 0; i < n_kv; ++i) {
        DBG_ITER(i);
        ptx::mbar_wait(&bars->s_full[t], i & 1);
        ptx::tc_fence_after();
        float s[C::BKV];
#pragma unroll
        for (int c = 0; c < C::BKV; c += 32) ptx::tmem_ld_32x32b_x32(tS + c, reinterpret_cast<uint32_t*>(s + c));
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars->s_free[t]);
        const int kv_valid = Lk - i * C::BKV;  // keys of this block inside the sequence
        if (kv_valid < C::BKV) {
#pragma unroll
          for (int c = 0; c < C::BKV; ++c) if (c >= kv_valid) s[c] = -INFINITY;
        }
        // the P buffer of this parity was last read by PV(i-2)
        if (i >= 2) ptx::mbar_wait(&bars->pv_done[t][i & 1], ((i - 2) >> 1) & 1);
        uint8_t* sP = sP0 + (i & 1) * C::P_TILE_BYTES;
        // exp2 (packed fp32x2 scale-and-shift), row sum, and P (bf16) straight into shared memory in the K-major
        // 128B-swizzled layout of the UMMA A operand; returns the row sum of this block
        auto emit_P = [&](float m_use) -> float {
          const float nmb = -m_use * scale_log2e;
          const unsigned long long sc2 = f2_pack(scale_log2e, scale_log2e), nmb2 = f2_pack(nmb, nmb);
          unsigned long long sum2[2] = {0ull, 0ull};
#pragma unroll
          for (int j = 0; j < C::BKV / 8; ++j) {
            float p[8];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              const unsigned long long x = f2_fma(f2_pack(s[8 * j + e], s[8 * j + e + 1]), sc2, nmb2);
              if (HD == 64 && e / 2 < LSVS_ATTN_POLY_PAIRS) {
                ex2_poly2(x, p[e], p[e + 1]);
              } else {
                p[e] = ex2(f2_lo(x));
                p[e + 1] = ex2(f2_hi(x));
              }
            }
            sum2[0] = f2_add(sum2[0], f2_add(f2_pack(p[0], p[1]), f2_pack(p[4], p[5])));
            sum2[1] = f2_add(sum2[1], f2_add(f2_pack(p[2], p[3]), f2_pack(p[6], p[7])));
            const int kb = j >> 3, chunk = j & 7;
            const uint4 v = make_uint4(ptx::pack_bf16(p[0], p[1]), ptx::pack_bf16(p[2], p[3]), ptx::pack_bf16(p[4], p[5]), ptx::pack_bf16(p[6], p[7]));
            *reinterpret_cast<uint4*>(sP + kb * (QT * 128) + ((chunk ^ (r & 7)) << 4)) = v;
          }
          const unsigned long long tot = f2_add(sum2[0], sum2[1]);
          return f2_lo(tot) + f2_hi(tot);
        };
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < C::BKV; c += 4) {
          mx[0] = fmaxf(mx[0], s[c]); mx[1] = fmaxf(mx[1], s[c + 1]); mx[2] = fmaxf(mx[2], s[c + 2]); mx[3] = fmaxf(mx[3], s[c + 3]);
        }
        const float m_blk = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        float blk_sum;
        if (i == 0) {
          m_ref = m_blk;
          blk_sum = emit_P(m_ref);
        } else {
          // optimistic: exponentiate against the trailing reference maximum (no dependence on this block's maximum, so
          // the MUFU work starts as soon as S is in registers); redo only if some row's maximum jumped by > 2^8
          blk_sum = emit_P(m_ref);
          const bool jump = (m_blk - m_ref) * scale_log2e > 8.0f;
          if (__any_sync(0xffffffffu, jump)) {
            // rescale O in TMEM: every earlier PV product must have retired (they complete in order)
            ptx::mbar_wait(&bars->pv_done[t][(i - 1) & 1], ((i - 1) >> 1) & 1);
            ptx::tc_fence_after();
            const float alpha = jump ? ex2((m_ref - m_blk) * scale_log2e) : 1.0f;
            if (jump) m_ref = m_blk;
            l_run *= alpha;
#pragma unroll
            for (int c = 0; c < HD; c += 16) {
              uint32_t o[16];
              ptx::tmem_ld_32x32b_x16(tO + c, o);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
              ptx::tmem_st_32x32b_x16(tO + c, o);
            }
            ptx::tmem_st_wait();
            blk_sum = emit_P(m_ref);
          }
        }
        l_run += blk_sum;
        // p_ready is double-buffered by iteration parity like the P tiles: P(i+2) is only written after PV(i) retired (wait
        // above), so the softmax warps can never lap the MMA warp on a barrier (parity waits only tell adjacent phases apart).
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars->p_ready[t][i & 1]);
      }
      // ---- epilogue: O / l -> bf16 -> global
      ptx::mbar_wait(&bars->pv_done[t][(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1);
      ptx::tc_fence_after();
      const float inv_l = 1.0f / l_run;
      const int q_local = q0 + t * QT + r;
      __nv_bfloat16* dst = O + (size_t)(batch * (size_t)Lq + q_local) * ldo + col0;
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t o[32];
        ptx::tmem_ld_32x32b_x32(tO + c, o);
        ptx::tmem_ld_wait();
        if (q_local < Lq) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = ptx::pack_bf16(__uint_as_float(o[8 * j + 0]) * inv_l, __uint_as_float(o[8 * j + 1]) * inv_l);
            u.y = ptx::pack_bf16(__uint_as_float(o[8 * j + 2]) * inv_l, __uint_as_float(o[8 * j + 3]) * inv_l);
            u.z = ptx::pack_bf16(__uint_as_float(o[8 * j + 4]) * inv_l, __uint_as_float(o[8 * j + 5]) * inv_l);
            u.w = ptx::pack_bf16(__uint_as_float(o[8 * j + 6]) * inv_l, __uint_as_float(o[8 * j + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + c + 8 * j) = u;
          }
        }
        __syncwarp();
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) ptx::tmem_dealloc(tmem, C::TMEM_COLS);
}

template <int HD>
int launch(const AttentionArgs& a, cudaStream_t st) {
  using C = Cfg<HD>;
  const size_t rows_q = (size_t)a.batches * a.Lq, rows_k = (size_t)a.batches * a.Lk;
  const CUtensorMap* tq = tmap_2d_bf16(a.q, (uint64_t)a.heads * HD, rows_q, (uint64_t)a.ldq * 2, 64, QT);
  const CUtensorMap* tk = tmap_2d_bf16(a.k, (uint64_t)a.heads * HD, rows_k, (uint64_t)a.ldk * 2, 64, C::BKV);
  const CUtensorMap* tv = tmap_2d_bf16(a.v, (uint64_t)a.heads * HD, rows_k, (uint64_t)a.ldv * 2, 64, C::BKV);
  if (!tq || !tk || !tv) return LSVS_ECUDA;
  auto kern = attention_fwd_tcgen05<HD>;
  static bool configured = false;
  if (!configured) {
    LSVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    configured = true;
  }
  dim3 grid((a.Lq + C::NT * QT - 1) / (C::NT * QT), a.heads, a.batches);
  const float scale_log2e = a.scale * 1.4426950408889634f;
  kern<<<grid, C::NTHREADS, C::SMEM, st>>>(*tq, *tk, *tv, reinterpret_cast<__nv_bfloat16*>(a.o), a.ldo, a.Lq, a.Lk, scale_log2e);
  LSVS_LAUNCH_CHECK();
  return LSVS_OK;
}

}  // namespace

int attention_fwd(const AttentionArgs& a, cudaStream_t st) {
  LSVS_CHECK_ARG(a.q && a.k && a.v && a.o, "attention: null pointer");
  LSVS_CHECK_ARG(a.batches > 0 && a.heads > 0 && a.Lq > 0 && a.Lk > 0, "attention: empty shape");
  LSVS_CHECK_ARG(a.batches <= 65535 && a.heads <= 65535, "attention: batch/head count exceeds the grid limit");
  LSVS_CHECK_ARG(a.head_dim == 64 || a.head_dim == 128, "attention: head_dim %d unsupported (64 or 128)", a.head_dim);
  const int D = a.heads * a.head_dim;
  LSVS_CHECK_ARG(a.ldq >= D && a.ldk >= D && a.ldv >= D && a.ldo >= D, "attention: leading dimension smaller than heads*head_dim");
  LSVS_CHECK_ARG(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0, "attention: leading dimensions must be multiples of 8");
  ProfScope prof(a.Lk >= 2048 ? PROF_ATTENTION_GLOBAL : PROF_ATTENTION, st, 4.0 * a.batches * (double)a.heads * a.Lq * (double)a.Lk * a.head_dim, 0);
  return a.head_dim == 64 ? launch<64>(a, st) : launch<128>(a, st);
}

}  // namespace lsvs

#ifdef LSVS_DEBUG_HANG
extern "C" int lsvs_debug_hang_read(int* out257, int* bar_base_offset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out257, ptx::g_lsvs_hang, 257 * sizeof(int));
  cudaMemcpyFromSymbol(out257 + 257, lsvs::g_dbg_iter, 64 * sizeof(int));
  *bar_base_offset = lsvs::Cfg<64>::OFF_BAR;
  static int zero[257] = {0};
  cudaMemcpyToSymbol(ptx::g_lsvs_hang, zero, sizeof(zero));
  return 0;
}
#endif
