// Fused flash-style attention  O = softmax(Q K^T * scale) V  on tcgen05 / TMEM / TMA (no mask; ragged tails).
//
// One CTA owns one 128-row query tile of one (batch, head), two CTAs per SM (NT = 1, the default at every sequence length since
// round 2: two independent CTAs fill each other's waits); the NT = 2 variant — two tiles (A, B) per CTA sharing every K/V block,
// one CTA per SM — is kept behind LSVS_ATTN_NT1_MAX_LK for comparison.  Everything the tensor core touches besides Q/K/V lives in
// tensor memory: S_t (fp32), O_t (fp32) and P_t (bf16 pairs, the A operand of the PV product).  Warp roles, NT = 2:
//   warps 0-3   softmax for tile A, warps 4-7 for tile B (registers moved to them with setmaxnreg): one thread per query row
//               (TMEM lane), so row sums need no shuffles.  exp2 with the softmax scale folded in, against a trailing
//               reference maximum (first block's maximum, raised only when a row sum proves an element 2^8 above it);
//               O stays in TMEM and is rescaled in place (tcgen05.ld/st) only then.  P goes back with tcgen05.st.
//   warp 8      TMA producer: Q tiles once, then the K blocks through a 4-stage ring
//   warp 9      tcgen05.mma issuer: S_t = Q_t K^T (SS form), O_t += P_t V (TS form: A from tensor memory, V MN-major)
//   warp 10     TMA producer of the V blocks (own ring: a V slot frees late, after PV, and must not delay the next K request)
// The two tiles ping-pong on the tensor core: while the softmax warps of one tile work, the MMAs of the other
// run.  S_t(i+1) is issued as soon as the softmax warps have pulled S_t(i) into registers (s_free barrier).
// Launched with programmatic dependent launch: the prologue overlaps the tail of the previous kernel.
//
// Replaces F.scaled_dot_product_attention behind UPSTREAM Attention (Aggregator frame/global/DINO blocks,
// alignment-head frame blocks alignment_head.py:363, camera-head trunk); q_norm/k_norm/RoPE are already applied
// by the QKV GEMM epilogue (csrc/gemm.cu).
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "attention.h"
#include "host_common.h"
#include "ptx.cuh"
#include "tensormap.h"

#ifndef LSVS_ATTN_POLY_DEFAULT
#define LSVS_ATTN_POLY_DEFAULT 0
#endif

namespace lsvs {
namespace {

constexpr int QT = 128;            // query rows per tile (UMMA M)
// threads = (NT + 1) warpgroups: one softmax warpgroup per query tile, the last one holds the K producer, the MMA warp, the V producer and an idle warp
constexpr int KV_STAGES = 4;


// POLY_: of every 4 pairs of scores, POLY_ pairs are exponentiated on the FMA / ALU pipes (Cody-Waite range reduction + a degree-3
// polynomial in packed fp32x2 arithmetic) instead of MUFU.EX2: at head dim 64 the softmax needs 2 x 128 x 128 exponentials per
// 128-key block and SM at 16 per clock = 2048 cycles, twice the tensor-core time of the block, so the XU pipe is the bound.
template <int HD, int NT_, int POLY_ = 0>
struct Cfg {
  static constexpr int POLY = POLY_;
  // NT_ = 2: two 128-row query tiles per CTA share every K/V block, one CTA per SM (long sequences).
  // NT_ = 1: one query tile, half the tensor memory and a 2-stage K/V ring so that TWO CTAs fit on an SM: for short
  //          sequences (frame attention, a few key blocks per CTA) the prologue / epilogue of one CTA hides behind the other.
  // 128-key blocks at head dim 64, 64-key blocks at head dim 128.
  // (Measured alternative at head dim 64: four tiles x 64-key blocks, i.e. 16 softmax warps: 455 vs 666 TFLOP/s —
  //  the per-block barrier / fence overhead doubles per key and outweighs the extra latency hiding.)
  static constexpr int NT = NT_;                            // query tiles per CTA
  static constexpr int KVS = (NT_ == 2) ? KV_STAGES : 2;    // K / V ring depth
  static constexpr int BKV = (HD == 64) ? 128 : 64;         // keys per block (UMMA N of S, K of PV)
  // (Measured alternatives, DESIGN.md 5: two threads per softmax row / 16 softmax warps, 701 vs 730 TFLOP/s; a share of the
  //  exponentials as a degree-3 polynomial on the FMA pipe, 708 vs 730; PRMT instead of F2FP packing, no change.)
  static constexpr int COLS = BKV;                          // S columns per softmax thread (one thread per query row)
  static constexpr int OCOLS = HD;                          // O columns per softmax thread (rescale / epilogue)
  static constexpr int NWG = NT;                            // softmax warpgroups
  static constexpr int NTHREADS = (NWG + 1) * 128;
  static constexpr int MAXNREG = NT_ == 2 ? 168 : 128;      // launch-time registers / thread (register file / resident threads)
  static constexpr int REG_SOFTMAX = 200;                   // after setmaxnreg: NWG*128*REG_SOFTMAX + 128*REG_SERVICE <= NTHREADS*MAXNREG
  static constexpr int REG_SERVICE = NT_ == 2 ? 96 : 56;
  static constexpr int KB = HD / 64;                        // 64-element (128 B) column blocks of the head dim
  static constexpr int Q_TILE_BYTES = QT * HD * 2;
  static constexpr int K_TILE_BYTES = BKV * HD * 2;
  static constexpr int V_TILE_BYTES = BKV * HD * 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + NT * Q_TILE_BYTES;
  static constexpr int OFF_V = OFF_K + KVS * K_TILE_BYTES;
  static constexpr int OFF_BAR = OFF_V + KVS * V_TILE_BYTES;
  static constexpr int SMEM = OFF_BAR + 512;                      // the dynamic shared window is 1024-aligned (no static smem)
  static constexpr int S_COL = 0;                            // TMEM columns: S_A, S_B, O_A, O_B, P_A, P_B
  static constexpr int O_COL = NT * BKV;
  static constexpr int P_COL = O_COL + NT * HD;              // P: bf16 pairs, BKV / 2 columns per tile
  static constexpr int PCOLS = BKV / 2;
  static constexpr int TMEM_COLS = (NT_ == 2) ? 512 : 256;
  static_assert(P_COL + NT * PCOLS <= TMEM_COLS, "TMEM budget");
  static_assert(SMEM <= (NT_ == 2 ? 232448 : 232448 / 2 - 1024), "shared memory budget");
  static_assert(NWG * 128 * REG_SOFTMAX + 128 * REG_SERVICE <= NTHREADS * MAXNREG, "register pool");
};

struct Bars {
  uint64_t q_full, q_empty;
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

template <bool V> struct BoolTag { static constexpr bool value = V; };

// register re-balancing between the service warpgroup and the softmax warpgroups (setmaxnreg, warpgroup-wide)
template <class C> __device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REG_SERVICE));
}
template <class C> __device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REG_SOFTMAX));
}

#ifdef LSVS_ATTN_PHASES
__device__ unsigned long long g_attn_phase[8 * 8];  // [warp][phase] cycle sums for CTA (0,0,0)
#define PH_DECL unsigned ph_t = clock(); unsigned long long ph_acc[6] = {0, 0, 0, 0, 0, 0}
#define PH(k) do { const unsigned now_ = clock(); ph_acc[k] += now_ - ph_t; ph_t = now_; } while (0)
#define PH_FLUSH() do { if (lane == 0 && blockIdx.x == 0) for (int k_ = 0; k_ < 6; ++k_) g_attn_phase[warp * 8 + k_] = ph_acc[k_]; } while (0)
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ unsigned long long g_attn_tl[8 * 8];   // timeline of CTA 0, softmax warp 0: [item ordinal][0: start, 1..6: after block 0..5, 7: rows stored]
#define TL(n_, k_) do { if (lane == 0 && warp == 0 && blockIdx.x == 0 && (n_) < 8 && (k_) < 8) g_attn_tl[(n_) * 8 + (k_)] = gtime(); } while (0)
#define MS_ENTRY const unsigned long long ms_t0 = gtime()
#define MS(k) do { if (threadIdx.x == 0 && blockIdx.x == gridDim.x - 1) g_attn_phase[48 + (k)] = gtime() - ms_t0; } while (0)
#else
#define TL(n_, k_) do {} while (0)
#define MS_ENTRY do {} while (0)
#define MS(k) do {} while (0)
#define PH_DECL do {} while (0)
#define PH(k) do {} while (0)
#define PH_FLUSH() do {} while (0)
#endif

#ifdef LSVS_DEBUG_HANG
__device__ int g_dbg_iter[64];
#define DBG_ITER(i) do { if (lane == 0 && blockIdx.x < 2 && ptx::g_lsvs_hang[0] == 0) g_dbg_iter[blockIdx.x * 32 + warp] = (i); } while (0)
#else
#define DBG_ITER(i) do {} while (0)
#endif

// 2^x for x <= ~8 without the XU pipe: x = n + f with n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial
// (max relative error 7.5e-5, far below the bf16 rounding of P); 2^n by adding n to the exponent field.  Inputs below -126
// are clamped (their weight is < 2^-126).  Packed fp32x2: two scores per FADD2 / FFMA2.
__device__ __forceinline__ void exp2_poly2(unsigned long long x2, float& p0, float& p1) {
  const float x0 = fmaxf(f2_lo(x2), -126.0f), x1 = fmaxf(f2_hi(x2), -126.0f);
  const unsigned long long xc = f2_pack(x0, x1);
  const unsigned long long magic = f2_pack(12582912.0f, 12582912.0f), nmagic = f2_pack(-12582912.0f, -12582912.0f);
  const unsigned long long r = f2_add(xc, magic);                       // integer part in the low mantissa bits
  const unsigned long long n = f2_add(r, nmagic);
  const unsigned long long f = f2_fma(n, f2_pack(-1.0f, -1.0f), xc);    // x - n
  unsigned long long p = f2_fma(f2_pack(0.05517167f, 0.05517167f), f, f2_pack(0.24261114f, 0.24261114f));
  p = f2_fma(p, f, f2_pack(0.69326097f, 0.69326097f));
  p = f2_fma(p, f, f2_pack(0.99992806f, 0.99992806f));
  p0 = __uint_as_float(__float_as_uint(f2_lo(p)) + (__float_as_uint(f2_lo(r)) << 23));
  p1 = __uint_as_float(__float_as_uint(f2_hi(p)) + (__float_as_uint(f2_hi(r)) << 23));
}

// How the grid maps to work (see the decode at the top of the kernel).
struct WorkSplit {
  int n_qt, heads;      // query-tile groups per sequence, heads: unit = (batch * heads + head) * n_qt + tile group
  int n_items;          // work items in all: n_full whole units + (units - n_full) * parts key ranges
  int n_full;           // units [0, n_full) run their whole key range in one CTA
  int parts;            // every later unit is cut into `parts` key ranges (CTAs n_full + (unit - n_full) * parts + part)
  float* partial;       // [(unit - n_full) * parts + part][NT * QT rows][HD + 2]: unnormalised O, reference maximum, row sum
};

// Merges the key-range partials of the split units:  O = sum_p w_p O_p / sum_p w_p l_p,  w_p = 2^((m_p - max_p m_p) * scale).
// One warp per query row, HD / 32 columns per lane.
template <int HD, int ROWS>
__global__ void attention_combine_kernel(const float* __restrict__ partial, __nv_bfloat16* __restrict__ O, int ldo, int Lq,
                                         float scale_log2e, int n_qt, int heads, int n_full, int parts) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int su = blockIdx.x / (ROWS / 4);                     // split unit
  const int row = (blockIdx.x % (ROWS / 4)) * 4 + warp;       // row within the unit
  const int unit = n_full + su;
  const int q_local = (unit % n_qt) * ROWS + row;
  if (q_local >= Lq) return;
  const int head = (unit / n_qt) % heads, batch = unit / (n_qt * heads);
  const float* base = partial + ((size_t)su * parts * ROWS + row) * (HD + 2);
  const size_t pstride = (size_t)ROWS * (HD + 2);
  float m = -INFINITY;
  for (int p = 0; p < parts; ++p) m = fmaxf(m, base[p * pstride + HD]);
  constexpr int CPL = HD / 32;
  float acc[CPL] = {}, l = 0.f;
  for (int p = 0; p < parts; ++p) {
    const float* src = base + p * pstride;
    const float w = exp2f((src[HD] - m) * scale_log2e);
    l += w * src[HD + 1];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] += w * src[lane * CPL + c];
  }
  const float inv = 1.0f / l;
  __nv_bfloat16* dst = O + (size_t)(batch * (size_t)Lq + q_local) * ldo + head * HD + lane * CPL;
  if constexpr (CPL == 2) {
    *reinterpret_cast<uint32_t*>(dst) = ptx::pack_bf16(acc[0] * inv, acc[1] * inv);
  } else {
    uint2 u; u.x = ptx::pack_bf16(acc[0] * inv, acc[1] * inv); u.y = ptx::pack_bf16(acc[2] * inv, acc[3] * inv);
    *reinterpret_cast<uint2*>(dst) = u;
  }
}

template <int HD, int NT_, int POLY_, bool SPLIT>
__global__ void __maxnreg__((Cfg<HD, NT_, POLY_>::MAXNREG))
attention_fwd_tcgen05(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ O, int ldo, int Lq, int Lk,
                      float scale_log2e, WorkSplit ws) {
  using C = Cfg<HD, NT_, POLY_>;
  MS_ENTRY;
  extern __shared__ __align__(1024) uint8_t smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem + C::OFF_BAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NT = C::NT;
  constexpr int W_TMA = 4 * C::NWG, W_MMA = 4 * C::NWG + 1, W_TMA_V = 4 * C::NWG + 2;  // warp ids of the two service warps
  // Work items (WorkSplit): item < n_full is a whole unit = (batch, head, NT query tiles) over every key block; the units of the
  // last, partial round of CTAs are cut into `parts` key ranges each, written as unnormalised partials and merged by
  // attention_combine_kernel — the tail of the grid then lasts 1/parts of a unit instead of a whole one.
  // PERSISTENT CTAs (NT = 1): CTA c runs items c, c + gridDim.x, ...  Barriers, the K / V rings and tensor memory are set up once;
  // every role walks the concatenated block sequence of its items, so the loads of item n+1 (Q once the last S of item n has
  // retired, K / V through their rings) and its first S MMA are in flight while the softmax warps finish item n and store its
  // rows — the ~2 us of set-up, first loads and tear-down that a 412-token CTA spent per 8 us of life are paid once.
  const int n_kv_seq = (Lk + C::BKV - 1) / C::BKV;
  struct Item { int q0, n_tiles, n_kv, Lkp, q_row0, kv_row0, col0, batch; bool partial; long long prow0; };
  auto decode = [&](int item) -> Item {
    int unit = item, part = 0, n_parts = 1;
    if (SPLIT && unit >= ws.n_full) { const int idx = unit - ws.n_full; unit = ws.n_full + idx / ws.parts; part = idx % ws.parts; n_parts = ws.parts; }
    const int qt = unit % ws.n_qt, head = (unit / ws.n_qt) % ws.heads;
    Item it;
    it.batch = unit / (ws.n_qt * ws.heads);
    it.q0 = qt * (NT * QT);                                   // first query row (within the sequence)
    it.n_tiles = min(NT, (Lq - it.q0 + QT - 1) / QT);         // only tiles with at least one valid row
    const int kb0 = (int)((long long)part * n_kv_seq / n_parts);   // key blocks [kb0, kb0 + n_kv) of the sequence
    it.n_kv = (int)((long long)(part + 1) * n_kv_seq / n_parts) - kb0;
    it.Lkp = min(Lk - kb0 * C::BKV, it.n_kv * C::BKV);        // valid keys of the range (ragged only at the sequence end)
    it.q_row0 = it.batch * Lq + it.q0;                        // global row of tile A's first query
    it.kv_row0 = it.batch * Lk + kb0 * C::BKV;
    it.col0 = head * HD;
    it.partial = SPLIT && n_parts > 1;
    it.prow0 = it.partial ? ((long long)(unit - ws.n_full) * ws.parts + part) * (NT * QT) : 0;
    return it;
  };
  // the last key block of a ragged sequence (412 = 3 x 128 + 28 keys) only spans n_last keys (a multiple of 32): a narrower
  // S = Q K^T (UMMA N = n_last) and a shorter P V reduction, and the softmax warps touch n_last columns instead of BKV
  auto last_cols = [](const Item& it) { return min(C::BKV, ((it.Lkp - (it.n_kv - 1) * C::BKV) + 31) & ~31); };
  const int item0 = blockIdx.x, item_step = gridDim.x;

  if (warp == W_TMA && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(&bars->q_full, 1); ptx::mbar_init(&bars->q_empty, 1);
    for (int i = 0; i < C::KVS; ++i) {
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
  pdl_launch_dependents();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;
  pdl_wait();  // the prologue above overlapped the previous kernel's tail (programmatic dependent launch)
  MS(0);  // setup done (barriers, TMEM allocation, CTA sync)

  if (warp == W_TMA) {
    // ============================================================ TMA producer (Q tiles, K blocks)
    reg_dec<C>();
    // K blocks only: V has its own producer warp so that a V slot that frees late (after PV(i-1)) never holds back the
    // request for K(i+2) — each ring then runs a full two key blocks ahead of its consumer
    int stage = 0;
    uint32_t phase = 0;
    int n = 0;
    for (int item = item0; item < ws.n_items; item += item_step, ++n) {
      const Item it = decode(item);
      if (n > 0) ptx::mbar_wait(&bars->q_empty, (n - 1) & 1);   // every S MMA of the previous item has read its Q tiles
      if (lane == 0) {
        ptx::mbar_expect_tx(&bars->q_full, it.n_tiles * C::Q_TILE_BYTES);
        for (int t = 0; t < it.n_tiles; ++t)
          for (int kb = 0; kb < C::KB; ++kb)
            ptx::tma_load_2d(smem + C::OFF_Q + t * C::Q_TILE_BYTES + kb * (QT * 128), &tmQ, &bars->q_full, it.col0 + kb * 64,
                             it.q_row0 + t * QT);
      }
      for (int i = 0; i < it.n_kv; ++i) {
        DBG_ITER(i);
        ptx::mbar_wait(&bars->k_empty[stage], phase ^ 1);
        if (lane == 0) {
          ptx::mbar_expect_tx(&bars->k_full[stage], C::K_TILE_BYTES);
          for (int kb = 0; kb < C::KB; ++kb)
            ptx::tma_load_2d(smem + C::OFF_K + stage * C::K_TILE_BYTES + kb * (C::BKV * 128), &tmK, &bars->k_full[stage],
                             it.col0 + kb * 64, it.kv_row0 + i * C::BKV);
        }
        __syncwarp();
        if (++stage == C::KVS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == W_TMA_V) {
    // ============================================================ TMA producer (V blocks)
    reg_dec<C>();
    int stage = 0;
    uint32_t phase = 0;
    for (int item = item0; item < ws.n_items; item += item_step) {
      const Item it = decode(item);
      for (int i = 0; i < it.n_kv; ++i) {
        ptx::mbar_wait(&bars->v_empty[stage], phase ^ 1);
        if (lane == 0) {
          ptx::mbar_expect_tx(&bars->v_full[stage], C::V_TILE_BYTES);
          for (int kb = 0; kb < C::KB; ++kb)
            ptx::tma_load_2d(smem + C::OFF_V + stage * C::V_TILE_BYTES + kb * (C::BKV * 128), &tmV, &bars->v_full[stage],
                             it.col0 + kb * 64, it.kv_row0 + i * C::BKV);
        }
        __syncwarp();
        if (++stage == C::KVS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == W_MMA) {
    // ============================================================ MMA issuer
    reg_dec<C>();
    constexpr uint32_t idesc_o = ptx::umma_idesc_bf16(QT, HD, 0, 1);      // O = P V   : A K-major, B (V) MN-major
    const uint32_t sQ = ptx::smem_u32(smem + C::OFF_Q), sK = ptx::smem_u32(smem + C::OFF_K);
    const uint32_t sV = ptx::smem_u32(smem + C::OFF_V);
    auto issue_S = [&](int t, int stage, int n_cols) {                    // S = Q K^T : both K-major, UMMA N = n_cols
      const uint32_t id = ptx::umma_idesc_bf16(QT, n_cols, 0, 0);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        const uint64_t a = ptx::umma_desc_sw128(sQ + t * C::Q_TILE_BYTES + (k / 4) * (QT * 128) + (k % 4) * 32, 16, 1024);
        const uint64_t b = ptx::umma_desc_sw128(sK + stage * C::K_TILE_BYTES + (k / 4) * (C::BKV * 128) + (k % 4) * 32, 16, 1024);
        ptx::umma_bf16_ss(tmem + C::S_COL + t * C::BKV, a, b, id, k != 0);
      }
    };
    auto issue_PV = [&](int t, int stage, bool accumulate, int n_keys) {
#pragma unroll
      for (int k = 0; k < C::BKV / 16; ++k) {
        if (16 * k >= n_keys) break;
        // V block: rows = keys (K dim), 64-wide column blocks (N dim) C::BKV*128 bytes apart; 16 keys per step
        const uint64_t b = ptx::umma_desc_sw128(sV + stage * C::V_TILE_BYTES + k * (16 * 128), C::BKV * 128, 1024);
        // A = P from tensor memory: lane = query row, 8 columns (16 packed bf16) per step
        ptx::umma_bf16_ts(tmem + C::O_COL + t * HD, tmem + C::P_COL + t * C::PCOLS + k * 8, b, idesc_o, (accumulate || k != 0) ? 1u : 0u);
      }
    };

    if (item0 < ws.n_items) {
      Item cur = decode(item0);
      int cur_last = last_cols(cur);
      ptx::mbar_wait(&bars->q_full, 0);
      ptx::mbar_wait(&bars->k_full[0], 0);
      ptx::tc_fence_after();
      if (lane == 0) {
        for (int t = 0; t < cur.n_tiles; ++t) { issue_S(t, 0, cur.n_kv == 1 ? cur_last : C::BKV); ptx::umma_commit(&bars->s_full[t]); }
        ptx::umma_commit(&bars->k_empty[0]);  // K(0) free once S_A(0), S_B(0) retire
        if (cur.n_kv == 1) ptx::umma_commit(&bars->q_empty);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0, g = 0;   // g: blocks issued so far by this CTA (barrier parities run across items)
      int n = 0;
      for (int item = item0; item < ws.n_items; item += item_step, ++n) {
        const bool more_items = item + item_step < ws.n_items;
        Item nxt = cur;
        int nxt_last = cur_last;
        if (more_items) { nxt = decode(item + item_step); nxt_last = last_cols(nxt); }
        for (int i = 0; i < cur.n_kv; ++i, ++g) {
          DBG_ITER(i);
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == C::KVS) { nstage = 0; nphase ^= 1; }
          const bool in_item = i + 1 < cur.n_kv;
          if (in_item || more_items) {
            // the next block's S_t as soon as the softmax warps hold this block's S_t in registers; across an item boundary the
            // next item's Q tiles must have landed as well
            const Item& ni = in_item ? cur : nxt;
            const int bi = in_item ? i + 1 : 0;                       // its index within its item
            const bool ni_last = bi + 1 == ni.n_kv;
            if (!in_item) ptx::mbar_wait(&bars->q_full, (n + 1) & 1);
            ptx::mbar_wait(&bars->k_full[nstage], nphase);
            for (int t = 0; t < ni.n_tiles; ++t) {
              ptx::mbar_wait(&bars->s_free[t], g & 1);
              ptx::tc_fence_after();
              if (lane == 0) { issue_S(t, nstage, ni_last ? (in_item ? cur_last : nxt_last) : C::BKV); ptx::umma_commit(&bars->s_full[t]); }
              __syncwarp();
            }
            if (lane == 0) {
              ptx::umma_commit(&bars->k_empty[nstage]);               // that K block is free once its S MMAs retire
              if (ni_last) ptx::umma_commit(&bars->q_empty);          // and so are the item's Q tiles after its last S
            }
            __syncwarp();
          }
          ptx::mbar_wait(&bars->v_full[stage], phase);
          for (int t = 0; t < cur.n_tiles; ++t) {
            ptx::mbar_wait(&bars->p_ready[t][g & 1], (g >> 1) & 1);
            ptx::tc_fence_after();
            if (lane == 0) { issue_PV(t, stage, i > 0, i + 1 == cur.n_kv ? cur_last : C::BKV); ptx::umma_commit(&bars->pv_done[t][g & 1]); }
            __syncwarp();
          }
          if (lane == 0) ptx::umma_commit(&bars->v_empty[stage]);
          __syncwarp();
          stage = nstage;
          phase = nphase;
        }
        cur = nxt;
        cur_last = nxt_last;
      }
    }
  } else if (warp > W_TMA_V) {
    reg_dec<C>();  // idle warp of the service warpgroup (setmaxnreg is warpgroup-wide)
  } else {
    // ============================================================ softmax / correction / epilogue
    reg_inc<C>();
    constexpr int COLS = C::COLS, OCOLS = C::OCOLS;
    const int t = warp >> 2;                   // query tile of this softmax warpgroup
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;         // query row within the tile == TMEM lane
    uint32_t g0 = 0;                           // blocks of the items before the current one (barrier parities run across items)
    int n_ord = 0;
    for (int item = item0; item < ws.n_items; item += item_step, ++n_ord) {
    const Item it = decode(item);
    TL(n_ord, 0);
    const int n_kv = it.n_kv, Lkp = it.Lkp, q0 = it.q0;
    if (t < it.n_tiles) {
      const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
      const uint32_t tS = tmem + lane_addr + C::S_COL + t * C::BKV;
      const uint32_t tO = tmem + lane_addr + C::O_COL + t * HD;
      const uint32_t tP = tmem + lane_addr + C::P_COL + t * C::PCOLS;
      // m_ref is the reference maximum used inside exp2: the maximum of the first key block, raised only when a later
      // element exceeds it by more than 8 (log2 units), so P <= 2^8 and O / l stay exact after the final division,
      // while the TMEM rescale of O only happens on the rare block where a row's maximum jumps by more.
      float m_ref = -INFINITY, l_run = 0.f;
      // a warp whose 32 query rows all lie past Lq (ragged last tile: 412 = 3 x 128 + 28 rows; 32-row temporal attention) only
      // keeps the barrier protocol going: its rows of S / P / O are never stored, and rows of an MMA are independent
      const bool warp_active = q0 + t * QT + quarter * 32 < Lq;
      const int n_last = last_cols(it);
      PH_DECL;
      for (int i = 0; i < n_kv; ++i) {
        DBG_ITER(i);
        const uint32_t g = g0 + i;                            // block index within this CTA's whole run
        const int n_cols = (i + 1 == n_kv) ? n_last : COLS;   // S columns of this key block (warp-uniform)
        ptx::mbar_wait(&bars->s_full[t], g & 1);
        ptx::tc_fence_after();
        PH(0);
        if (i == 0) MS(1);  // first S tile ready (Q, K(0) landed, first MMA retired)
        float s[COLS];
        const int kv_valid = Lkp - i * C::BKV;  // keys of this block inside the sequence
        // steady-state block (not the first, all columns valid): only the first 32 columns of S are waited for here; the other
        // tensor-memory loads stay in flight under the first exponentials, and so does the wait for P V(i-1) (emit_full's hooks)
        const bool split = warp_active && i >= 1 && kv_valid >= COLS;
        if (split) {
          ptx::tmem_ld_32x32b_x32(tS, reinterpret_cast<uint32_t*>(s));
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 32; c < COLS; c += 32) ptx::tmem_ld_32x32b_x32(tS + c, reinterpret_cast<uint32_t*>(s + c));
        } else {
          if (warp_active) {
            if (n_cols == COLS) {
#pragma unroll
              for (int c = 0; c < COLS; c += 32) ptx::tmem_ld_32x32b_x32(tS + c, reinterpret_cast<uint32_t*>(s + c));
            } else {
#pragma unroll
              for (int c = 0; c < COLS; c += 32)
                if (c < n_cols) ptx::tmem_ld_32x32b_x32(tS + c, reinterpret_cast<uint32_t*>(s + c));
            }
            ptx::tmem_ld_wait();
          }
          PH(1);
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bars->s_free[t]);
          if (warp_active && kv_valid < COLS) {
#pragma unroll
            for (int c = 0; c < COLS; ++c) if (c >= kv_valid) s[c] = -INFINITY;
          }
        }
        auto rest_of_S = [&]() {   // before the first use of columns >= 32
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 32; c < COLS; ++c) asm volatile("" : "+f"(s[c]));   // no use of these registers may move above the wait
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bars->s_free[t]);
        };
        auto p_free = [&]() {      // before the first tcgen05.st of P: P V(i-1) has consumed the previous P (and O is quiescent)
          ptx::mbar_wait(&bars->pv_done[t][(g - 1) & 1], ((g - 1) >> 1) & 1);
          ptx::tc_fence_after();
        };
        auto no_hook = []() {};
        // exp2 (packed fp32x2 scale-and-shift), row sum, and P packed to bf16 pairs and stored to its tensor-memory tile 32
        // keys at a time (measured alternative: keep all of P in registers and store after the wait for PV(i-1): 714 vs 730)
        // FULL (compile-time): the block spans all COLS columns — every block but a ragged sequence's last one; the narrow variant
        // carries the run-time column bound.  The row sums and the bf16 packing of a group of 8 exponentials are issued one group
        // LATE, after the next group's MUFU.EX2 instructions: nothing in the stream then waits on the MUFU result latency (ncu
        // source view, profiles/r2_ncu_attention_poly.md: the FADD2 / F2FP right behind their MUFUs cost 5.3 / 2.9 cycles each).
        auto group_exp = [&](int j, float* p, unsigned long long sc2, unsigned long long nmb2) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const unsigned long long x = f2_fma(f2_pack(s[8 * j + e], s[8 * j + e + 1]), sc2, nmb2);
            if (e / 2 < C::POLY) {   // this pair on the FMA / ALU pipes
              exp2_poly2(x, p[e], p[e + 1]);
            } else {
              p[e] = ex2(f2_lo(x));
              p[e + 1] = ex2(f2_hi(x));
            }
          }
        };
        auto group_out = [&](int j, const float* p, unsigned long long* sum2, uint32_t* pk) {
          sum2[0] = f2_add(sum2[0], f2_add(f2_pack(p[0], p[1]), f2_pack(p[4], p[5])));
          sum2[1] = f2_add(sum2[1], f2_add(f2_pack(p[2], p[3]), f2_pack(p[6], p[7])));
#pragma unroll
          for (int e = 0; e < 4; ++e) pk[4 * (j & 3) + e] = ptx::pack_bf16(p[2 * e], p[2 * e + 1]);
          if ((j & 3) == 3) ptx::tmem_st_32x32b_x16(tP + (j - 3) * 4, pk);   // 32 keys = 16 packed columns
        };
        // full block (every block but a ragged sequence's last one): all COLS columns, no run-time bounds
        auto emit_full = [&](float m_use, auto&& need_rest, auto&& first_store) -> float {
          const float nmb = -m_use * scale_log2e;
          const unsigned long long sc2 = f2_pack(scale_log2e, scale_log2e), nmb2 = f2_pack(nmb, nmb);
          unsigned long long sum2[2] = {0ull, 0ull};
          uint32_t pk[16];
          float pa[8], pb[8];
          group_exp(0, pa, sc2, nmb2);
#pragma unroll
          for (int j = 1; j < COLS / 8; j += 2) {
            group_exp(j, pb, sc2, nmb2);
            group_out(j - 1, pa, sum2, pk);
            if (j == 3) need_rest();          // group 4 is the first one in columns >= 32
            if (j + 1 < COLS / 8) group_exp(j + 1, pa, sc2, nmb2);
            if (j == 3) first_store();        // group_out(3) issues the first tcgen05.st of P
            group_out(j, pb, sum2, pk);
          }
          const unsigned long long tot = f2_add(sum2[0], sum2[1]);
          return f2_lo(tot) + f2_hi(tot);
        };
        // ragged last block: only the first n_cols columns
        auto emit_part = [&](float m_use) -> float {
          const float nmb = -m_use * scale_log2e;
          const unsigned long long sc2 = f2_pack(scale_log2e, scale_log2e), nmb2 = f2_pack(nmb, nmb);
          unsigned long long sum2[2] = {0ull, 0ull};
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < COLS / 8; ++j) {
            if (8 * j >= n_cols) break;
            float p[8];
            group_exp(j, p, sc2, nmb2);
            group_out(j, p, sum2, pk);
          }
          const unsigned long long tot = f2_add(sum2[0], sum2[1]);
          return f2_lo(tot) + f2_hi(tot);
        };
        auto block_max = [&](auto full_tag) -> float {
          constexpr bool FULL = decltype(full_tag)::value;
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c = 0; c < COLS; c += 4) {
            if (!FULL && c >= n_cols) break;
            mx[0] = fmaxf(mx[0], s[c]); mx[1] = fmaxf(mx[1], s[c + 1]); mx[2] = fmaxf(mx[2], s[c + 2]); mx[3] = fmaxf(mx[3], s[c + 3]);
          }
          return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        };
        using FullT = BoolTag<true>;
        using PartT = BoolTag<false>;
        const bool full_block = n_cols == COLS;
        auto emit = [&](float m_use) -> float { return full_block ? emit_full(m_use, no_hook, no_hook) : emit_part(m_use); };
        auto bmax = [&]() -> float { return full_block ? block_max(FullT{}) : block_max(PartT{}); };
        // P is single-buffered in tensor memory: PV(i-1) must have consumed it (they retire in order, so O is quiescent too)
        if (i >= 1 && !split) ptx::mbar_wait(&bars->pv_done[t][(g - 1) & 1], ((g - 1) >> 1) & 1);
        ptx::tc_fence_after();
        PH(2);
        float blk_sum = 0.f;
        if (!warp_active) {
          // nothing to compute for these rows
        } else if (i == 0) {
          m_ref = bmax();
          blk_sum = emit(m_ref);
        } else {
          // optimistic: exponentiate against the trailing reference maximum (no dependence on this block's maximum, so
          // the MUFU work starts as soon as S is in registers).  The block maximum itself is only needed when a row may have
          // exceeded the reference by more than 2^8: a row sum <= 2^8 proves that no element did (every p <= its row sum), so
          // the 128-element max pass is skipped on all other blocks.
          blk_sum = split ? emit_full(m_ref, rest_of_S, p_free) : emit(m_ref);
          const bool suspect = !(blk_sum <= 256.0f);
          if (__any_sync(0xffffffffu, suspect)) {
            const float m_blk = bmax();
            const bool jump = (m_blk - m_ref) * scale_log2e > 8.0f;
            if (__any_sync(0xffffffffu, jump)) {
              // rescale O in TMEM, redo P against the new reference
              const float alpha = jump ? ex2((m_ref - m_blk) * scale_log2e) : 1.0f;
              if (jump) m_ref = m_blk;
              l_run *= alpha;
#pragma unroll
              for (int c = 0; c < OCOLS; c += 16) {
                uint32_t o[16];
                ptx::tmem_ld_32x32b_x16(tO + c, o);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
                ptx::tmem_st_32x32b_x16(tO + c, o);
              }
              blk_sum = emit(m_ref);
            }
          }
        }
        l_run += blk_sum;
        PH(3);
        // p_ready alternates between two barriers by block parity: P(i+1) is only stored after PV(i) retired (wait above), so
        // the softmax warps can never lap the MMA warp on a barrier (parity waits only tell adjacent phases apart).
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars->p_ready[t][g & 1]);
        PH(4);
        TL(n_ord, 1 + i);
      }
      PH_FLUSH();
      MS(2);  // key loop done
      // ---- epilogue: O / l -> bf16 -> global
      ptx::mbar_wait(&bars->pv_done[t][(g0 + n_kv - 1) & 1], ((g0 + n_kv - 1) >> 1) & 1);
      ptx::tc_fence_after();
      const float inv_l = 1.0f / l_run;
      const int q_local = q0 + t * QT + r;
      __nv_bfloat16* dst = O + (size_t)(it.batch * (size_t)Lq + q_local) * ldo + it.col0;
      // split unit: this CTA saw only its key range — unnormalised O (fp32), reference maximum and row sum go to the workspace
      float* pdst = nullptr;
      if (SPLIT && it.partial) {
        const size_t prow = (size_t)it.prow0 + t * QT + r;
        pdst = ws.partial + prow * (HD + 2);
        if (q_local < Lq) { pdst[HD] = m_ref; pdst[HD + 1] = l_run; }
      }
#pragma unroll
      for (int c = 0; c < OCOLS; c += 32) {
        uint32_t o[32];
        ptx::tmem_ld_32x32b_x32(tO + c, o);
        ptx::tmem_ld_wait();
        if (SPLIT && pdst != nullptr) {
          if (q_local < Lq) {
#pragma unroll
            for (int j = 0; j < 16; ++j) *reinterpret_cast<float2*>(pdst + c + 2 * j) = make_float2(__uint_as_float(o[2 * j]), __uint_as_float(o[2 * j + 1]));
          }
        } else if (q_local < Lq) {
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
    TL(n_ord, 7);
    g0 += (uint32_t)n_kv;
    }  // items of this CTA
  }
  MS(3);  // epilogue stores issued
  ptx::tc_fence_before();
  __syncthreads();
  MS(4);
  if (warp == W_MMA) ptx::tmem_dealloc(tmem, C::TMEM_COLS);
  MS(5);
}

// Workspace of the split units' partials: one buffer per stream (launches on one stream are ordered, so a buffer is reused by
// the next launch only after the previous merge has read it), grown on demand; never allocated while the stream is being
// captured into a CUDA graph (the launch then simply does not split).
float* split_workspace(cudaStream_t st, size_t bytes) {
  struct Buf { float* p = nullptr; size_t cap = 0; };
  static std::mutex mu;
  static std::map<cudaStream_t, Buf> bufs;
  std::lock_guard<std::mutex> lock(mu);
  Buf& b = bufs[st];
  if (b.cap >= bytes) return b.p;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return nullptr; }
  // a buffer that is outgrown is retired, not freed: a CUDA graph captured earlier on this stream may hold its address
  static std::vector<float*> retired;
  if (b.p) { retired.push_back(b.p); b.p = nullptr; b.cap = 0; }
  const size_t cap = bytes + bytes / 4;
  if (cudaMalloc(&b.p, cap) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; return nullptr; }
  b.cap = cap;
  return b.p;
}

template <int HD, int NT_, int POLY_ = 0>
int launch(const AttentionArgs& a, cudaStream_t st) {
  using C = Cfg<HD, NT_, POLY_>;
  const size_t rows_q = (size_t)a.batches * a.Lq, rows_k = (size_t)a.batches * a.Lk;
  const CUtensorMap* tq = tmap_2d_bf16(a.q, (uint64_t)a.heads * HD, rows_q, (uint64_t)a.ldq * 2, 64, QT);
  const CUtensorMap* tk = tmap_2d_bf16(a.k, (uint64_t)a.heads * HD, rows_k, (uint64_t)a.ldk * 2, 64, C::BKV);
  const CUtensorMap* tv = tmap_2d_bf16(a.v, (uint64_t)a.heads * HD, rows_k, (uint64_t)a.ldv * 2, 64, C::BKV);
  if (!tq || !tk || !tv) return LSVS_ECUDA;
  constexpr bool CAN_SPLIT = (HD == 64 && NT_ == 1);   // the key-range split is built for the long head-dim-64 passes only
  static bool configured = false;
  if (!configured) {
    LSVS_CUDA(cudaFuncSetAttribute(attention_fwd_tcgen05<HD, NT_, POLY_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    if (CAN_SPLIT) LSVS_CUDA(cudaFuncSetAttribute(attention_fwd_tcgen05<HD, NT_, POLY_, CAN_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    configured = true;
  }
  const float scale_log2e = a.scale * 1.4426950408889634f;
  WorkSplit ws;
  ws.n_qt = (a.Lq + C::NT * QT - 1) / (C::NT * QT);
  ws.heads = a.heads;
  const long long n_units = (long long)ws.n_qt * a.heads * a.batches;
  LSVS_CHECK_ARG(n_units < (1ll << 30), "attention: too many (batch, head, query tile) units");
  ws.n_full = (int)n_units; ws.parts = 1; ws.partial = nullptr;
  // Tail balancing: CTAs are dispatched in rounds of `slots`; a last round that fills only part of the SMs still lasts a whole
  // unit.  Cutting its units into p key ranges makes it last ceil(leftover * p / slots) / p of a unit (+ the merge).
  static const int split_max = [] { const char* e = getenv("LSVS_ATTN_SPLIT_MAX"); return e ? atoi(e) : 4; }();
  const int slots = num_sms() * (NT_ == 1 ? 2 : 1);
  const int n_kv_seq = (a.Lk + C::BKV - 1) / C::BKV;
  const int leftover = (int)(n_units % slots);
  if (CAN_SPLIT && n_units > slots && leftover > 0 && split_max > 1) {
    int best_p = 1; double best = 1.0;
    for (int p = 2; p <= split_max && n_kv_seq / p >= 8; ++p) {
      const double cost = (double)(((long long)leftover * p + slots - 1) / slots) / p + 0.03 * p;
      if (cost < best - 0.05) { best = cost; best_p = p; }
    }
    if (best_p > 1) {
      const size_t need = (size_t)leftover * best_p * (C::NT * QT) * (HD + 2) * sizeof(float);
      float* buf = split_workspace(st, need);
      if (buf) { ws.n_full = (int)(n_units - leftover); ws.parts = best_p; ws.partial = buf; }
    }
  }
  ws.n_items = (int)(ws.n_full + (n_units - ws.n_full) * ws.parts);
  // persistent CTAs (one-tile kernel, short sequences): one per slot, each walking its items.  Measured in the step
  // (profiles/r2b_attention_experiments.md): 412-token passes 4.45 -> 4.17 ms; on the long global pass the static round-robin
  // loses to the hardware's dynamic CTA dispatch (21.5 -> 22.2 ms), so long sequences keep one CTA per item.
  // LSVS_ATTN_PERSIST=0 / 2: never / always (A/B runs).
  static const int persist = [] { const char* e = getenv("LSVS_ATTN_PERSIST"); return e ? atoi(e) : 1; }();
  const bool persistent = NT_ == 1 && ws.n_items > slots && (persist == 2 || (persist == 1 && n_kv_seq <= 16));
  const dim3 grid((unsigned)(persistent ? slots : ws.n_items));
  auto kern = ws.parts > 1 ? attention_fwd_tcgen05<HD, NT_, POLY_, CAN_SPLIT> : attention_fwd_tcgen05<HD, NT_, POLY_, false>;
  LSVS_CUDA(launch_pdl(kern, grid, dim3(C::NTHREADS), C::SMEM, st, *tq, *tk, *tv, reinterpret_cast<__nv_bfloat16*>(a.o), a.ldo, a.Lq, a.Lk, scale_log2e, ws));
  LSVS_LAUNCH_CHECK();
  if (ws.parts > 1) {
    constexpr int ROWS = C::NT * QT;
    auto ck = attention_combine_kernel<HD, ROWS>;
    LSVS_CUDA(launch_pdl(ck, dim3((unsigned)((n_units - ws.n_full) * (ROWS / 4))), dim3(128), 0, st, (const float*)ws.partial,
                         reinterpret_cast<__nv_bfloat16*>(a.o), a.ldo, a.Lq, scale_log2e, ws.n_qt, ws.heads, ws.n_full, ws.parts));
    LSVS_LAUNCH_CHECK();
  }
  return LSVS_OK;
}

}  // namespace

int attention_fwd(const AttentionArgs& a, cudaStream_t st) {
  LSVS_CHECK_ARG(a.q && a.k && a.v && a.o, "attention: null pointer");
  LSVS_CHECK_ARG(a.batches > 0 && a.heads > 0 && a.Lq > 0 && a.Lk > 0, "attention: empty shape");
  LSVS_CHECK_ARG(a.head_dim == 64 || a.head_dim == 128, "attention: head_dim %d unsupported (64 or 128)", a.head_dim);
  const int D = a.heads * a.head_dim;
  LSVS_CHECK_ARG(a.ldq >= D && a.ldk >= D && a.ldv >= D && a.ldo >= D, "attention: leading dimension smaller than heads*head_dim");
  LSVS_CHECK_ARG(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0, "attention: leading dimensions must be multiples of 8");
  ProfScope prof(a.Lk >= 2048 ? PROF_ATTENTION_GLOBAL : PROF_ATTENTION, st, 4.0 * a.batches * (double)a.heads * a.Lq * (double)a.Lk * a.head_dim, 0);
  // One query tile per CTA and two CTAs per SM at every sequence length.  Short sequences: the prologue / epilogue of one CTA
  // hides behind the other.  Long sequences (measured, profiles/r2b_attention_experiments.md): two independent CTAs beat one
  // CTA whose two tiles share each K/V block — 746-752 vs 695-700 TFLOP/s at 13 184 tokens, 824 vs 732 at 21 984 — although
  // every K/V block is then fetched twice per SM: the two tiles of one CTA run in lock-step (their softmax warps wait for
  // their barriers and tensor-memory loads at the same time, leaving the XU pipe idle), two CTAs drift apart and fill
  // each other's waits.  LSVS_ATTN_NT1_MAX_LK=<n> restores the two-tile kernel for Lk > n (A/B runs).
  static const int nt1_max_lk = [] { const char* e = getenv("LSVS_ATTN_NT1_MAX_LK"); return e ? atoi(e) : 0x7fffffff; }();
  const bool one_tile = a.Lk <= nt1_max_lk;
  if (a.head_dim == 128) return one_tile ? launch<128, 1>(a, st) : launch<128, 2>(a, st);
  // head dim 64: share of the exponentials taken off the XU pipe (pairs out of 4; LSVS_ATTN_POLY overrides for A/B runs).  With two
  // independent CTAs per SM the XU pipe is the contended resource on long sequences (one pair of four on the FMA / ALU pipes:
  // 787 -> 850 TFLOP/s at 13 184 tokens, 842 -> 915 at 21 984; two pairs: slower); on 412-token sequences it costs 3 %.
  static const int poly_env = [] { const char* e = getenv("LSVS_ATTN_POLY"); return e ? atoi(e) : -1; }();
  const int poly = poly_env >= 0 ? poly_env : (a.Lk > 1024 ? 1 : LSVS_ATTN_POLY_DEFAULT);
  if (one_tile) return poly == 1 ? launch<64, 1, 1>(a, st) : poly == 2 ? launch<64, 1, 2>(a, st) : launch<64, 1, 0>(a, st);
  return poly == 1 ? launch<64, 2, 1>(a, st) : poly == 2 ? launch<64, 2, 2>(a, st) : launch<64, 2, 0>(a, st);
}

}  // namespace lsvs

#ifdef LSVS_ATTN_PHASES
extern "C" int lsvs_debug_attn_timeline(unsigned long long* out64) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out64, lsvs::g_attn_tl, 64 * sizeof(unsigned long long));
}
extern "C" int lsvs_debug_attn_phases(unsigned long long* out64) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out64, lsvs::g_attn_phase, 64 * sizeof(unsigned long long));
}
#endif

#ifdef LSVS_DEBUG_HANG
extern "C" int lsvs_debug_hang_read(int* out257, int* bar_base_offset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out257, ptx::g_lsvs_hang, 257 * sizeof(int));
  cudaMemcpyFromSymbol(out257 + 257, lsvs::g_dbg_iter, 64 * sizeof(int));
  *bar_base_offset = lsvs::Cfg<64, 2, 0>::OFF_BAR;
  static int zero[257] = {0};
  cudaMemcpyToSymbol(ptx::g_lsvs_hang, zero, sizeof(zero));
  return 0;
}
#endif
