// tcgen05 / TMEM flash attention for head dims < 64 (the 64 x 64 level: d_head 40, 4096 + G keys), third layout:
// ONE CTA per SM that carries THREE query tiles (384 rows) through the key tiles, one softmax warpgroup and one MMA
// issuer warp per query tile.
//
// What the clock probes of the earlier layouts showed (scripts/micro/attn_l0.py, PROBE=1; profiles/r2_attention_l0.md):
// attention_tc.cu (row pairs), attention_tc2.cu (two warpgroups splitting the key tiles, two CTAs per SM) and a
// four-warpgroup variant of it all ran the level at ~435 us whatever was removed from them - exp2, TMEM loads, even the
// MMAs (knock-outs).  The time was in two serial chains:
//   * ONE MMA-issuer thread per CTA doing wait(P_j) -> issue P.V_j -> wait(K_{j+n}) -> issue QK^T_{j+n} -> commit for every
//     tile of every warpgroup: ~750 clk of single-thread latency per 128 x 64 block against ~190 clk of tensor-pipe work;
//   * one K / V stream per 128 query rows: 10 KB of 80-byte rows per block per SM - the tiles arrived later than a
//     12-deep ring could hide.
// Here
//   * warpgroup g owns query tile g and ALL key tiles: no cross-warpgroup merge, and every K_j / V_j tile that lands in
//     shared memory feeds three query tiles (a third of the L2 -> SM traffic and TMA requests per score);
//   * every warpgroup has its own issuer warp (warps 1..3): the wait -> issue -> commit chains of different warpgroups
//     never queue behind each other;
//   * S_g, P_g and O_g have their own TMEM columns (3 x (64 + 32 + DPAD) <= 512).  A softmax thread copies its whole
//     64-score row to registers with ONE tcgen05.ld round trip (128-register budget: 16 warps, one CTA per SM) and frees
//     S_g at once: QK^T_{j+1} runs under the softmax of tile j, P.V_j under the start of tile j+1;
//   * the ones column of V (row sums come out of P.V: the spare column d_head of O_g) is written by the TMA warp a few
//     tiles behind its loads, off every issuer's path.
// What bounds it now: the MUFU (16 exp2 / clk / SM = 512 clk per 128 x 64 block; the three warps of a scheduler share
// one MUFU port) - so a share of the exponentials runs as a Cody-Waite + cubic polynomial on the FMA pipe (POLY).
//
//   warp 0      : TMA producer (Q_0..2 once, then K_j / V_j through a 10-stage ring) + ones column of V_{j-4}
//   warps 1..3  : MMA issuer of warpgroup g:  QK^T_{j+1} -> S_g when S_g is free,  P.V_j -> O_g when P_j is there
//   warps 4..15 : softmax warpgroups (warp & 3 = TMEM lane quarter)
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string>

#include "common.cuh"
#include "internal.h"
#include "../../include/gligen_b200.h"

namespace glg {

struct AttnTc3Params {
  bf16* o; long long o_row, o_batch;
  int heads, d, Lq, Lk;
  float scale_log2;
  int stagger;            // clocks by which warpgroup g delays its first tile (g * stagger): takes the warpgroups' exp2 phases out of lockstep
  long long* probe;       // KO & 64 instantiation only: per-phase clock sums of CTA (0,0,0), [warp][8]
};

namespace atc3 {
constexpr int BM = 128, BN = 64, NWG = 3, STAGES = 10, THREADS = 128 + NWG * 128;
constexpr int Q_BYTES = BM * 64 * 2;        // 16 KB per warpgroup (one 64-column swizzle atom: d_head <= 64)
constexpr int KV_BYTES = BN * 64 * 2;       // 8 KB
constexpr int SMEM_BYTES = NWG * Q_BYTES + STAGES * 2 * KV_BYTES + 1024 + 512;
constexpr int TMEM_COLS = 512;
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units
}  // namespace atc3

__device__ __forceinline__ float ex2_approx3(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA / ALU pipes (Cody-Waite split + Taylor cubic on [-0.5, 0.5]: relative error 1.2e-4 mean, 7.9e-4 max at |f| = 0.5,
// against 2e-3 for the bf16 rounding of P; tests/test_exp2_poly_cpu.py): a share of the exponentials leaves the MUFU pipe (16 exp2 / clk / SM).
__device__ __forceinline__ float ex2_poly3(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;                 // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.05550411f, f, 0.24022651f);
  p = fmaf(p, f, 0.69314718f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// POLY: of every 8 score pairs, the last POLY pairs take their exp2 from ex2_poly3 instead of the MUFU
// KO (timing experiments only, results are wrong): 1 no exp2 (plain FMA instead), 2 K / V tiles loaded only for the first ring
// pass, 4 no P.V MMAs, 8 no QK^T MMAs, 16 no TMEM score loads, 32 no P store, 64 clock probes of CTA (0,0,0)
// VAR (measured variants of the per-tile chains, profiles/r2_attention_l0.md; 4 = the issuer warps' waits by lane 0 only;
// 8 / 16 = busy-polling (mbarrier.test_wait) waits in the softmax / issuer warps):
//   1  mbarrier waits by lane 0 only + __syncwarp (32 lanes polling the same mbarrier serialise in the shared-memory pipe)
//   2  the exponentials of the whole tile are computed BEFORE the wait for P.V of the previous tile (only the P stores and the rare
//      O rescale need it), taking that wait off the per-tile critical path
template <int DPAD, int POLY = 0, int KO = 0, int VAR = 0>      // head dim rounded up to a multiple of 16 with one spare column for the row sums: d_head < DPAD <= 64
__global__ void __launch_bounds__(atc3::THREADS, 1)
attn_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnTc3Params p) {
  using namespace atc3;
  constexpr int P_COL = NWG * BN;                      // S_g: 64 fp32 columns each | P_g: 32 columns (64 bf16) each | O_g: DPAD each
  constexpr int O_COL = P_COL + NWG * (BN / 2);
  static_assert(O_COL + NWG * DPAD <= TMEM_COLS, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sKV = base + NWG * Q_BYTES;           // per stage: K | V
  const uint32_t bar_base = sKV + STAGES * 2 * KV_BYTES;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };                       // TMA landed (K and V)
  auto kv_ready = [&](int s) { return bar_base + 8u * (1 + STAGES + s); };             // ... and V carries its ones column
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + 2 * STAGES + s); };         // P.V_j done with the stage
  auto s_full = [&](int g) { return bar_base + 8u * (1 + 3 * STAGES + g); };           // QK^T_j in S_g
  auto s_free = [&](int g) { return bar_base + 8u * (1 + 3 * STAGES + NWG + g); };     // S_g copied to registers (4 warps)
  auto p_full = [&](int g) { return bar_base + 8u * (1 + 3 * STAGES + 2 * NWG + g); }; // P_j in P_g (4 warps)
  auto p_free = [&](int g) { return bar_base + 8u * (1 + 3 * STAGES + 3 * NWG + g); }; // P.V_j done: P_g reusable, O_g up to date
  auto o_full = [&](int g) { return bar_base + 8u * (1 + 3 * STAGES + 4 * NWG + g); }; // last P.V of warpgroup g done
  const uint32_t tmem_slot = bar_base + 8u * (1 + 3 * STAGES + 5 * NWG);
  static_assert(8 * (2 + 3 * STAGES + 5 * NWG) <= 512, "barrier area");

  pdl_trigger();
  long long t_prev = 0, stamp[6] = {0, 0, 0, 0, 0, 0};
  const bool probe_on = (KO & 64) && p.probe && (threadIdx.x & 31) == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define TC3_STAMP(k) do { if ((KO & 64) && probe_on) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_)); stamp[k] += t_ - t_prev; t_prev = t_; } } while (0)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (NWG * BM);              // warpgroup g owns query rows [q0 + g * 128, +128)
  const int h = blockIdx.y, b = blockIdx.z;
  const int nkt = (p.Lk + BN - 1) / BN;
  const int nact = min(NWG, (p.Lq - q0 + BM - 1) / BM);   // warpgroups with at least one real query row

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_ready(s), 1); mbar_init(kv_empty(s), nact); }
    for (int g = 0; g < NWG; ++g) { mbar_init(s_full(g), 1); mbar_init(s_free(g), 4); mbar_init(p_full(g), 4); mbar_init(p_free(g), 1); mbar_init(o_full(g), 1); }
    fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();
  if ((KO & 64) && probe_on) asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_prev));

  if (warp == 0) {
    // ===================== TMA producer + ones column of V (row sums come out of P.V) =====================
    // The ones are written FIX_LAG tiles behind the loads: that tile has long landed (12-stage ring), so the producer
    // never stalls on it and no issuer ever waits for it.
    constexpr int FIX_LAG = 4;
    const bool leader = elect_one();
    if (leader) {
      mbar_arrive_expect_tx(q_full, nact * Q_BYTES);
      for (int g = 0; g < nact; ++g) tma_load_4d(sQ + g * Q_BYTES, &tmQ, q_full, 0, h, q0 + g * BM, b);
    }
    auto fix_v = [&](int jj) {       // V_jj[key][d_head] = 1 (128-byte swizzled rows: 16-byte chunk index ^ (row & 7))
      const int stage = jj % STAGES;
      mbar_wait(kv_full(stage), (uint32_t)((jj / STAGES) & 1));
      const uint32_t svt = sKV + stage * 2 * KV_BYTES + KV_BYTES;
#pragma unroll
      for (int r = lane; r < BN; r += 32) {
        const uint32_t a = svt + (uint32_t)r * 128u + ((((uint32_t)p.d >> 3) ^ ((uint32_t)r & 7u)) << 4) + (((uint32_t)p.d & 7u) << 1);
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)0x3F80) : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(kv_ready(stage));
    };
    int stage = 0; uint32_t phase = 0;
    for (int j = 0; j < nkt; ++j) {
      mbar_wait(kv_empty(stage), phase ^ 1u);
      TC3_STAMP(0);
      if (leader) {
        if ((KO & 2) && j >= STAGES) {
          mbar_arrive_expect_tx(kv_full(stage), 0);
        } else {
          mbar_arrive_expect_tx(kv_full(stage), 2 * KV_BYTES);
          const uint32_t sk = sKV + stage * 2 * KV_BYTES;
          tma_load_4d(sk, &tmK, kv_full(stage), 0, h, j * BN, b);
          tma_load_4d(sk + KV_BYTES, &tmV, kv_full(stage), 0, h, j * BN, b);
        }
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      TC3_STAMP(1);
      if (j >= FIX_LAG) fix_v(j - FIX_LAG);
      TC3_STAMP(2);
    }
    for (int jj = (nkt > FIX_LAG ? nkt - FIX_LAG : 0); jj < nkt; ++jj) fix_v(jj);
  } else if (warp < 1 + NWG) {
    // ===================== MMA issuer of warpgroup g: S_g = Q_g K_j^T and O_g += P_j V_j for every key tile j =====================
    // Every warpgroup has its own issuer, so the (slow, single-thread) wait -> issue -> commit chains of different
    // warpgroups never queue behind each other.  QK^T of the NEXT tile goes out as soon as the softmax warps have copied
    // S_g to registers (early in their tile), P.V when they are done with it.
    const int g = warp - 1;
    if (g < nact) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, BN);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DPAD, true);    // B (V) is MN-major
      const bool leader = elect_one();
      const uint32_t sQg = sQ + g * Q_BYTES;
      auto mbar_wait_i = [&](uint32_t bar, uint32_t parity) {      // VAR & 4: the issuer warps' waits by lane 0 only; & 16: busy polls
        if (VAR & 16) {
          mbar_wait_spin(bar, parity);
        } else if (VAR & 4) {
          if (lane == 0) mbar_wait(bar, parity);
          __syncwarp();
        } else {
          mbar_wait(bar, parity);
        }
      };
      auto issue_qk = [&](int j) {
        const int stage = j % STAGES;
        mbar_wait_i(kv_full(stage), (uint32_t)((j / STAGES) & 1));
        tc_fence_after();
        if (leader) {
          const uint32_t sk = sKV + stage * 2 * KV_BYTES;
#pragma unroll
          for (int kk = 0; kk < ((KO & 8) ? 0 : DPAD / 16); ++kk)
            umma_bf16(tmem_base + g * BN, umma_desc_kmajor_sw128(sQg) + 2 * kk, umma_desc_kmajor_sw128(sk) + 2 * kk, idesc_s, kk != 0 ? 1u : 0u);
          umma_commit(s_full(g));
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < nkt; ++j) {
        if (j + 1 < nkt) {
          mbar_wait_i(s_free(g), (uint32_t)(j & 1));      // S_g of tile j is in registers
          TC3_STAMP(0);
          issue_qk(j + 1);
          TC3_STAMP(1);
        }
        const int stage = j % STAGES;
        mbar_wait_i(kv_ready(stage), (uint32_t)((j / STAGES) & 1));
        mbar_wait_i(p_full(g), (uint32_t)(j & 1));
        tc_fence_after();
        TC3_STAMP(2);
        if (leader) {
          const uint64_t vdesc = umma_desc_mnmajor_sw128(sKV + stage * 2 * KV_BYTES + KV_BYTES, KV_BYTES);
#pragma unroll
          for (int k = 0; k < ((KO & 4) ? 0 : BN / 16); ++k)           // 16 keys per step: +8 TMEM columns of P, +16 rows (2048 B) of V
            umma_bf16_ts(tmem_base + O_COL + g * DPAD, tmem_base + P_COL + g * (BN / 2) + 8 * k, vdesc + 128 * k, idesc_o, (j > 0 || k != 0) ? 1u : 0u);
          umma_commit(kv_empty(stage));                  // one of nact arrivals: every warpgroup's P.V_j must be done with the stage
          umma_commit(p_free(g));
        }
        __syncwarp();
        TC3_STAMP(3);
      }
      if (leader) umma_commit(o_full(g));
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int g = (warp - 4) >> 2;               // warpgroup: owns query tile g of this CTA, all key tiles
    const int q = warp & 3;                      // TMEM lane quarter
    const int rloc = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t s_addr = lane_base + g * BN;
    const uint32_t p_addr = lane_base + P_COL + g * (BN / 2);
    const uint32_t o_addr = lane_base + O_COL + g * DPAD;
    const int row = q0 + g * BM + rloc;
    const float sl2 = p.scale_log2;
    float m_ref = -INFINITY;
    if (p.stagger > 0 && g > 0) {
      const long long t_end = clock64() + (long long)g * p.stagger;
      while (clock64() < t_end) {}
    }
    auto wait_bar = [&](uint32_t bar, uint32_t parity) {
      if (VAR & 8) {
        mbar_wait_spin(bar, parity);
      } else if (VAR & 1) {
        if (lane == 0) mbar_wait(bar, parity);
        __syncwarp();
      } else {
        mbar_wait(bar, parity);
      }
    };
    for (int j = 0, it = 0; j < (g < nact ? nkt : 0); ++j, ++it) {
      wait_bar(s_full(g), (uint32_t)(it & 1));
      tc_fence_after();
      TC3_STAMP(0);
      const int valid = p.Lk - j * BN;           // < 64 only in the last tile: keys [valid, 64) are padding
      // the thread's whole row of this tile: ONE trip through the TMEM read port; S_g is free again right after
      uint32_t lo[32], hi[32];
#define SC(i) ((i) < 32 ? lo[(i) & 31] : hi[(i) & 31])
      if (KO & 16) {
#pragma unroll
        for (int i = 0; i < 32; ++i) { lo[i] = __float_as_uint((float)(i + j) * 0.01f); hi[i] = __float_as_uint((float)(i - j) * 0.01f); }
      } else {
        tmem_ld32(s_addr, lo);
        tmem_ld32(s_addr + 32, hi);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free(g));
      if (valid < BN) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= valid) lo[i] = 0xff800000u;
          if (32 + i >= valid) hi[i] = 0xff800000u;
        }
      }
      TC3_STAMP(1);
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(SC(i)), __uint_as_float(SC(i + 1))));
        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(SC(i + 2)), __uint_as_float(SC(i + 3))));
      }
      const float m_new = fmaxf(m_ref, fmaxf(mx0, mx1));
      bool waited = false;
      if (it == 0) {
        m_ref = m_new;
      } else {
        const bool need = (m_new - m_ref) * sl2 > RESCALE_THRESHOLD;
        const bool any_need = __any_sync(0xffffffffu, need);     // tcgen05.ld / st are warp-collective: the whole warp rescales its rows
        if (!(VAR & 2) || any_need) {
          // P.V of this warpgroup's previous tile must be done before P_g is overwritten and before O_g is rescaled
          wait_bar(p_free(g), (uint32_t)((it - 1) & 1));
          tc_fence_after();
          waited = true;
        }
        if (any_need) {
          const float f = need ? ex2_approx3((m_ref - m_new) * sl2) : 1.0f;
          if (need) m_ref = m_new;
          // rare path, 4 columns at a time: the 64 scores stay live across it
#pragma unroll 1
          for (int c = 0; c < DPAD; c += 4) {
            uint32_t o0, o1, o2, o3;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(o_addr + c) : "memory");
            tmem_ld_wait();
            o0 = __float_as_uint(__uint_as_float(o0) * f); o1 = __float_as_uint(__uint_as_float(o1) * f);
            o2 = __float_as_uint(__uint_as_float(o2) * f); o3 = __float_as_uint(__uint_as_float(o3) * f);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o_addr + c), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
          }
          tmem_st_wait();
        }
      }
      TC3_STAMP(2);
      // P = exp2((S - m_ref) * scale) -> bf16 in P_g
      const float ms = m_ref * sl2;
      if (VAR & 2) {
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x0 = fmaf(__uint_as_float(SC(2 * i)), sl2, -ms), x1 = fmaf(__uint_as_float(SC(2 * i + 1)), sl2, -ms);
          const float a0 = (KO & 1) ? x0 * 0.001f : ((i & 7) >= 8 - POLY) ? ex2_poly3(x0) : ex2_approx3(x0);
          const float a1 = (KO & 1) ? x1 * 0.001f : ((i & 7) >= 8 - POLY) ? ex2_poly3(x1) : ex2_approx3(x1);
          pk[i] = pack_bf16x2(a0, a1);
        }
        if (it > 0 && !waited) {
          wait_bar(p_free(g), (uint32_t)((it - 1) & 1));
          tc_fence_after();
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                       ::"r"(p_addr + k * 8), "r"(pk[8 * k]), "r"(pk[8 * k + 1]), "r"(pk[8 * k + 2]), "r"(pk[8 * k + 3]), "r"(pk[8 * k + 4]), "r"(pk[8 * k + 5]),
                         "r"(pk[8 * k + 6]), "r"(pk[8 * k + 7]) : "memory");
      } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x0 = fmaf(__uint_as_float(SC(k * 16 + 2 * i)), sl2, -ms), x1 = fmaf(__uint_as_float(SC(k * 16 + 2 * i + 1)), sl2, -ms);
          const float a0 = (KO & 1) ? x0 * 0.001f : (i >= 8 - POLY) ? ex2_poly3(x0) : ex2_approx3(x0);     // padding keys: exp2(-inf) = 0 on both paths
          const float a1 = (KO & 1) ? x1 * 0.001f : (i >= 8 - POLY) ? ex2_poly3(x1) : ex2_approx3(x1);
          pk[i] = pack_bf16x2(a0, a1);
        }
        if (KO & 32) { if (pk[0] + pk[1] + pk[2] + pk[3] + pk[4] + pk[5] + pk[6] + pk[7] == 0x12345u) m_ref += 1.f; }
        else asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"r"(p_addr + k * 8), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
      }
      }
      TC3_STAMP(3);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(g));
      TC3_STAMP(4);
#undef SC
    }
    // ---- O_g / l_g -> global (the row sums l_g rode in column d_head of O_g)
    if (g < nact) {
      mbar_wait(o_full(g), 0);
      tc_fence_after();
      const uint32_t lu = tmem_ld1(o_addr + p.d);
      tmem_ld_wait();
      const float inv = 1.0f / __uint_as_float(lu);
      bf16* orow = p.o + (long long)b * p.o_batch + (long long)row * p.o_row + (long long)h * p.d;
#pragma unroll
      for (int c = 0; c < DPAD / 16; ++c) {
        uint32_t a[16];
        tmem_ld16(o_addr + c * 16, a);
        tmem_ld_wait();
        if (row < p.Lq) {
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const int col = c * 16 + gg * 8;
            if (col < p.d) {           // d is a multiple of 8: whole 8-column groups are valid or not
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(a[gg * 8 + 0]) * inv, __uint_as_float(a[gg * 8 + 1]) * inv);
              u.y = pack_bf16x2(__uint_as_float(a[gg * 8 + 2]) * inv, __uint_as_float(a[gg * 8 + 3]) * inv);
              u.z = pack_bf16x2(__uint_as_float(a[gg * 8 + 4]) * inv, __uint_as_float(a[gg * 8 + 5]) * inv);
              u.w = pack_bf16x2(__uint_as_float(a[gg * 8 + 6]) * inv, __uint_as_float(a[gg * 8 + 7]) * inv);
              *reinterpret_cast<uint4*>(orow + col) = u;
            }
          }
        }
      }
    }
  }

  if ((KO & 64) && probe_on) {
    TC3_STAMP(5);
    for (int k = 0; k < 6; ++k) p.probe[warp * 8 + k] = stamp[k];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, atc3::TMEM_COLS);
  }
}

extern long long* g_attn_probe;  // attention_tc.cu: glg_debug_attn_probe
extern int g_attn_tc2_poly;      // attention_tc2.cu: GLG_ATTN_POLY / glg_debug_attn_poly_share; pairs of 8 whose exp2 runs on the FMA pipe

int g_attn_tc3_ko = 0;
int g_attn_tc3_stagger = -1;     // -1: GLG_ATTN_STAGGER env (default 0)

int g_attn_tc3_var = -1;         // -1: GLG_ATTN_TC3_VAR env (default below)

template <int DPAD, int POLY, int KO = 0, int VAR = 0>
static int launch_attn_tc3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTc3Params& p, int B, cudaStream_t st) {
  static bool attr_set = false;
  auto kern = attn_tc3_kernel<DPAD, POLY, KO, VAR>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, atc3::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(std::string("cudaFuncSetAttribute(attn_tc3): ") + cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((p.Lq + atc3::NWG * atc3::BM - 1) / (atc3::NWG * atc3::BM), p.heads, B);
  launch_k(kern, grid, dim3(atc3::THREADS), atc3::SMEM_BYTES, st, 1, tq, tk, tv, p);
  count_launch();
  return check_launch("attention_tc3 launch");
}

// Returns 1 if this path does not apply (caller falls back), 0 on success, -1 on error.
// Applies to d_head < 64 with a spare column up to the next multiple of 16 (d_head = 40: DPAD = 48), i.e. d_head % 16 != 0.
int attention_tc3(const GlgAttnArgs* a, cudaStream_t st) {
  const int dpad = (a->d_head + 15) / 16 * 16;
  if (dpad > 64 || dpad == a->d_head) return 1;
  if ((a->d_head % 8) || (a->o_row % 8) || (a->o_batch % 8) || ((uintptr_t)a->out & 15)) return 1;
  CUtensorMap tq, tk, tv;
  const uint64_t d = a->d_head, hd = a->heads;
  {
    const uint64_t dims[4] = {d, hd, (uint64_t)a->Lq, (uint64_t)a->B};
    const uint64_t str[3] = {d * 2, (uint64_t)a->q_row * 2, (uint64_t)a->q_batch * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)atc3::BM, 1};
    if (get_tmap_bf16(&tq, a->q, 4, dims, str, box)) return -1;
  }
  {
    const uint64_t dims[4] = {d, hd, (uint64_t)a->Lk, (uint64_t)a->B};
    const uint64_t strk[3] = {d * 2, (uint64_t)a->k_row * 2, (uint64_t)a->k_batch * 2};
    const uint64_t strv[3] = {d * 2, (uint64_t)a->v_row * 2, (uint64_t)a->v_batch * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)atc3::BN, 1};
    if (get_tmap_bf16(&tk, a->k, 4, dims, strk, box)) return -1;
    if (get_tmap_bf16(&tv, a->v, 4, dims, strv, box)) return -1;
  }
  AttnTc3Params p;
  p.o = (bf16*)a->out; p.o_row = a->o_row; p.o_batch = a->o_batch;
  p.heads = a->heads; p.d = a->d_head; p.Lq = a->Lq; p.Lk = a->Lk;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.probe = g_attn_probe;
  if (g_attn_tc3_stagger < 0) { const char* e = getenv("GLG_ATTN_STAGGER"); g_attn_tc3_stagger = e ? atoi(e) : 0; }
  p.stagger = g_attn_tc3_stagger;
  // FMA-pipe share of the exponentials: measured best at 2 of every 8 score pairs for this kernel (level 0, 2B = 8: 409.6 us
  // all-MUFU, 391.3 / 389.9 / 405.5 / 437.5 us at 1 / 2 / 3 / 4 of 8; profiles/r2_attention_l0.md).  GLG_ATTN_POLY or the debug
  // setter override it; the two-warpgroup kernel keeps its own default (0).
  int poly = g_attn_tc2_poly;
  if (poly < 0) { const char* e = getenv("GLG_ATTN_POLY"); poly = e ? atoi(e) : 2; }
  // measured (profiles/r2_attention_l0.md, level 0, 2B = 8): variant 0 389.2 us, 1 (lane-0 waits) 434.0, 2 (exponentials before the
  // P.V wait) 385.9, 3 413.9, 5 / 7 (lane-0 waits in the issuer warps) 653 / 647 -> default 2
  if (g_attn_tc3_var < 0) { const char* e = getenv("GLG_ATTN_TC3_VAR"); g_attn_tc3_var = e ? atoi(e) : 2; }
  const int var = g_attn_tc3_var;
  switch (dpad) {
    case 16: return launch_attn_tc3<16, 0>(tq, tk, tv, p, a->B, st);
    case 32: return launch_attn_tc3<32, 0>(tq, tk, tv, p, a->B, st);
    case 48:
      switch (poly) {
        case 1: return launch_attn_tc3<48, 1>(tq, tk, tv, p, a->B, st);
        case 2:
          switch (var) {
            case 1: return launch_attn_tc3<48, 2, 0, 1>(tq, tk, tv, p, a->B, st);
            case 2: return launch_attn_tc3<48, 2, 0, 2>(tq, tk, tv, p, a->B, st);
            case 3: return launch_attn_tc3<48, 2, 0, 3>(tq, tk, tv, p, a->B, st);
            case 5: return launch_attn_tc3<48, 2, 0, 5>(tq, tk, tv, p, a->B, st);
            case 7: return launch_attn_tc3<48, 2, 0, 7>(tq, tk, tv, p, a->B, st);
            case 10: return launch_attn_tc3<48, 2, 0, 10>(tq, tk, tv, p, a->B, st);
            case 18: return launch_attn_tc3<48, 2, 0, 18>(tq, tk, tv, p, a->B, st);
            case 26: return launch_attn_tc3<48, 2, 0, 26>(tq, tk, tv, p, a->B, st);
            default: return launch_attn_tc3<48, 2>(tq, tk, tv, p, a->B, st);
          }
        case 3: return launch_attn_tc3<48, 3>(tq, tk, tv, p, a->B, st);
        case 4: return launch_attn_tc3<48, 4>(tq, tk, tv, p, a->B, st);
        default:
          switch (g_attn_tc3_ko) {
            case 1: return launch_attn_tc3<48, 0, 1>(tq, tk, tv, p, a->B, st);
            case 2: return launch_attn_tc3<48, 0, 2>(tq, tk, tv, p, a->B, st);
            case 4: return launch_attn_tc3<48, 0, 4>(tq, tk, tv, p, a->B, st);
            case 8: return launch_attn_tc3<48, 0, 8>(tq, tk, tv, p, a->B, st);
            case 12: return launch_attn_tc3<48, 0, 12>(tq, tk, tv, p, a->B, st);
            case 16: return launch_attn_tc3<48, 0, 16>(tq, tk, tv, p, a->B, st);
            case 17: return launch_attn_tc3<48, 0, 17>(tq, tk, tv, p, a->B, st);
            case 49: return launch_attn_tc3<48, 0, 49>(tq, tk, tv, p, a->B, st);
            case 14: return launch_attn_tc3<48, 0, 14>(tq, tk, tv, p, a->B, st);
            case 64: return launch_attn_tc3<48, 0, 64>(tq, tk, tv, p, a->B, st);
            default: return launch_attn_tc3<48, 0>(tq, tk, tv, p, a->B, st);
          }
      }
    case 64: return launch_attn_tc3<64, 0>(tq, tk, tv, p, a->B, st);
  }
  return 1;
}

}  // namespace glg

extern "C" void glg_debug_attn_tc3_knockout(int ko) { glg::g_attn_tc3_ko = ko; }
extern "C" void glg_debug_attn_tc3_stagger(int clocks) { glg::g_attn_tc3_stagger = clocks; }
extern "C" void glg_debug_attn_tc3_variant(int var) { glg::g_attn_tc3_var = var; }
