// tcgen05 / TMEM flash attention for head dims <= 160 (d_head = 40 at the 64x64-latent level carries 88 % of the
// attention FLOPs, d_head = 80 at 32x32 most of the rest, d_head = 160 at 16x16 / 8x8).  Head dims above 64
// use two or three 64-column swizzle atoms per operand tile.
//
//   per CTA: one (batch, head, 128-query tile); key/value tiles of 64 keys stream through a TMA ring.
//   warp 0 : TMA producer (Q once, then K_j / V_j tiles; 4-D tensor maps {d, head, row, batch},
//            box {64, 1, rows, 1}: columns >= d and rows >= L are zero-filled by TMA)
//   warp 1 : MMA issuer   S_j = Q K_j^T          (SS: both operands K-major in smem, N = 64)
//                         O  += P_j V_j           (TS: P_j bf16 in TMEM, V_j MN-major in smem)
//            Both run the whole warp through their loops with one elect.sync leader issuing (common.cuh elect_one()).
//   warps 2..9 : softmax  two threads per query row (32 of the tile's 64 keys each; partial row maxima are
//            exchanged through smem): tcgen05.ld S_j -> running max (lazy rescale of O only when the max grows by
//            > 2^8; only that rare path waits for P.V) -> exp2 -> P_j bf16 -> tcgen05.st.
//   TMEM (256 columns; 512 for d_head > 128): S[NSB] (64 fp32 each; P_j, 32 packed-bf16 columns, overwrites the first
//   half of S_j once both warps of a row pair hold their scores in registers) | O (<= 160 fp32).  NSB = 3 for
//   d_head <= 64 and > 128 (QK^T runs two key tiles ahead), 2 between.  Two CTAs are resident per SM (one for d_head > 128).
//   When d_head is not a multiple of 16 the spare V column carries 1.0, so the row sums come out of the P.V MMA.
//
// Measured on B200 (profiles/r1_attention_pipeline.md): exp2 16/clk/SM, tcgen05.ld ~466 B/clk/SM, so a 128x128 score
// block costs >= 1024 clk of MUFU against ~400 clk of tensor pipe: the kernel is bound by the softmax warps.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string>

#include "common.cuh"
#include "internal.h"
#include "../../include/gligen_b200.h"

namespace glg {

struct AttnTcParams {
  bf16* o; long long o_row, o_batch;
  int heads, d, Lq, Lk;
  float scale_log2;
  long long* probe;      // PROBE instantiation only: per-phase clock stamps of one softmax warp (scripts/micro)
  int dbg;               // PROBE instantiation only: bit flags that knock out one pipeline piece (scripts/micro)
};

namespace atc {
constexpr int BM = 128, BN = 64;
constexpr int QA_BYTES = BM * 64 * 2;       // one 64-column atom of Q: 16 KB
constexpr int KVA_BYTES = BN * 64 * 2;      // one atom of K or V: 8 KB
constexpr int XCH_BYTES = 2 * 2 * 128 * 4;  // [parity][column half][row] partial maxima / sums
constexpr int S_COL = 0;                    // S buffers first, then O; P_j lives in the first 32 columns of S_j
template <int DPAD> struct Cfg {
  static constexpr int NATOM = (DPAD + 63) / 64;
  static constexpr int STAGES = NATOM == 1 ? 4 : (DPAD > 128 ? 3 : 2);     // >= NSB: QK^T_{j+NSB-1} is issued before P.V_j frees its stage
  // d_head <= 128: S buffers + O fit 256 columns (two CTAs per SM); 144 / 160 take the whole TMEM (one CTA per SM)
  static constexpr int TMEM_COLS = DPAD <= 128 ? 256 : 512;
  static constexpr int NSB = (DPAD <= 64 || DPAD > 128) ? 3 : 2;
  static constexpr int Q_BYTES = NATOM * QA_BYTES;
  static constexpr int KV_BYTES = NATOM * KVA_BYTES;          // K (or V) of one stage
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * 2 * KV_BYTES + 1024 + 256 + XCH_BYTES;
};
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units
}  // namespace atc

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#define ATTN_STAMP(k) do { if (PROBE && probe_on) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_)); stamp[k] += t_ - t_prev; t_prev = t_; } } while (0)

#define EX2PAIR(x) ((PROBE && (p.dbg & 1)) ? (x) : ex2_approx(x))
#define ATTN_EV(jj, slot) do { if (PROBE && ev_on && (jj) >= 8 && (jj) < 16) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_)); p.probe[((jj) - 8) * 16 + (slot)] = t_; } } while (0)
// ONES (d_head < DPAD): the MMA warp writes 1.0 into column d_head of every V tile, so O[:, d_head] accumulates the
// row sum of the bf16 P the tensor core actually multiplied - the softmax loop then carries no FADD per score and no
// running sum to rescale.  (Measured at level 0: no change in time - profiles/r1_attention_pipeline.md shows the loop is
// latency-, not issue-bound - kept because the row sum then matches the numerator's bf16 rounding exactly.)
template <int DPAD, bool PROBE = false, bool ONES = false>   // head dim rounded up to a multiple of 16 (<= 160)
__global__ void __launch_bounds__(320, DPAD <= 128 ? 2 : 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  using namespace atc;
  using C = Cfg<DPAD>;
  constexpr int NATOM = C::NATOM, STAGES = C::STAGES;
  // S buffers: three when O leaves room (d_head <= 64), so QK^T runs TWO key tiles ahead of the softmax warps and
  // their s_full wait never sees the MMA issuer's per-tile latency (wake-up, P.V issue, commit, next QK^T issue).
  constexpr int NSB = C::NSB;
  constexpr int TMEM_COLS = C::TMEM_COLS;
  constexpr int O_COL = NSB * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sKV = base + C::Q_BYTES;              // per stage: K atoms | V atoms
  const uint32_t bar_base = sKV + STAGES * 2 * C::KV_BYTES;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + STAGES + s); };
  auto s_full = [&](int b) { return bar_base + 8u * (1 + 2 * STAGES + b); };
  auto p_full = [&](int b) { return bar_base + 8u * (1 + 2 * STAGES + NSB + b); };
  auto pv_done = [&](int b) { return bar_base + 8u * (1 + 2 * STAGES + 2 * NSB + b); };
  const uint32_t o_full = bar_base + 8u * (1 + 2 * STAGES + 3 * NSB);
  const uint32_t tmem_slot = bar_base + 8u * (2 + 2 * STAGES + 3 * NSB);
  const uint32_t xch = bar_base + 256;          // float [2][2][128]

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nkt = (p.Lk + BN - 1) / BN;

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int i = 0; i < NSB; ++i) { mbar_init(s_full(i), 1); mbar_init(p_full(i), 8); mbar_init(pv_done(i), 1); }
    mbar_init(o_full, 1);
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

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    const bool leader = elect_one();
    const bool ev_on = PROBE && (p.dbg & 64) && p.probe && leader && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    if (leader) {
      mbar_arrive_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
      for (int a = 0; a < NATOM; ++a) tma_load_4d(sQ + a * QA_BYTES, &tmQ, q_full, 64 * a, h, q0, b);
    }
    int stage = 0; uint32_t phase = 0;
    for (int j = 0; j < nkt; ++j) {
      mbar_wait(kv_empty(stage), phase ^ 1u);
      ATTN_EV(j, 0);
      if (leader) {
        if (PROBE && (p.dbg & 16)) {
          mbar_arrive(kv_full(stage));
        } else {
          mbar_arrive_expect_tx(kv_full(stage), 2 * C::KV_BYTES);
          const uint32_t sk = sKV + stage * 2 * C::KV_BYTES, sv = sk + C::KV_BYTES;
#pragma unroll
          for (int a = 0; a < NATOM; ++a) {
            tma_load_4d(sk + a * KVA_BYTES, &tmK, kv_full(stage), 64 * a, h, j * BN, b);
            tma_load_4d(sv + a * KVA_BYTES, &tmV, kv_full(stage), 64 * a, h, j * BN, b);
          }
        }
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, BN);            // S = Q K^T : N = 64 keys
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, DPAD, true);    // O += P V  : N = DPAD, B (V) is MN-major
    const bool leader = elect_one();
    const bool ev_on = PROBE && (p.dbg & 64) && p.probe && leader && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    mbar_wait(q_full, 0);
    tc_fence_after();
    auto issue_qk = [&](int j) {
      const int stage = j % STAGES;
      mbar_wait(kv_full(stage), (uint32_t)((j / STAGES) & 1));
      tc_fence_after();
      ATTN_EV(j, 1);
      const uint32_t sk = sKV + stage * 2 * C::KV_BYTES;
      if constexpr (ONES) {     // V_j[key][d_head] = 1 (128-byte swizzled rows: 16-byte chunk index ^ (row & 7))
        const uint32_t svt = sk + C::KV_BYTES;
#pragma unroll
        for (int r = lane; r < BN; r += 32) {
          const uint32_t a = svt + (uint32_t)r * 128u + ((((uint32_t)p.d >> 3) ^ ((uint32_t)r & 7u)) << 4);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)0x3F80) : "memory");
        }
        fence_proxy_async();    // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
      }
      const uint32_t d_tmem = tmem_base + S_COL + (j % NSB) * BN;
      if (leader) {
        if (!(PROBE && (p.dbg & 2)))
#pragma unroll
        for (int kk = 0; kk < DPAD / 16; ++kk) {        // 16 head-dim columns per step: atom kk/4, +32 B inside the atom
          const uint64_t qd = umma_desc_kmajor_sw128(sQ + (kk >> 2) * QA_BYTES) + 2 * (kk & 3);
          const uint64_t kd = umma_desc_kmajor_sw128(sk + (kk >> 2) * KVA_BYTES) + 2 * (kk & 3);
          umma_bf16(d_tmem, qd, kd, idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(s_full(j % NSB));
      }
      ATTN_EV(j, 2);
    };
    for (int j = 0; j < NSB - 1 && j < nkt; ++j) issue_qk(j);
    for (int j = 0; j < nkt; ++j) {
      if (j + NSB - 1 < nkt) issue_qk(j + NSB - 1);   // its S buffer held P_{j-1}: PV_{j-1} is already queued ahead of it
      const int stage = j % STAGES, bsel = j % NSB;
      mbar_wait(p_full(bsel), (uint32_t)((j / NSB) & 1));
      tc_fence_after();
      ATTN_EV(j, 3);
      // V tile: [64 keys][DPAD cols] as NATOM atoms of 64 columns, KVA_BYTES apart (LBO); MN-major B operand
      const uint64_t vdesc = umma_desc_mnmajor_sw128(sKV + stage * 2 * C::KV_BYTES + C::KV_BYTES, KVA_BYTES);
      const uint32_t a_tmem = tmem_base + S_COL + bsel * BN;           // P_j aliases the first 32 columns of S_j
      if (leader) {
        if (!(PROBE && (p.dbg & 4)))
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)           // 16 keys per step: +8 TMEM columns of P, +16 rows (2048 B) of V
          umma_bf16_ts(tmem_base + O_COL, a_tmem + 8 * k, vdesc + 128 * k, idesc_o, (j | k) != 0 ? 1u : 0u);
        umma_commit(kv_empty(stage));
        umma_commit(pv_done(bsel));
      }
      ATTN_EV(j, 4);
    }
    if (leader) umma_commit(o_full);
  } else if (warp >= 2) {
    // ===================== softmax / correction / output =====================
    const int q = warp & 3;                      // TMEM lane quarter
    const int hc = (warp - 2) >> 2;              // column half of the key tile this thread owns
    const int rloc = q * 32 + lane;              // row inside the 128-query tile
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int row = q0 + rloc;
    const float sl2 = p.scale_log2;
    auto xch_addr = [&](int par, int half, int r) { return xch + (uint32_t)(((par * 2 + half) * 128 + r) * 4); };
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); };
    float m_ref = -INFINITY;        // reference max (raw score units) the stored P / O are relative to
    float l = 0.f;                  // this thread's partial row sum (its 32 keys per tile)
    long long stamp[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t_prev = 0;
    const bool probe_on = PROBE && !(p.dbg & 64) && p.probe != nullptr && lane == 0 && (warp == 2 || warp == 7);
    const bool ev_on = PROBE && (p.dbg & 64) && p.probe && lane == 0 && warp == 2 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    if (PROBE && probe_on) asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_prev));
    for (int j = 0; j < nkt; ++j) {
      const int bsel = j % NSB;
      ATTN_STAMP(0);
      mbar_wait(s_full(bsel), (uint32_t)((j / NSB) & 1));
      ATTN_STAMP(1);
      tc_fence_after();
      ATTN_STAMP(2);
      ATTN_EV(j, 5);
      uint32_t sv[32];
      tmem_ld32(lane_base + S_COL + bsel * BN + hc * 32, sv);
      ATTN_STAMP(3);
      tmem_ld_wait();
      ATTN_STAMP(4);
      if (j == nkt - 1 && (p.Lk & (BN - 1))) {
        const int valid = p.Lk - j * BN - hc * 32;      // keys [valid, 32) of this half are padding
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= valid) sv[i] = 0xff800000u;          // -inf
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
      // row max over both column halves: exchange through smem with the partner warp (same rows).  After this
      // barrier BOTH warps hold their scores in registers, so P may overwrite the S columns.
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch_addr(j & 1, hc, rloc)), "f"(mx) : "memory");
      ATTN_STAMP(5);
      pair_sync();
      ATTN_STAMP(6);
      float other;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(xch_addr(j & 1, hc ^ 1, rloc)) : "memory");
      mx = fmaxf(mx, other);
      const float m_new = fmaxf(m_ref, mx);
      if (j == 0) {
        m_ref = m_new;
      } else {
        const bool need = (m_new - m_ref) * sl2 > RESCALE_THRESHOLD;
        const bool any_need = __any_sync(0xffffffffu, need);
        ATTN_STAMP(7);
        if (any_need) {     // tcgen05.ld/st are warp-collective: whole warp rescales
          // Only the (rare) rescale touches O, so only it waits for PV_{j-1}: the common path never waits on the
          // P.V MMAs, and the softmax warps run back to back while the MMA issuer trails one tile behind.  Skipping
          // waits is phase-safe: PV_{j+1} cannot complete before this thread's p_full(j+1) arrival, so pv_done never
          // runs more than one phase ahead of the parity tested here.
          mbar_wait(pv_done((j - 1) % NSB), (uint32_t)(((j - 1) / NSB) & 1));
          tc_fence_after();
          const float f = need ? ex2_approx((m_ref - m_new) * sl2) : 1.0f;
          if (need) { m_ref = m_new; if constexpr (!ONES) l *= f; }
#pragma unroll
          for (int c = 0; c < DPAD / 16; ++c) {
            if ((c & 1) != hc) continue;         // the two warps of a lane quarter split the O columns
            uint32_t o[16];
            tmem_ld16(lane_base + O_COL + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st16(lane_base + O_COL + c * 16, o);
          }
          tmem_st_wait();
        }
      }
      ATTN_STAMP(8);
      const float ms = m_ref * sl2;
      uint32_t pk[16];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a0 = EX2PAIR(fmaf(__uint_as_float(sv[2 * i]), sl2, -ms));
        const float a1 = EX2PAIR(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, -ms));
        if constexpr (!ONES) sum += a0 + a1;
        pk[i] = pack_bf16x2(a0, a1);
      }
      if constexpr (!ONES) l += sum;
      ATTN_STAMP(9);
      tmem_st16(lane_base + S_COL + bsel * BN + hc * 16, pk);
      ATTN_STAMP(10);
      tmem_st_wait();
      ATTN_STAMP(11);
      tc_fence_before();
      ATTN_STAMP(12);
      __syncwarp();
      ATTN_STAMP(13);
      if (lane == 0) mbar_arrive(p_full(bsel));       // one arrival per warp: 256 same-address arrivals serialise in the smem pipe
      ATTN_STAMP(14);
      ATTN_EV(j, 6);
    }
    if (PROBE && probe_on) {
      const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      long long* dst = p.probe + ((long long)cta * 2 + (warp == 7)) * 16;
      for (int i = 0; i < 16; ++i) dst[i] = stamp[i];
    }
    // ---- epilogue: total row sum, O / l -> bf16
    float inv;
    if constexpr (ONES) {
      mbar_wait(o_full, 0);
      tc_fence_after();
      const uint32_t lsum = tmem_ld1(lane_base + O_COL + p.d);              // column d_head of O = sum_j P_j . 1
      tmem_ld_wait();
      inv = 1.0f / __uint_as_float(lsum);
    } else {
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch_addr(nkt & 1, hc, rloc)), "f"(l) : "memory");
      pair_sync();
      float l_other;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l_other) : "r"(xch_addr(nkt & 1, hc ^ 1, rloc)) : "memory");
      mbar_wait(o_full, 0);
      tc_fence_after();
      inv = 1.0f / (l + l_other);
    }
    bf16* orow = p.o + (long long)b * p.o_batch + (long long)row * p.o_row + (long long)h * p.d;
#pragma unroll
    for (int c = 0; c < DPAD / 16; ++c) {
      if ((c & 1) != hc) continue;
      uint32_t o[16];
      tmem_ld16(lane_base + O_COL + c * 16, o);
      tmem_ld_wait();
      if (row < p.Lq) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int col = c * 16 + g * 8;
          if (col < p.d) {           // d is a multiple of 8: whole 8-column groups are valid or not
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + col) = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


int g_attn_tc_variant = 0;     // test hook: 0 = auto, 1 = PROBE build where asked (scripts/micro), 3 = no ones-column row sum
int g_attn_dbg = 0;            // PROBE build only: knock-out flags (1 MUFU, 2 QK^T MMAs, 4 P.V MMAs, 16 K/V TMA), 64 = event trace

long long* g_attn_probe = nullptr;

template <int DPAD, bool PROBE = false, bool ONES = false>
static int launch_attn_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcParams& p, int B, cudaStream_t st) {
  static bool attr_set = false;
  auto kern = attn_tc_kernel<DPAD, PROBE, ONES>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::Cfg<DPAD>::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(std::string("cudaFuncSetAttribute(attn_tc): ") + cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((p.Lq + atc::BM - 1) / atc::BM, p.heads, B);
  launch_k(kern, grid, dim3(320), atc::Cfg<DPAD>::SMEM_BYTES, st, 1, tq, tk, tv, p);
  count_launch();
  return check_launch("attention_tc launch");
}

// Returns 1 if this path does not apply (caller falls back to the mma.sync kernel), 0 on success, -1 on error.
int attention_tc(const GlgAttnArgs* a, cudaStream_t st) {
  if (a->d_head > 160) return 1;
  // tensor-map constraints: 16-byte aligned bases and strides; the output is written with 16-byte stores
  if ((a->d_head % 8) || (a->o_row % 8) || (a->o_batch % 8) || ((uintptr_t)a->out & 15)) return 1;
  CUtensorMap tq, tk, tv;
  const uint64_t d = a->d_head, hd = a->heads;
  {
    const uint64_t dims[4] = {d, hd, (uint64_t)a->Lq, (uint64_t)a->B};
    const uint64_t str[3] = {d * 2, (uint64_t)a->q_row * 2, (uint64_t)a->q_batch * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)atc::BM, 1};
    if (get_tmap_bf16(&tq, a->q, 4, dims, str, box)) return -1;
  }
  {
    const uint64_t dims[4] = {d, hd, (uint64_t)a->Lk, (uint64_t)a->B};
    const uint64_t strk[3] = {d * 2, (uint64_t)a->k_row * 2, (uint64_t)a->k_batch * 2};
    const uint64_t strv[3] = {d * 2, (uint64_t)a->v_row * 2, (uint64_t)a->v_batch * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)atc::BN, 1};
    if (get_tmap_bf16(&tk, a->k, 4, dims, strk, box)) return -1;
    if (get_tmap_bf16(&tv, a->v, 4, dims, strv, box)) return -1;
  }
  AttnTcParams p;
  p.o = (bf16*)a->out; p.o_row = a->o_row; p.o_batch = a->o_batch;
  p.heads = a->heads; p.d = a->d_head; p.Lq = a->Lq; p.Lk = a->Lk;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.probe = g_attn_probe;
  p.dbg = g_attn_dbg;
  const int dpad = (a->d_head + 15) / 16 * 16;
  const bool ones = a->d_head < dpad && g_attn_tc_variant != 3;      // a spare V column carries the row sum
  switch (dpad) {
    case 16: return ones ? launch_attn_tc<16, false, true>(tq, tk, tv, p, a->B, st) : launch_attn_tc<16>(tq, tk, tv, p, a->B, st);
    case 32: return ones ? launch_attn_tc<32, false, true>(tq, tk, tv, p, a->B, st) : launch_attn_tc<32>(tq, tk, tv, p, a->B, st);
    case 48:
      if ((g_attn_probe || g_attn_dbg) && g_attn_tc_variant == 1) return launch_attn_tc<48, true>(tq, tk, tv, p, a->B, st);
      return ones ? launch_attn_tc<48, false, true>(tq, tk, tv, p, a->B, st) : launch_attn_tc<48>(tq, tk, tv, p, a->B, st);
    case 64: return ones ? launch_attn_tc<64, false, true>(tq, tk, tv, p, a->B, st) : launch_attn_tc<64>(tq, tk, tv, p, a->B, st);
    case 80: return launch_attn_tc<80>(tq, tk, tv, p, a->B, st);
    case 96: return launch_attn_tc<96>(tq, tk, tv, p, a->B, st);
    case 112: return launch_attn_tc<112>(tq, tk, tv, p, a->B, st);
    case 128: return launch_attn_tc<128>(tq, tk, tv, p, a->B, st);
    case 144: return launch_attn_tc<144>(tq, tk, tv, p, a->B, st);
    case 160: return launch_attn_tc<160>(tq, tk, tv, p, a->B, st);
  }
  return 1;
}

}  // namespace glg

// test hook: device buffer of [ctas][2 warps][8] int64 phase-clock sums for the d_head = 40 kernel (nullptr = off)
extern "C" void glg_debug_attn_probe(void* buf) { glg::g_attn_probe = reinterpret_cast<long long*>(buf); }
extern "C" void glg_debug_attn_tc_variant(int v) { glg::g_attn_tc_variant = v; }
extern "C" void glg_debug_attn_poly(int v) { glg::g_attn_dbg = v; }   // knock-out / trace flags of the PROBE instantiation (scripts/micro)
