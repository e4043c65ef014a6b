// tcgen05 / TMEM flash attention for head dims <= 64, second layout: the key tiles of a 128-query tile are dealt
// alternately to TWO independent softmax warpgroups (even tiles -> warps 2..5, odd tiles -> warps 6..9).
//
// Why: profiles/r1_attention_pipeline.md shows the row-pair kernel (attention_tc.cu: two threads per query row, row
// maximum exchanged through shared memory + bar.sync every tile) bound by the softmax warps' own latency chain
// (~1650 clk per 128 x 128 score block against 1024 clk of MUFU), a third of it in the exchange.  Here
//   * one THREAD owns one query row of one warpgroup's tiles (64 keys per tile): running maximum, lazy-rescale decision
//     and the exp2 pass need no other thread - no st/ld.shared, no bar.sync, no vote-dependent hand-off per tile;
//   * each warpgroup runs its own online softmax over its half of the key tiles into its OWN accumulator O_g (with its
//     own reference maximum m_g); the halves are merged once per query tile:
//         O = (2^(m_0-m) O_0 + 2^(m_1-m) O_1) / (2^(m_0-m) l_0 + 2^(m_1-m) l_1),   m = max(m_0, m_1)
//     (the row sums l_g ride in a spare column of O_g: the MMA warp writes 1.0 into column d_head of every V tile);
//   * one S buffer per warpgroup is enough: while warpgroup g waits for P.V_j and QK^T_{j+2} (issued back to back, the
//     tensor pipe executes in order, so S_g is never overwritten before P_j has been consumed), the other warpgroup and
//     the second CTA on the SM keep the MUFU busy: four warpgroups per SM, each needing 512 MUFU-clk per tile.
//
//   warp 0 : TMA producer (Q once, then K_j / V_j tiles through a 4-stage ring)
//   warp 1 : MMA issuer    QK^T_0, QK^T_1, then per tile j: P.V_j -> O_{j&1};  QK^T_{j+2} -> S_{j&1}
//   TMEM (256 columns, two CTAs per SM): S_0 | S_1 (64 fp32 each; the bf16 P_j overwrites the first 32 columns of its S)
//                                        | O_0 | O_1 (DPAD fp32 each, DPAD <= 64)
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string>

#include "common.cuh"
#include "internal.h"
#include "../../include/gligen_b200.h"

namespace glg {

struct AttnTc2Params {
  bf16* o; long long o_row, o_batch;
  int heads, d, Lq, Lk;
  float scale_log2;
};

namespace atc2 {
constexpr int BM = 128, BN = 64, STAGES = 4;
constexpr int Q_BYTES = BM * 64 * 2;        // 16 KB (one 64-column swizzle atom: d_head <= 64)
constexpr int KV_BYTES = BN * 64 * 2;       // 8 KB
constexpr int XCH_BYTES = 2 * 128 * 4;      // m_g per row
constexpr int SMEM_BYTES = Q_BYTES + STAGES * 2 * KV_BYTES + 1024 + 256 + XCH_BYTES;
constexpr int TMEM_COLS = 256;
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units
}  // namespace atc2

__device__ __forceinline__ float ex2_approx2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA / ALU pipes (Cody-Waite split + degree-3 polynomial on [-0.5, 0.5]; relative error 1.2e-4 mean / 7.9e-4 max, below the
// bf16 rounding of P): a share of the exponentials is taken off the MUFU pipe, which bounds this kernel (16 exp2 / clk / SM).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;                 // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.05550411f, f, 0.24022651f);
  p = fmaf(p, f, 0.69314718f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// POLY: of every 8 score pairs, the last POLY pairs take their exp2 from ex2_poly instead of the MUFU
template <int DPAD, int POLY = 0>      // head dim rounded up to a multiple of 16 with one spare column for the row sums: d_head < DPAD <= 64
__global__ void __launch_bounds__(320, 2)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnTc2Params p) {
  using namespace atc2;
  constexpr int O_COL = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sKV = base + Q_BYTES;                 // per stage: K | V
  const uint32_t bar_base = sKV + STAGES * 2 * KV_BYTES;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + STAGES + s); };
  auto s_full = [&](int g) { return bar_base + 8u * (1 + 2 * STAGES + g); };
  auto p_full = [&](int g) { return bar_base + 8u * (3 + 2 * STAGES + g); };
  const uint32_t o_full = bar_base + 8u * (5 + 2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (6 + 2 * STAGES);
  const uint32_t xch = bar_base + 256;                  // float [2][128]: reference maxima of the two warpgroups

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nkt = (p.Lk + BN - 1) / BN;

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(s_full(g), 1); mbar_init(p_full(g), 4); }
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
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    if (leader) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_4d(sQ, &tmQ, q_full, 0, h, q0, b);
    }
    int stage = 0; uint32_t phase = 0;
    for (int j = 0; j < nkt; ++j) {
      mbar_wait(kv_empty(stage), phase ^ 1u);
      if (leader) {
        mbar_arrive_expect_tx(kv_full(stage), 2 * KV_BYTES);
        const uint32_t sk = sKV + stage * 2 * KV_BYTES;
        tma_load_4d(sk, &tmK, kv_full(stage), 0, h, j * BN, b);
        tma_load_4d(sk + KV_BYTES, &tmV, kv_full(stage), 0, h, j * BN, b);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, BN);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, DPAD, true);    // B (V) is MN-major
    const bool leader = elect_one();
    mbar_wait(q_full, 0);
    tc_fence_after();
    auto issue_qk = [&](int j) {
      const int stage = j % STAGES;
      mbar_wait(kv_full(stage), (uint32_t)((j / STAGES) & 1));
      tc_fence_after();
      const uint32_t sk = sKV + stage * 2 * KV_BYTES;
      {   // V_j[key][d_head] = 1 (128-byte swizzled rows: 16-byte chunk index ^ (row & 7)): row sums come out of P.V
        const uint32_t svt = sk + KV_BYTES;
#pragma unroll
        for (int r = lane; r < BN; r += 32) {
          const uint32_t a = svt + (uint32_t)r * 128u + ((((uint32_t)p.d >> 3) ^ ((uint32_t)r & 7u)) << 4) + (((uint32_t)p.d & 7u) << 1);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)0x3F80) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
      }
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < DPAD / 16; ++kk)
          umma_bf16(tmem_base + (j & 1) * BN, umma_desc_kmajor_sw128(sQ) + 2 * kk, umma_desc_kmajor_sw128(sk) + 2 * kk, idesc_s, kk != 0 ? 1u : 0u);
        umma_commit(s_full(j & 1));
      }
    };
    for (int j = 0; j < 2 && j < nkt; ++j) issue_qk(j);
    for (int j = 0; j < nkt; ++j) {
      const int g = j & 1, stage = j % STAGES;
      mbar_wait(p_full(g), (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      const uint64_t vdesc = umma_desc_mnmajor_sw128(sKV + stage * 2 * KV_BYTES + KV_BYTES, KV_BYTES);
      if (leader) {
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)           // 16 keys per step: +8 TMEM columns of P, +16 rows (2048 B) of V
          umma_bf16_ts(tmem_base + O_COL + g * DPAD, tmem_base + g * BN + 8 * k, vdesc + 128 * k, idesc_o, (j >= 2 || k != 0) ? 1u : 0u);
        umma_commit(kv_empty(stage));
      }
      if (j + 2 < nkt) issue_qk(j + 2);             // overwrites S_g: queued behind P.V_j, the tensor pipe executes in order
    }
    if (leader) umma_commit(o_full);
  } else {
    // ===================== softmax warpgroups =====================
    const int g = (warp - 2) >> 2;               // warpgroup: owns key tiles j with (j & 1) == g
    const int q = warp & 3;                      // TMEM lane quarter
    const int rloc = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t s_addr = lane_base + g * BN;
    const uint32_t o_addr = lane_base + O_COL + g * DPAD;
    const int row = q0 + rloc;
    const float sl2 = p.scale_log2;
    float m_ref = -INFINITY;
    int it = 0;
    for (int j = g; j < nkt; j += 2, ++it) {
      mbar_wait(s_full(g), (uint32_t)(it & 1));
      tc_fence_after();
      const int valid = p.Lk - j * BN;           // < 64 only in the last tile: keys [valid, 64) are padding
      // ---- pass 1: row maximum of this tile: all 64 scores with ONE TMEM round trip (the registers die right after)
      float mx = -INFINITY;
      {
        uint32_t lo[32], hi[32];
        tmem_ld32(s_addr, lo);
        tmem_ld32(s_addr + 32, hi);
        tmem_ld_wait();
        if (valid < BN) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i >= valid) lo[i] = 0xff800000u;
            if (32 + i >= valid) hi[i] = 0xff800000u;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          mx = fmaxf(mx, fmaxf(__uint_as_float(lo[i]), __uint_as_float(lo[i + 1])));
          mx = fmaxf(mx, fmaxf(__uint_as_float(hi[i]), __uint_as_float(hi[i + 1])));
        }
      }
      // pass 2 re-reads the scores in 16-column pieces, each requested one piece ahead of its use
      uint32_t cur[16];
      tmem_ld16(s_addr, cur);
      const float m_new = fmaxf(m_ref, mx);
      if (it == 0) {
        m_ref = m_new;
      } else {
        const bool need = (m_new - m_ref) * sl2 > RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, need)) {     // tcgen05.ld / st are warp-collective: the whole warp rescales its rows
          // s_full(g) of this tile was committed after QK^T_j, which was issued after P.V_{j-2}: O_g is up to date
          const float f = need ? ex2_approx2((m_ref - m_new) * sl2) : 1.0f;
          if (need) m_ref = m_new;
#pragma unroll
          for (int c = 0; c < DPAD / 16; ++c) {
            uint32_t o[16];
            tmem_ld16(o_addr + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st16(o_addr + c * 16, o);
          }
          tmem_st_wait();
        }
      }
      // ---- pass 2: P = exp2((S - m_ref) * scale) -> bf16, written over the first 32 columns of S.  P piece k lands in
      //      columns [8k, 8k+8), which only cover score pieces <= k - all already in registers when it is written
      const float ms = m_ref * sl2;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        tmem_ld_wait();
        uint32_t nxt[16];
        if (k < 3) tmem_ld16(s_addr + (k + 1) * 16, nxt);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x0 = fmaf(__uint_as_float(cur[2 * i]), sl2, -ms), x1 = fmaf(__uint_as_float(cur[2 * i + 1]), sl2, -ms);
          float a0 = (i >= 8 - POLY) ? ex2_poly(x0) : ex2_approx2(x0);
          float a1 = (i >= 8 - POLY) ? ex2_poly(x1) : ex2_approx2(x1);
          if (valid < BN) {
            if (k * 16 + 2 * i >= valid) a0 = 0.f;
            if (k * 16 + 2 * i + 1 >= valid) a1 = 0.f;
          }
          pk[i] = pack_bf16x2(a0, a1);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"r"(s_addr + k * 8), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
        if (k < 3) {
#pragma unroll
          for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(g));
    }
    // ---- merge the two warpgroups' partial results and write the rows
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch + (uint32_t)((g * 128 + rloc) * 4)), "f"(m_ref) : "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float m_other;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(m_other) : "r"(xch + (uint32_t)(((g ^ 1) * 128 + rloc) * 4)) : "memory");
    const float m0 = g == 0 ? m_ref : m_other, m1 = g == 0 ? m_other : m_ref;
    const float m = fmaxf(m0, m1);
    const bool has1 = nkt > 1;                   // warpgroup 1 saw no tile: O_1 was never written
    const float f0 = ex2_approx2((m0 - m) * sl2);
    const float f1 = has1 ? ex2_approx2((m1 - m) * sl2) : 0.f;
    mbar_wait(o_full, 0);
    tc_fence_after();
    const uint32_t o0 = lane_base + O_COL, o1 = lane_base + O_COL + DPAD;
    const uint32_t l0u = tmem_ld1(o0 + p.d);                    // column d_head of O_g = sum_j P_j . 1
    const uint32_t l1u = has1 ? tmem_ld1(o1 + p.d) : 0u;
    tmem_ld_wait();
    float l = f0 * __uint_as_float(l0u);
    if (has1) l = fmaf(f1, __uint_as_float(l1u), l);
    const float w0 = f0 / l, w1 = f1 / l;
    bf16* orow = p.o + (long long)b * p.o_batch + (long long)row * p.o_row + (long long)h * p.d;
    // warpgroup 0 writes columns [0, 32), warpgroup 1 the rest
    constexpr int NCH = DPAD / 16;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if ((c < 2 ? 0 : 1) != g) continue;
      uint32_t a[16], bb[16];
      tmem_ld16(o0 + c * 16, a);
      if (has1) tmem_ld16(o1 + c * 16, bb);
      tmem_ld_wait();
      if (row < p.Lq) {
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          const int col = c * 16 + gg * 8;
          if (col < p.d) {           // d is a multiple of 8: whole 8-column groups are valid or not
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[i] = w0 * __uint_as_float(a[gg * 8 + i]);
              if (has1) v[i] = fmaf(w1, __uint_as_float(bb[gg * 8 + i]), v[i]);
            }
            uint4 u;
            u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
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
    tmem_dealloc(tmem_base, atc2::TMEM_COLS);
  }
}

int g_attn_tc2_poly = -1;      // -1: GLG_ATTN_POLY env (default 0); pairs of 8 whose exp2 runs on the FMA pipe (0, 1, 2, 3)

template <int DPAD, int POLY>
static int launch_attn_tc2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTc2Params& p, int B, cudaStream_t st) {
  static bool attr_set = false;
  auto kern = attn_tc2_kernel<DPAD, POLY>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, atc2::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(std::string("cudaFuncSetAttribute(attn_tc2): ") + cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((p.Lq + atc2::BM - 1) / atc2::BM, p.heads, B);
  launch_k(kern, grid, dim3(320), atc2::SMEM_BYTES, st, 1, tq, tk, tv, p);
  count_launch();
  return check_launch("attention_tc2 launch");
}

// Returns 1 if this path does not apply (caller falls back), 0 on success, -1 on error.
// Applies to d_head < 64 with a spare column up to the next multiple of 16 (d_head = 40: DPAD = 48), i.e. d_head % 16 != 0.
int attention_tc2(const GlgAttnArgs* a, cudaStream_t st) {
  const int dpad = (a->d_head + 15) / 16 * 16;
  if (dpad > 64 || dpad == a->d_head) return 1;
  if ((a->d_head % 8) || (a->o_row % 8) || (a->o_batch % 8) || ((uintptr_t)a->out & 15)) return 1;
  CUtensorMap tq, tk, tv;
  const uint64_t d = a->d_head, hd = a->heads;
  {
    const uint64_t dims[4] = {d, hd, (uint64_t)a->Lq, (uint64_t)a->B};
    const uint64_t str[3] = {d * 2, (uint64_t)a->q_row * 2, (uint64_t)a->q_batch * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)atc2::BM, 1};
    if (get_tmap_bf16(&tq, a->q, 4, dims, str, box)) return -1;
  }
  {
    const uint64_t dims[4] = {d, hd, (uint64_t)a->Lk, (uint64_t)a->B};
    const uint64_t strk[3] = {d * 2, (uint64_t)a->k_row * 2, (uint64_t)a->k_batch * 2};
    const uint64_t strv[3] = {d * 2, (uint64_t)a->v_row * 2, (uint64_t)a->v_batch * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)atc2::BN, 1};
    if (get_tmap_bf16(&tk, a->k, 4, dims, strk, box)) return -1;
    if (get_tmap_bf16(&tv, a->v, 4, dims, strv, box)) return -1;
  }
  AttnTc2Params p;
  p.o = (bf16*)a->out; p.o_row = a->o_row; p.o_batch = a->o_batch;
  p.heads = a->heads; p.d = a->d_head; p.Lq = a->Lq; p.Lk = a->Lk;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  if (g_attn_tc2_poly < 0) { const char* e = getenv("GLG_ATTN_POLY"); g_attn_tc2_poly = e ? atoi(e) : 0; }
  switch (dpad) {
    case 16: return launch_attn_tc2<16, 0>(tq, tk, tv, p, a->B, st);
    case 32: return launch_attn_tc2<32, 0>(tq, tk, tv, p, a->B, st);
    case 48:
      switch (g_attn_tc2_poly) {
        case 1: return launch_attn_tc2<48, 1>(tq, tk, tv, p, a->B, st);
        case 2: return launch_attn_tc2<48, 2>(tq, tk, tv, p, a->B, st);
        case 3: return launch_attn_tc2<48, 3>(tq, tk, tv, p, a->B, st);
        default: return launch_attn_tc2<48, 0>(tq, tk, tv, p, a->B, st);
      }
    case 64: return launch_attn_tc2<64, 0>(tq, tk, tv, p, a->B, st);
  }
  return 1;
}

}  // namespace glg

extern "C" void glg_debug_attn_poly_share(int pairs_of_8) { glg::g_attn_tc2_poly = pairs_of_8; }
