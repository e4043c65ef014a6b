// tcgen05 / TMEM attention for SHORT key sets (Lk <= 128): the cross-attention to the 77-token text context
// (CrossAttention.forward, attention.py:127-149; 16 layers per UNet pass, Lq = 4096 / 1024 / 256 / 64 queries).
//
// All keys fit one tile, so there is no online softmax: per 128-query tile
//     S = Q K^T  (one MMA group, N = keys rounded up to 16)  ->  P = exp2((S - rowmax) * scale)  ->  O = P V
// K and V of the (batch, head) stay resident in shared memory for the CTA's lifetime; the CTA walks a range of query
// tiles with two tiles in flight (two Q stages in smem, two S/P/O buffers in TMEM, two softmax warpgroups):
//   warp 0      : TMA producer  (K, V once; then Q tiles through a 2-stage ring)
//   warp 1      : MMA issuer    QK^T(i) is issued before P.V(i-1), so the tensor pipe never waits for a softmax
//   warps 2..5  : softmax + epilogue of even tiles   (one thread per query row: no cross-thread reduction at all)
//   warps 6..9  : softmax + epilogue of odd tiles
// TMEM per buffer: S (NKP fp32 columns; the bf16 P overwrites its first NKP/2 columns chunk by chunk, always behind
// the columns already read) | O (DPAD fp32 columns).  The 77-key context (NKP = 80) needs 2 x (80 + DPAD) columns:
// 256 for d_head = 40 (two CTAs per SM), <= 480 up to d_head = 160; shapes that would need more than 512 columns are
// left to the caller's other kernels.
//
// The floor is the MUFU exp2 rate (16 / clk / SM): Lq * NKP exponentials per (batch, head).
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string>

#include "common.cuh"
#include "internal.h"
#include "../../include/gligen_b200.h"

namespace glg {

struct AttnShortParams {
  bf16* o; long long o_row, o_batch;
  int heads, d, Lq, Lk, nkp;          // nkp = Lk rounded up to a multiple of 16
  int causal;                         // query row r sees keys [0, r] only (the CLIP text encoder's causal mask)
  int tiles_per_cta, n_qtiles;
  int s_cols;                         // TMEM columns reserved for S in each buffer (= nkp)
  int tmem_cols;                      // allocation: 256 or 512
  float scale_log2;
};

namespace ast {
constexpr int BM = 128;
constexpr int QA_BYTES = BM * 64 * 2;       // one 64-column atom of a Q tile
template <int DPAD> struct Cfg {
  static constexpr int NATOM = (DPAD + 63) / 64;
  static constexpr int Q_BYTES = NATOM * QA_BYTES;
  static int smem_bytes(int nkp) { return 2 * Q_BYTES + 2 * NATOM * nkp * 128 + 1024 + 256; }
};
}  // namespace ast

__device__ __forceinline__ float ex2_approx_s(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DPAD>
__global__ void __launch_bounds__(320, DPAD <= 48 ? 2 : 1)
attn_short_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnShortParams p) {
  using namespace ast;
  using C = Cfg<DPAD>;
  constexpr int NATOM = C::NATOM;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t katom = (uint32_t)p.nkp * 128u;                 // bytes of one 64-column atom of K (or V)
  const uint32_t sQ = base;                                      // [2][NATOM][128 x 128 B]
  const uint32_t sK = base + 2 * C::Q_BYTES;                     // [NATOM][nkp x 128 B]
  const uint32_t sV = sK + NATOM * katom;
  const uint32_t bar_base = sV + NATOM * katom;                  // nkp % 16 == 0 -> 2 KB granules: 8-byte aligned
  const uint32_t kv_full = bar_base;
  auto q_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto q_empty = [&](int s) { return bar_base + 8u * (3 + s); };
  auto s_full = [&](int s) { return bar_base + 8u * (5 + s); };
  auto p_full = [&](int s) { return bar_base + 8u * (7 + s); };
  auto o_full = [&](int s) { return bar_base + 8u * (9 + s); };
  auto o_done = [&](int s) { return bar_base + 8u * (11 + s); };
  const uint32_t tmem_slot = bar_base + 8u * 13;

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int t0 = blockIdx.x * p.tiles_per_cta;
  const int t1 = min(p.n_qtiles, t0 + p.tiles_per_cta);
  const int ntile = t1 - t0;
  const int buf_cols = p.s_cols + DPAD;                          // TMEM columns per in-flight tile

  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); mbar_init(s_full(s), 1);
      mbar_init(p_full(s), 4); mbar_init(o_full(s), 1); mbar_init(o_done(s), 4);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
    __syncwarp();
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
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
      mbar_arrive_expect_tx(kv_full, 2u * NATOM * katom);
#pragma unroll
      for (int a = 0; a < NATOM; ++a) {
        tma_load_4d(sK + a * katom, &tmK, kv_full, 64 * a, h, 0, b);
        tma_load_4d(sV + a * katom, &tmV, kv_full, 64 * a, h, 0, b);
      }
    }
    for (int i = 0; i < ntile; ++i) {
      const int s = i & 1;
      mbar_wait(q_empty(s), (uint32_t)(((i >> 1) & 1) ^ 1));
      if (leader) {
        mbar_arrive_expect_tx(q_full(s), C::Q_BYTES);
#pragma unroll
        for (int a = 0; a < NATOM; ++a) tma_load_4d(sQ + s * C::Q_BYTES + a * QA_BYTES, &tmQ, q_full(s), 64 * a, h, (t0 + i) * BM, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = umma_idesc_bf16(128, p.nkp);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, DPAD, true);          // B (V) is MN-major
    const bool leader = elect_one();
    mbar_wait(kv_full, 0);
    tc_fence_after();
    auto issue_pv = [&](int i) {
      const int s = i & 1;
      mbar_wait(p_full(s), (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const uint32_t tb = tmem_base + s * buf_cols;
      const uint64_t vdesc = umma_desc_mnmajor_sw128(sV, katom);
      if (leader) {
        for (int k = 0; k < p.nkp / 16; ++k)          // 16 keys per step: +8 TMEM columns of P, +16 rows (2048 B) of V
          umma_bf16_ts(tb + p.s_cols, tb + 8 * k, vdesc + 128 * k, idesc_o, k != 0 ? 1u : 0u);
        umma_commit(o_full(s));
      }
    };
    for (int i = 0; i < ntile; ++i) {
      const int s = i & 1;
      mbar_wait(q_full(s), (uint32_t)((i >> 1) & 1));
      mbar_wait(o_done(s), (uint32_t)(((i >> 1) & 1) ^ 1));       // TMEM buffer s: tile i-2 has been read out
      tc_fence_after();
      const uint32_t tb = tmem_base + s * buf_cols;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < DPAD / 16; ++kk) {        // 16 head-dim columns per step: atom kk/4, +32 B inside the atom
          const uint64_t qd = umma_desc_kmajor_sw128(sQ + s * C::Q_BYTES + (kk >> 2) * QA_BYTES) + 2 * (kk & 3);
          const uint64_t kd = umma_desc_kmajor_sw128(sK + (kk >> 2) * katom) + 2 * (kk & 3);
          umma_bf16(tb, qd, kd, idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(s_full(s));
        umma_commit(q_empty(s));
      }
      if (i > 0) issue_pv(i - 1);
    }
    if (ntile > 0) issue_pv(ntile - 1);
  } else {
    // ===================== softmax + epilogue: warpgroup g owns tiles of parity g =====================
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;                                    // TMEM lane quarter of this warp
    const int rloc = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + g * buf_cols;
    const float sl2 = p.scale_log2;
    const int nch = p.nkp >> 4;
    for (int i = g; i < ntile; i += 2) {
      const uint32_t par = (uint32_t)((i >> 1) & 1);
      mbar_wait(s_full(g), par);
      tc_fence_after();
      // keys this row may see: all Lk, or - causal - the first (row + 1)
      const int lim = p.causal ? min(p.Lk, (t0 + i) * BM + rloc + 1) : p.Lk;
      // pass 1: row maximum (scores are re-read from TMEM in pass 2: 80-128 live registers would cost the 2nd CTA/SM)
      float mx = -INFINITY;
      for (int c = 0; c < nch; ++c) {
        uint32_t sv[16];
        tmem_ld16(lane_base + c * 16, sv);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c * 16 + j < lim) mx = fmaxf(mx, __uint_as_float(sv[j]));
      }
      const float ms = mx * sl2;
      float l = 0.f;
      for (int c = 0; c < nch; ++c) {
        uint32_t sv[16];
        tmem_ld16(lane_base + c * 16, sv);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a0 = ex2_approx_s(fmaf(__uint_as_float(sv[2 * j]), sl2, -ms));
          float a1 = ex2_approx_s(fmaf(__uint_as_float(sv[2 * j + 1]), sl2, -ms));
          if (c * 16 + 2 * j >= lim) a0 = 0.f;                  // padding keys (K rows zero-filled by TMA) / masked keys
          if (c * 16 + 2 * j + 1 >= lim) a1 = 0.f;
          l += a0 + a1;
          pk[j] = pack_bf16x2(a0, a1);
        }
        // P chunk c -> columns [8c, 8c+8): inside score chunks <= c, all of which this thread has already consumed
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"r"(lane_base + c * 8), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(g));
      // ---- epilogue of the same tile: O / l -> bf16
      const float inv = 1.0f / l;
      const int row = (t0 + i) * BM + rloc;
      bf16* orow = p.o + (long long)b * p.o_batch + (long long)row * p.o_row + (long long)h * p.d;
      mbar_wait(o_full(g), par);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < DPAD / 16; ++c) {
        uint32_t o[16];
        tmem_ld16(lane_base + p.s_cols + c * 16, o);
        tmem_ld_wait();
        if (row < p.Lq) {
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const int col = c * 16 + gg * 8;
            if (col < p.d) {           // d is a multiple of 8: whole 8-column groups are valid or not
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(o[gg * 8 + 0]) * inv, __uint_as_float(o[gg * 8 + 1]) * inv);
              u.y = pack_bf16x2(__uint_as_float(o[gg * 8 + 2]) * inv, __uint_as_float(o[gg * 8 + 3]) * inv);
              u.z = pack_bf16x2(__uint_as_float(o[gg * 8 + 4]) * inv, __uint_as_float(o[gg * 8 + 5]) * inv);
              u.w = pack_bf16x2(__uint_as_float(o[gg * 8 + 6]) * inv, __uint_as_float(o[gg * 8 + 7]) * inv);
              *reinterpret_cast<uint4*>(orow + col) = u;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_done(g));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

template <int DPAD>
static int launch_attn_short(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, AttnShortParams& p, int B, cudaStream_t st) {
  using C = ast::Cfg<DPAD>;
  p.s_cols = p.nkp;
  const int need = 2 * (p.s_cols + DPAD);
  if (need > 512) return 1;                                // not applicable (caller falls back)
  p.tmem_cols = need <= 256 ? 256 : 512;
  static int attr_set = 0;
  auto kern = attn_short_tc_kernel<DPAD>;
  const int smem = C::smem_bytes(p.nkp);
  if (attr_set < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(std::string("cudaFuncSetAttribute(attn_short_tc): ") + cudaGetErrorString(e));
    attr_set = smem;
  }
  // about two CTAs per SM over the whole grid; every CTA re-loads K/V (a few KB), so finer is nearly free
  int qsplit = (2 * num_sms() + p.heads * B - 1) / (p.heads * B);
  if (qsplit > p.n_qtiles) qsplit = p.n_qtiles;
  if (qsplit < 1) qsplit = 1;
  p.tiles_per_cta = (p.n_qtiles + qsplit - 1) / qsplit;
  qsplit = (p.n_qtiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  dim3 grid(qsplit, p.heads, B);
  launch_k(kern, grid, dim3(320), (size_t)smem, st, 1, tq, tk, tv, p);
  count_launch();
  return check_launch("attention_short_tc launch");
}

// Returns 1 if this path does not apply (caller falls back), 0 on success, -1 on error.
int attention_short_tc(const GlgAttnArgs* a, cudaStream_t st) {
  if (a->Lk > 128 || a->d_head > 160) return 1;
  if ((a->d_head % 8) || (a->o_row % 8) || (a->o_batch % 8) || ((uintptr_t)a->out & 15)) return 1;
  AttnShortParams p;
  p.o = (bf16*)a->out; p.o_row = a->o_row; p.o_batch = a->o_batch;
  p.heads = a->heads; p.d = a->d_head; p.Lq = a->Lq; p.Lk = a->Lk;
  p.nkp = (a->Lk + 15) / 16 * 16;
  p.causal = a->causal;
  p.n_qtiles = (a->Lq + ast::BM - 1) / ast::BM;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.tiles_per_cta = 1; p.s_cols = 0;
  CUtensorMap tq, tk, tv;
  const uint64_t d = a->d_head, hd = a->heads;
  {
    const uint64_t dims[4] = {d, hd, (uint64_t)a->Lq, (uint64_t)a->B};
    const uint64_t str[3] = {d * 2, (uint64_t)a->q_row * 2, (uint64_t)a->q_batch * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)ast::BM, 1};
    if (get_tmap_bf16(&tq, a->q, 4, dims, str, box)) return -1;
  }
  {
    const uint64_t dims[4] = {d, hd, (uint64_t)a->Lk, (uint64_t)a->B};
    const uint64_t strk[3] = {d * 2, (uint64_t)a->k_row * 2, (uint64_t)a->k_batch * 2};
    const uint64_t strv[3] = {d * 2, (uint64_t)a->v_row * 2, (uint64_t)a->v_batch * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)p.nkp, 1};
    if (get_tmap_bf16(&tk, a->k, 4, dims, strk, box)) return -1;
    if (get_tmap_bf16(&tv, a->v, 4, dims, strv, box)) return -1;
  }
  const int dpad = (a->d_head + 15) / 16 * 16;
  switch (dpad) {
    case 16: return launch_attn_short<16>(tq, tk, tv, p, a->B, st);
    case 32: return launch_attn_short<32>(tq, tk, tv, p, a->B, st);
    case 48: return launch_attn_short<48>(tq, tk, tv, p, a->B, st);
    case 64: return launch_attn_short<64>(tq, tk, tv, p, a->B, st);
    case 80: return launch_attn_short<80>(tq, tk, tv, p, a->B, st);
    case 96: return launch_attn_short<96>(tq, tk, tv, p, a->B, st);
    case 112: return launch_attn_short<112>(tq, tk, tv, p, a->B, st);
    case 128: return launch_attn_short<128>(tq, tk, tv, p, a->B, st);
    case 144: return launch_attn_short<144>(tq, tk, tv, p, a->B, st);
    case 160: return launch_attn_short<160>(tq, tk, tv, p, a->B, st);
  }
  return 1;
}

}  // namespace glg
