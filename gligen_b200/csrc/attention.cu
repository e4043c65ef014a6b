// Fused (flash-style) multi-head attention for the GLIGEN transformer blocks.
//
//   O = softmax(Q K^T * scale) V      per (batch, head); online softmax in fp32; scores stay on chip.
//
// v1 data path: 64-query x 64-key tiles, 4 warps (16 query rows each), bf16 mma.sync m16n8k16 with
// ldmatrix-fed operands, K/V tiles double-buffered with cp.async.  Head dims 8..160 (multiple of 8)
// are zero-padded to a multiple of 16 in shared memory only.  Ragged key lengths (T+G grounding
// tokens, 77 text tokens) are masked to -inf in the last tile.
//
// At d_head = 40 (the 64x64 level, 88% of the attention FLOPs) the kernel is bound by the exp2
// (MUFU) rate, not by the tensor pipe, which is why this legacy-MMA path is a sound first
// version; the tcgen05/TMEM variant is the planned upgrade for d_head = 80/160.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <string>

#include "common.cuh"
#include "internal.h"
#include "../../include/gligen_b200.h"

namespace glg {

struct AttnKParams {
  const bf16* q; const bf16* k; const bf16* v; bf16* o;
  long long q_row, k_row, v_row, o_row, q_batch, k_batch, v_batch, o_batch;
  int heads, d, Lq, Lk;
  float scale_log2;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DPAD>
struct AttnCfg {
  static constexpr int BM = 64, BN = 64;
  static constexpr int LDS = DPAD + 8;                 // padded row (elements): (DPAD+8)*2 B == 16*odd (mod 128) -> conflict-free ldmatrix
  static constexpr int TILE_ELEMS = 64 * LDS;
  static constexpr int SMEM_BYTES = 5 * TILE_ELEMS * 2;  // Q + 2xK + 2xV
};

// Copy a 64-row tile of one head into shared memory: rows [row0, row0+64) of a [L, *] matrix with row
// stride `ld`, first d columns (d/8 16-byte chunks per row).  Rows >= L are zero-filled.
template <int DPAD>
__device__ __forceinline__ void load_tile(uint32_t smem_tile, const bf16* g, long long ld, int row0, int L, int d) {
  constexpr int LDS = AttnCfg<DPAD>::LDS;
  const int chunks = d >> 3;
  const int total = 64 * chunks;
  for (int i = threadIdx.x; i < total; i += 128) {
    const int r = i / chunks, c = i - r * chunks;
    const int gr = row0 + r;
    const bool ok = gr < L;
    const bf16* src = g + (ok ? (long long)gr * ld + c * 8 : 0);
    cp_async16(smem_tile + (uint32_t)(r * LDS + c * 8) * 2u, src, ok ? 16 : 0);
  }
}

template <int DPAD, bool SHORT>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnKParams p) {
  // SHORT (Lk <= 128, the 77-token text context of cross attention): K and V (<= 2 tiles) are loaded ONCE and the
  // CTA walks QT = 4 consecutive 64-row query tiles with a double-buffered Q, instead of one CTA (and one K/V
  // fetch, one launch-latency chain) per 64 query rows.
  pdl_trigger();
  pdl_wait();
  using Cfg = AttnCfg<DPAD>;
  constexpr int LDS = Cfg::LDS;
  constexpr int KSTEPS = DPAD / 16;       // k-steps of QK^T
  constexpr int DT = DPAD / 8;            // n-tiles of the output
  constexpr int NQ = SHORT ? 2 : 1;       // Q buffers
  constexpr int QT = SHORT ? 4 : 1;       // query tiles per CTA
  extern __shared__ __align__(16) uint8_t attn_smem[];
  const uint32_t sQ0 = smem_u32(attn_smem);
  const uint32_t sK = sQ0 + NQ * Cfg::TILE_ELEMS * 2;
  const uint32_t sV = sK + 2 * Cfg::TILE_ELEMS * 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int d = p.d;
  const bf16* gq = p.q + (long long)b * p.q_batch + (long long)h * d;
  const bf16* gk = p.k + (long long)b * p.k_batch + (long long)h * d;
  const bf16* gv = p.v + (long long)b * p.v_batch + (long long)h * d;

  // zero the padded columns [d, DPAD) of every tile once (cp.async never writes them)
  if (d < DPAD) {
    const int padc = DPAD - d;                   // multiple of 8
    for (int i = threadIdx.x; i < (4 + NQ) * 64 * (padc / 8); i += 128) {
      const int row = i / (padc / 8), c = i - row * (padc / 8);
      *reinterpret_cast<uint4*>(attn_smem + ((size_t)row * LDS + d + c * 8) * 2) = make_uint4(0, 0, 0, 0);
    }
  }
  const int nkt = (p.Lk + 63) / 64;
  const int qbase = blockIdx.x * 64 * QT;
  load_tile<DPAD>(sQ0, gq, p.q_row, qbase, p.Lq, d);
  load_tile<DPAD>(sK, gk, p.k_row, 0, p.Lk, d);
  load_tile<DPAD>(sV, gv, p.v_row, 0, p.Lk, d);
  if (SHORT && nkt > 1) {
    load_tile<DPAD>(sK + Cfg::TILE_ELEMS * 2, gk, p.k_row, 64, p.Lk, d);
    load_tile<DPAD>(sV + Cfg::TILE_ELEMS * 2, gv, p.v_row, 64, p.Lk, d);
  }
  cp_async_commit();

  float o_acc[DT][4];
  float m_run[2], l_run[2];
  auto reset_state = [&]() {
#pragma unroll
    for (int j = 0; j < DT; ++j) { o_acc[j][0] = o_acc[j][1] = o_acc[j][2] = o_acc[j][3] = 0.f; }
    m_run[0] = m_run[1] = -INFINITY;
    l_run[0] = l_run[1] = 0.f;
  };
  reset_state();

  const uint32_t q_lane_off = (uint32_t)((warp * 16 + (lane & 15)) * LDS + (lane >> 4) * 8) * 2u;
  // K (non-transposed) ldmatrix lane offsets: matrix m = lane/8 -> key += (m/2)*8, col += (m%2)*8
  const uint32_t k_lane_off = (uint32_t)((((lane >> 4) * 8) + (lane & 7)) * LDS + ((lane >> 3) & 1) * 8) * 2u;
  // V (transposed) ldmatrix lane offsets: matrix m = lane/8 -> key += (m%2)*8, dcol += (m/2)*8
  const uint32_t v_lane_off = (uint32_t)(((((lane >> 3) & 1) * 8) + (lane & 7)) * LDS + (lane >> 4) * 8) * 2u;

  // one 64-key tile: S = Q K^T, mask, online softmax, O += P V
  auto compute_tile = [&](uint32_t sQ, uint32_t sKb, uint32_t sVb, int kt) {
    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < KSTEPS; ++kk) {
      uint32_t a0, a1, a2, a3;
      ldsm_x4(sQ + q_lane_off + kk * 32, a0, a1, a2, a3);
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sKb + k_lane_off + (uint32_t)(jp * 16 * LDS) * 2u + kk * 32, b0, b1, b2, b3);
        mma_bf16_16816(s[2 * jp], a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(s[2 * jp + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    // ---- mask ragged tail
    if (kt == nkt - 1 && (p.Lk & 63)) {
      const int kbase = kt * 64 + 2 * (lane & 3);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = kbase + j * 8;
        if (key >= p.Lk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (key + 1 >= p.Lk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    // ---- online softmax (rows g = lane/4 and g+8)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
    const float corr0 = fast_exp2((m_run[0] - mn0) * p.scale_log2);
    const float corr1 = fast_exp2((m_run[1] - mn1) * p.scale_log2);
    m_run[0] = mn0; m_run[1] = mn1;
    const float ms0 = mn0 * p.scale_log2, ms1 = mn1 * p.scale_log2;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pa[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = fast_exp2(fmaf(s[j][0], p.scale_log2, -ms0));
      const float p1 = fast_exp2(fmaf(s[j][1], p.scale_log2, -ms0));
      const float p2 = fast_exp2(fmaf(s[j][2], p.scale_log2, -ms1));
      const float p3 = fast_exp2(fmaf(s[j][3], p.scale_log2, -ms1));
      // the row sum uses the bf16-rounded probabilities that actually enter the PV product
      const uint32_t u01 = pack_bf16x2(p0, p1), u23 = pack_bf16x2(p2, p3);
      const float2 f01 = unpack_bf16x2(u01), f23 = unpack_bf16x2(u23);
      rs0 += f01.x + f01.y;
      rs1 += f23.x + f23.y;
      pa[j][0] = u01; pa[j][1] = u23;
    }
    l_run[0] = l_run[0] * corr0 + rs0;
    l_run[1] = l_run[1] * corr1 + rs1;
#pragma unroll
    for (int j = 0; j < DT; ++j) {
      o_acc[j][0] *= corr0; o_acc[j][1] *= corr0;
      o_acc[j][2] *= corr1; o_acc[j][3] *= corr1;
    }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t a0 = pa[2 * kk][0], a1 = pa[2 * kk][1], a2 = pa[2 * kk + 1][0], a3 = pa[2 * kk + 1][1];
#pragma unroll
      for (int jp = 0; jp < DT / 2; ++jp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sVb + v_lane_off + (uint32_t)(kk * 16 * LDS + jp * 16) * 2u, b0, b1, b2, b3);
        mma_bf16_16816(o_acc[2 * jp], a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(o_acc[2 * jp + 1], a0, a1, a2, a3, b2, b3);
      }
    }
  };
  // O /= l, write bf16 rows [q0, q0 + 64)
  auto finalize = [&](int q0) {
  float l0 = l_run[0], l1 = l_run[1];
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  bf16* go = p.o + (long long)b * p.o_batch + (long long)h * d;
#pragma unroll
  for (int j = 0; j < DT; ++j) {
    const int col = j * 8 + 2 * (lane & 3);
    if (col < d) {
      if (r0 < p.Lq) *reinterpret_cast<uint32_t*>(go + (long long)r0 * p.o_row + col) = pack_bf16x2(o_acc[j][0] * inv0, o_acc[j][1] * inv0);
      if (r1 < p.Lq) *reinterpret_cast<uint32_t*>(go + (long long)r1 * p.o_row + col) = pack_bf16x2(o_acc[j][2] * inv1, o_acc[j][3] * inv1);
    }
  }
  };

  if constexpr (!SHORT) {
    for (int kt = 0; kt < nkt; ++kt) {
      const int buf = kt & 1;
      if (kt + 1 < nkt) {
        load_tile<DPAD>(sK + (buf ^ 1) * Cfg::TILE_ELEMS * 2, gk, p.k_row, (kt + 1) * 64, p.Lk, d);
        load_tile<DPAD>(sV + (buf ^ 1) * Cfg::TILE_ELEMS * 2, gv, p.v_row, (kt + 1) * 64, p.Lk, d);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      compute_tile(sQ0, sK + buf * Cfg::TILE_ELEMS * 2, sV + buf * Cfg::TILE_ELEMS * 2, kt);
      __syncthreads();
    }
    finalize(qbase);
  } else {
    for (int s = 0; s < QT; ++s) {
      const int q0 = qbase + s * 64;
      if (q0 >= p.Lq) break;
      const bool more = (s + 1 < QT) && (q0 + 64 < p.Lq);
      if (more) {
        load_tile<DPAD>(sQ0 + ((s + 1) & 1) * Cfg::TILE_ELEMS * 2, gq, p.q_row, q0 + 64, p.Lq, d);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      reset_state();
      for (int kt = 0; kt < nkt; ++kt)
        compute_tile(sQ0 + (s & 1) * Cfg::TILE_ELEMS * 2, sK + kt * Cfg::TILE_ELEMS * 2, sV + kt * Cfg::TILE_ELEMS * 2, kt);
      finalize(q0);
      __syncthreads();            // this Q buffer is refilled two iterations later
    }
  }
}

template <int DPAD, bool SHORT>
static int launch_attn2(const AttnKParams& p, int B, cudaStream_t st) {
  using Cfg = AttnCfg<DPAD>;
  static bool attr_set = false;
  auto kern = attn_fwd_kernel<DPAD, SHORT>;
  constexpr int smem = (SHORT ? 6 : 5) * Cfg::TILE_ELEMS * 2;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(std::string("cudaFuncSetAttribute(attn): ") + cudaGetErrorString(e));
    attr_set = true;
  }
  const int rows_per_cta = SHORT ? 256 : 64;
  dim3 grid((p.Lq + rows_per_cta - 1) / rows_per_cta, p.heads, B);
  launch_k(kern, dim3(grid), dim3(128), smem, st, 1, p);
  count_launch();
  return check_launch("attention launch");
}
template <int DPAD>
static int launch_attn(const AttnKParams& p, int B, cudaStream_t st) {
  return p.Lk <= 128 ? launch_attn2<DPAD, true>(p, B, st) : launch_attn2<DPAD, false>(p, B, st);
}

}  // namespace glg

namespace glg {
int attention_tc(const GlgAttnArgs* a, cudaStream_t st);         // attention_tc.cu: tcgen05 flash attention (streamed key tiles)
int attention_short_tc(const GlgAttnArgs* a, cudaStream_t st);   // attention_short_tc.cu: tcgen05, all keys in one tile (Lk <= 128)
int attention_tc2(const GlgAttnArgs* a, cudaStream_t st);        // attention_tc2.cu: tcgen05, key tiles dealt to two independent softmax warpgroups
int attention_tc3(const GlgAttnArgs* a, cudaStream_t st);        // attention_tc3.cu: tcgen05, one CTA per SM, four softmax warpgroups, one TMEM read per score
int g_attn_mode = 0;          // test hooks: 0 = auto, 1 = force the mma.sync kernel, 2 = force the streamed tcgen05 kernel,
                              // 3 = force the short-key tcgen05 kernel, 4 = force the two-warpgroup tcgen05 kernel
                              // 5 = force the four-warpgroup single-read tcgen05 kernel (each "where it applies")
}

using namespace glg;

extern "C" void glg_debug_attn_mode(int mode) { glg::g_attn_mode = mode; }

extern "C" int glg_attention(const GlgAttnArgs* a, void* stream) {
  if (!a) return set_error("glg_attention: null args");
  if (a->d_head <= 0 || a->d_head % 8 || a->d_head > 160) return set_error("glg_attention: d_head must be a multiple of 8 in [8,160]");
  if (a->Lq <= 0 || a->Lk <= 0 || a->B <= 0 || a->heads <= 0) return set_error("glg_attention: bad sizes");
  if ((a->q_row | a->k_row | a->v_row | a->q_batch | a->k_batch | a->v_batch) % 8) return set_error("glg_attention: q/k/v strides must be multiples of 8 elements");
  if ((a->o_row | a->o_batch) % 2) return set_error("glg_attention: output strides must be even");
  if (((uintptr_t)a->q | (uintptr_t)a->k | (uintptr_t)a->v) & 15) return set_error("glg_attention: q/k/v must be 16-byte aligned");
  if (a->causal) {                    // causal mask: only the short-key kernel implements it (the 77-token CLIP text encoder)
    const int rc = a->Lk <= 128 ? attention_short_tc(a, reinterpret_cast<cudaStream_t>(stream)) : 1;
    return rc <= 0 ? rc : set_error("glg_attention: causal attention needs Lk <= 128, d_head % 8 == 0 and 16-byte aligned output rows");
  }
  if ((g_attn_mode == 0 && a->Lk <= 128) || g_attn_mode == 3) {     // short key sets (the 77-token text context)
    const int rc = attention_short_tc(a, reinterpret_cast<cudaStream_t>(stream));
    if (rc <= 0) return rc;
  }
  if ((g_attn_mode == 0 && a->Lk >= 512) || g_attn_mode == 5) {     // d_head < 64 with a spare column (d_head = 40: the 64x64 level)
    const int rc = attention_tc3(a, reinterpret_cast<cudaStream_t>(stream));
    if (rc <= 0) return rc;
  }
  if ((g_attn_mode == 0 && a->Lk > 128) || g_attn_mode == 4 || g_attn_mode == 5) {
    const int rc = attention_tc2(a, reinterpret_cast<cudaStream_t>(stream));
    if (rc <= 0) return rc;
  }
  if ((g_attn_mode == 0 && a->Lk > 128) || g_attn_mode == 2 || g_attn_mode == 4 || g_attn_mode == 5) {
    const int rc = attention_tc(a, reinterpret_cast<cudaStream_t>(stream));
    if (rc <= 0) return rc;          // 0 = launched, -1 = error; 1 = not applicable -> mma.sync kernel below
  }
  AttnKParams p;
  p.q = (const bf16*)a->q; p.k = (const bf16*)a->k; p.v = (const bf16*)a->v; p.o = (bf16*)a->out;
  p.q_row = a->q_row; p.k_row = a->k_row; p.v_row = a->v_row; p.o_row = a->o_row;
  p.q_batch = a->q_batch; p.k_batch = a->k_batch; p.v_batch = a->v_batch; p.o_batch = a->o_batch;
  p.heads = a->heads; p.d = a->d_head; p.Lq = a->Lq; p.Lk = a->Lk;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int dpad = (a->d_head + 15) / 16 * 16;
  switch (dpad) {
    case 16: return launch_attn<16>(p, a->B, st);
    case 32: return launch_attn<32>(p, a->B, st);
    case 48: return launch_attn<48>(p, a->B, st);
    case 64: return launch_attn<64>(p, a->B, st);
    case 80: return launch_attn<80>(p, a->B, st);
    case 96: return launch_attn<96>(p, a->B, st);
    case 112: return launch_attn<112>(p, a->B, st);
    case 128: return launch_attn<128>(p, a->B, st);
    case 144: return launch_attn<144>(p, a->B, st);
    case 160: return launch_attn<160>(p, a->B, st);
  }
  return set_error("glg_attention: unsupported head dim");
}
