// tcgen05 / TMEM / TMA GEMM and implicit-GEMM 3x3 convolution for sm_100a.
//
//   out[M,N] = epilogue( A[M,K] . W[N,K]^T )           bf16 operands, fp32 accumulation in TMEM
//
// One persistent CTA per SM, 10 warps, warp-specialised:
//   warp 0 lane 0 : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1 lane 0 : MMA issuer    (tcgen05.mma cta_group::1, M=128, N=BN, K=16 per instruction)
//   warps 2..9    : epilogue      (tcgen05.ld 32x32b -> registers -> bias/act/gate/residual/GEGLU -> global);
//                   two warps per TMEM lane quarter, each taking half of the tile's column chunks
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.
//
// conv_mode: the A operand is gathered by a 4-D tiled tensor map over the NHWC activation
// (C, W, H, B); for tap (dy,dx) the box origin is shifted by (dx-1, dy-1) and TMA's out-of-bounds
// zero fill implements the padding, so a 3x3 convolution is 9*Cin/64 K-steps of the same pipeline
// with no im2col buffer.  A 128-row M tile is 128/W image rows (or 128/(H*W) whole images).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <unordered_map>
#include <string>

#include "common.cuh"
#include "internal.h"
#include "../../include/gligen_b200.h"

namespace glg {

struct GemmKParams {
  int M, N, num_kb, kb_per_tap;
  int tiles_m, tiles_n;
  int conv, HW, Wd;
  void* out; long long ldc; int out_fp32;
  const float* bias; const float* rowbias; long long ld_rowbias; int rows_per_batch;
  int act; const float* gate; const bf16* residual; long long ldr;
  // LayerNorm fold (consumer side): per-row partial (sum, sumsq) of A over K, column sums of the weights
  const float* ln_stats; int ln_slots; const float* ln_colsum; float ln_eps; float inv_k;
  // producer side: per-row partial (sum, sumsq) of the bf16 values this GEMM stores
  float* stats_out; int stats_slots;
  // batch-strided output rows: address = (row / orpb) * obs + (row % orpb) * ldc   (orpb == 0: uniform rows)
  int orpb; long long obs;
};

template <int BN> struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 160) ? 5 : (BN == 128) ? 6 : 8;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void epi_store_bf16(bf16* dst, const float (&v)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    d4[i] = u;
  }
}

template <int BN, bool GEGLU>
__global__ void __launch_bounds__(320, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmKParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); }
    fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int total_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
      int b0 = 0, y0 = 0;
      if (p.conv) {
        const int p0 = m_blk * 128;
        b0 = p0 / p.HW;
        y0 = (p0 - b0 * p.HW) / p.Wd;
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
        const uint32_t a_dst = base + stage * Cfg::STAGE_BYTES;
        const uint32_t b_dst = a_dst + Cfg::A_BYTES;
        if (p.conv) {
          const int tap = kb / p.kb_per_tap;
          const int cb = kb - tap * p.kb_per_tap;
          const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
          tma_load_4d(a_dst, &tmA, full_bar(stage), cb * 64, dx, y0 + dy, b0);
          tma_load_2d(b_dst, &tmB, full_bar(stage), cb * 64, tap * p.N + n_blk * BN);
        } else {
          tma_load_2d(a_dst, &tmA, full_bar(stage), kb * 64, m_blk * 128);
          tma_load_2d(b_dst, &tmB, full_bar(stage), kb * 64, n_blk * BN);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = base + stage * Cfg::STAGE_BYTES;
        const uint64_t adesc = umma_desc_kmajor_sw128(a_addr);
        const uint64_t bdesc = umma_desc_kmajor_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // 4 x K=16 inside one 64-wide (128 B) swizzle atom: +32 B per step
          umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(empty_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tfull_bar(acc));
      acc ^= 1; if (acc == 0) acc_phase ^= 1u;
    }
  } else if (warp >= 2) {
    // ===================== epilogue =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;             // which half of the column chunks this warp handles
    int acc = 0; uint32_t acc_phase = 0;
    const float gate = p.gate ? __ldg(p.gate) : 1.0f;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * 128 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      // LayerNorm fold: y = rstd * (acc - mu * colsum[n]) + bias'[n]   (gamma folded into W, beta into bias')
      float ln_mu = 0.f, ln_rstd = 1.f;
      if (p.ln_stats && row_ok) {
        const float2* sp = reinterpret_cast<const float2*>(p.ln_stats) + (size_t)row * p.ln_slots;
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < p.ln_slots; ++i) { const float2 t = __ldg(sp + i); s1 += t.x; s2 += t.y; }   // fixed order
        ln_mu = s1 * p.inv_k;
        ln_rstd = rsqrtf(fmaxf(s2 * p.inv_k - ln_mu * ln_mu, 0.f) + p.ln_eps);
      }
      const size_t out_off = p.orpb ? (size_t)(row / p.orpb) * p.obs + (size_t)(row % p.orpb) * p.ldc : (size_t)row * p.ldc;
      float st_sum = 0.f, st_sq = 0.f;
      if constexpr (GEGLU) {
        constexpr int HALF = BN / 2;
        constexpr int NCH = HALF / 32;
#pragma unroll 1
        for (int c = half ? (NCH + 1) / 2 : 0; c < (half ? NCH : (NCH + 1) / 2); ++c) {
          uint32_t rx[32], rg[32];
          tmem_ld32(taddr + c * 32, rx);
          tmem_ld32(taddr + HALF + c * 32, rg);
          tmem_ld_wait();
          float v[32];
          const float* bx = p.bias + (size_t)n_blk * BN + c * 32;
          const float* bg = bx + HALF;
          if (p.ln_stats) {
            const float* sx = p.ln_colsum + (size_t)n_blk * BN + c * 32;
            const float* sg = sx + HALF;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 ux = __ldg(reinterpret_cast<const float4*>(sx + j));
              const float4 ug = __ldg(reinterpret_cast<const float4*>(sg + j));
              rx[j + 0] = __float_as_uint(ln_rstd * (__uint_as_float(rx[j + 0]) - ln_mu * ux.x));
              rx[j + 1] = __float_as_uint(ln_rstd * (__uint_as_float(rx[j + 1]) - ln_mu * ux.y));
              rx[j + 2] = __float_as_uint(ln_rstd * (__uint_as_float(rx[j + 2]) - ln_mu * ux.z));
              rx[j + 3] = __float_as_uint(ln_rstd * (__uint_as_float(rx[j + 3]) - ln_mu * ux.w));
              rg[j + 0] = __float_as_uint(ln_rstd * (__uint_as_float(rg[j + 0]) - ln_mu * ug.x));
              rg[j + 1] = __float_as_uint(ln_rstd * (__uint_as_float(rg[j + 1]) - ln_mu * ug.y));
              rg[j + 2] = __float_as_uint(ln_rstd * (__uint_as_float(rg[j + 2]) - ln_mu * ug.z));
              rg[j + 3] = __float_as_uint(ln_rstd * (__uint_as_float(rg[j + 3]) - ln_mu * ug.w));
            }
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 tx = __ldg(reinterpret_cast<const float4*>(bx + j));
            const float4 tg = __ldg(reinterpret_cast<const float4*>(bg + j));
            v[j + 0] = (__uint_as_float(rx[j + 0]) + tx.x) * gelu_erf_f(__uint_as_float(rg[j + 0]) + tg.x);
            v[j + 1] = (__uint_as_float(rx[j + 1]) + tx.y) * gelu_erf_f(__uint_as_float(rg[j + 1]) + tg.y);
            v[j + 2] = (__uint_as_float(rx[j + 2]) + tx.z) * gelu_erf_f(__uint_as_float(rg[j + 2]) + tg.z);
            v[j + 3] = (__uint_as_float(rx[j + 3]) + tx.w) * gelu_erf_f(__uint_as_float(rg[j + 3]) + tg.w);
          }
          if (row_ok) epi_store_bf16(reinterpret_cast<bf16*>(p.out) + out_off + (size_t)n_blk * HALF + c * 32, v);
        }
      } else {
        const float* rb = (p.rowbias && row_ok) ? p.rowbias + (size_t)(row / p.rows_per_batch) * p.ld_rowbias : nullptr;
        constexpr int NCH = BN / 32;
#pragma unroll 1
        for (int c = half ? (NCH + 1) / 2 : 0; c < (half ? NCH : (NCH + 1) / 2); ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          const int n0 = n_blk * BN + c * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.ln_stats) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + n0 + j));
              v[j] = ln_rstd * (v[j] - ln_mu * t.x); v[j + 1] = ln_rstd * (v[j + 1] - ln_mu * t.y);
              v[j + 2] = ln_rstd * (v[j + 2] - ln_mu * t.z); v[j + 3] = ln_rstd * (v[j + 3] - ln_mu * t.w);
            }
          }
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
              v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
            }
          }
          if (rb) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(rb + n0 + j));
              v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
            }
          }
          if (p.act == GLG_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
          }
          if (p.gate) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= gate;
          }
          if (row_ok) {
            if (p.residual) {
              const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + (size_t)row * p.ldr + n0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = __ldg(r4 + i);
                float2 f;
                f = unpack_bf16x2(u.x); v[8 * i + 0] += f.x; v[8 * i + 1] += f.y;
                f = unpack_bf16x2(u.y); v[8 * i + 2] += f.x; v[8 * i + 3] += f.y;
                f = unpack_bf16x2(u.z); v[8 * i + 4] += f.x; v[8 * i + 5] += f.y;
                f = unpack_bf16x2(u.w); v[8 * i + 6] += f.x; v[8 * i + 7] += f.y;
              }
            }
            if (p.out_fp32) {
              float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_off + n0);
#pragma unroll
              for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
              if (p.stats_out) {      // statistics of the values as stored (bf16-rounded): what the consumer GEMM will read
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float rv = __bfloat162float(__float2bfloat16(v[j]));
                  st_sum += rv; st_sq = fmaf(rv, rv, st_sq);
                }
              }
              epi_store_bf16(reinterpret_cast<bf16*>(p.out) + out_off + n0, v);
            }
          }
        }
      }
      if (p.stats_out && row_ok) {
        float2* so = reinterpret_cast<float2*>(p.stats_out) + (size_t)row * p.stats_slots;
        so[n_blk * 2 + half] = make_float2(st_sum, st_sq);
        if (n_blk == 0 && half == 0)
          for (int i = 2 * p.tiles_n; i < p.stats_slots; ++i) so[i] = make_float2(0.f, 0.f);
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      acc ^= 1; if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor-map cache + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

struct TmapKey {
  const void* ptr; uint64_t d[4]; uint64_t s[3]; uint32_t box[4]; int rank;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    return h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static std::mutex g_tmap_mu;

// bf16 tensor map, 128B swizzle, zero OOB fill.  dims/strides innermost first; strides in bytes (rank-1 of them).
int get_tmap_bf16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                  const uint32_t* box) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.rank = rank;
  for (int i = 0; i < rank; ++i) { key.d[i] = dims[i]; key.box[i] = box[i]; }
  for (int i = 0; i < rank - 1; ++i) key.s[i] = strides[i];
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) { *out = it->second; return 0; }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gd[4]; cuuint64_t gs[3]; cuuint32_t bx[4]; cuuint32_t es[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides[i];
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu stride0 %llu box %u %u %u %u ptr %p",
             (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
             (unsigned long long)(rank > 1 ? strides[0] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
             rank > 3 ? box[3] : 0, ptr);
    return set_error(buf);
  }
  g_tmaps.emplace(key, m);
  *out = m;
  return 0;
}

static int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, bool GEGLU>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmKParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  auto kern = gemm_tc_kernel<BN, GEGLU>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(std::string("cudaFuncSetAttribute(gemm): ") + cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, 320, Cfg::SMEM_BYTES, st>>>(ta, tb, p);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(std::string("gemm launch: ") + cudaGetErrorString(e));
  return 0;
}

static int pick_bn(int M, int N, int forced) {
  if (forced) return forced;
  const int cands[4] = {256, 160, 128, 64};
  const float penalty[4] = {0.0f, 0.04f, 0.08f, 0.30f};
  const int tiles_m = (M + 127) / 128;
  int best = 0; float best_cost = 1e30f;
  for (int i = 0; i < 4; ++i) {
    if (N % cands[i]) continue;
    const int tiles = tiles_m * (N / cands[i]);
    const int waves = (tiles + num_sms() - 1) / num_sms();
    const float cost = (float)waves * (cands[i] + 24.0f) * (1.0f + penalty[i]);
    if (cost < best_cost) { best_cost = cost; best = cands[i]; }
  }
  return best;
}

int g_force_bn = 0;   // test hook (glg_debug_force_bn)

}  // namespace glg

using namespace glg;

extern "C" void glg_debug_force_bn(int bn) { glg::g_force_bn = bn; }

extern "C" int glg_gemm(const GlgGemmArgs* a, void* stream) {
  if (!a) return set_error("glg_gemm: null args");
  if (a->K <= 0 || a->K % 64) return set_error("glg_gemm: K must be a positive multiple of 64");
  if (a->M <= 0 || a->N <= 0) return set_error("glg_gemm: M, N must be positive");
  if ((a->lda % 8) || (a->ldc % 8) || (a->residual && (a->ldr % 8))) return set_error("glg_gemm: leading dims must be multiples of 8");
  if (((uintptr_t)a->A | (uintptr_t)a->W | (uintptr_t)a->out | (uintptr_t)a->residual) & 15) return set_error("glg_gemm: pointers must be 16-byte aligned");
  if (a->rowbias && ((a->ld_rowbias % 4) || a->rows_per_batch <= 0)) return set_error("glg_gemm: bad rowbias args");
  int bn;
  if (a->geglu) {
    if (a->N % 256 || !a->bias || a->out_fp32) return set_error("glg_gemm: geglu needs N % 256 == 0, a bias and bf16 output");
    bn = 256;
  } else {
    bn = pick_bn(a->M, a->N, (g_force_bn && a->N % g_force_bn == 0) ? g_force_bn : 0);
    if (!bn) return set_error("glg_gemm: N must be a multiple of 64");
  }
  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->M; p.N = a->N;
  p.kb_per_tap = a->K / 64;
  p.num_kb = a->conv_mode ? 9 * p.kb_per_tap : p.kb_per_tap;
  p.tiles_m = (a->M + 127) / 128;
  p.tiles_n = a->N / bn;
  p.conv = a->conv_mode;
  p.out = a->out; p.ldc = a->ldc; p.out_fp32 = a->out_fp32;
  p.bias = a->bias; p.rowbias = a->rowbias; p.ld_rowbias = a->ld_rowbias; p.rows_per_batch = a->rows_per_batch > 0 ? a->rows_per_batch : 1;
  p.act = a->act; p.gate = a->gate; p.residual = reinterpret_cast<const bf16*>(a->residual); p.ldr = a->ldr;
  if (a->ln_stats) {
    if (!a->ln_colsum || a->ln_slots <= 0 || a->conv_mode) return set_error("glg_gemm: LayerNorm fold needs ln_colsum, ln_slots > 0 and a plain GEMM");
    if (((uintptr_t)a->ln_stats & 7) || ((uintptr_t)a->ln_colsum & 15)) return set_error("glg_gemm: ln_stats / ln_colsum alignment");
    p.ln_stats = a->ln_stats; p.ln_slots = a->ln_slots; p.ln_colsum = a->ln_colsum; p.ln_eps = a->ln_eps; p.inv_k = 1.0f / (float)a->K;
  }
  if (a->stats_out) {
    if (a->geglu || a->out_fp32 || 2 * p.tiles_n > a->stats_slots || ((uintptr_t)a->stats_out & 7))
      return set_error("glg_gemm: stats_out needs a bf16 non-GEGLU output and stats_slots >= 2 * ceil(N / tile)");
    p.stats_out = a->stats_out; p.stats_slots = a->stats_slots;
  }
  if (a->out_rows_per_batch > 0) {
    if (a->out_batch_stride % 8) return set_error("glg_gemm: out_batch_stride must be a multiple of 8");
    p.orpb = a->out_rows_per_batch; p.obs = a->out_batch_stride;
  }
  if (a->bias && ((uintptr_t)a->bias & 15)) return set_error("glg_gemm: bias must be 16-byte aligned");

  CUtensorMap ta, tb;
  if (a->conv_mode) {
    const int H = a->H, W = a->Wd, B = a->Bn;
    if (H <= 0 || W <= 0 || B <= 0 || (long long)B * H * W != a->M) return set_error("glg_gemm: conv dims do not match M");
    if (W > 128 || (128 % W)) return set_error("glg_gemm: conv width must divide 128");
    const int HW = H * W;
    uint32_t box[4];
    if (HW >= 128) {
      if (HW % 128) return set_error("glg_gemm: conv H*W must be a multiple of 128 (or divide it)");
      box[0] = 64; box[1] = W; box[2] = 128 / W; box[3] = 1;
    } else {
      if (128 % HW) return set_error("glg_gemm: conv H*W must divide 128");
      box[0] = 64; box[1] = W; box[2] = H; box[3] = 128 / HW;
    }
    p.HW = HW; p.Wd = W;
    const uint64_t dims[4] = {(uint64_t)a->K, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)a->lda * 2, (uint64_t)a->lda * 2 * W, (uint64_t)a->lda * 2 * HW};
    if (get_tmap_bf16(&ta, a->A, 4, dims, str, box)) return -1;
    const uint64_t wd[2] = {(uint64_t)a->K, (uint64_t)a->N * 9};
    const uint64_t ws[1] = {(uint64_t)a->K * 2};
    const uint32_t wb[2] = {64, (uint32_t)bn};
    if (get_tmap_bf16(&tb, a->W, 2, wd, ws, wb)) return -1;
  } else {
    const uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
    const uint64_t str[1] = {(uint64_t)a->lda * 2};
    const uint32_t box[2] = {64, 128};
    if (get_tmap_bf16(&ta, a->A, 2, dims, str, box)) return -1;
    const uint64_t wd[2] = {(uint64_t)a->K, (uint64_t)a->N};
    const uint64_t ws[1] = {(uint64_t)a->K * 2};
    const uint32_t wb[2] = {64, (uint32_t)bn};
    if (get_tmap_bf16(&tb, a->W, 2, wd, ws, wb)) return -1;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->geglu) return launch_gemm<256, true>(ta, tb, p, st);
  switch (bn) {
    case 256: return launch_gemm<256, false>(ta, tb, p, st);
    case 160: return launch_gemm<160, false>(ta, tb, p, st);
    case 128: return launch_gemm<128, false>(ta, tb, p, st);
    case 64:  return launch_gemm<64, false>(ta, tb, p, st);
  }
  return set_error("glg_gemm: internal: bad BN");
}
