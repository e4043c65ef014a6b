// GroupNorm (+SiLU) and LayerNorm over channels-last bf16 activations; fp32 statistics.
// HBM-bound streaming kernels: 16-byte vector accesses, fully coalesced rows.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <string>

#include "common.cuh"
#include "internal.h"
#include "../../include/gligen_b200.h"

namespace glg {

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  float2 f;
  f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  float2 f;
  f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ---- GroupNorm, ONE kernel: statistics -> per-sample barrier -> normalise + affine (+SiLU) ---------------------
// grid (chunks, B), every CTA co-resident (the launcher sizes the grid by the occupancy of this kernel);
// blockDim = (C/8) * rpi, thread -> (fixed channel vector cv of 8 channels, row lane rl); a CTA owns rows
// [chunk * rows_per_chunk, ...) of one sample.
//   1. SHIFTED moments: every channel is accumulated relative to its own pivot p_c = x[b, row 0, c] (one 16-byte load
//      per thread, no reduction), so neither sum ever holds a large mean; when the channels of a group are merged the
//      sums are moved onto the group's common pivot P_g = p_(first channel of g) by exact algebra
//          S1' = S1 + n d,   S2' = S2 + 2 d S1 + n d^2,   d = p_c - P_g.
//      The variance S2'/N - (S1'/N)^2 then cancels at most by the group's own between-channel spread - the raw
//      E[x^2] - E[x]^2 form loses everything when |mean| >> std (util.py:223-225 runs F.group_norm in fp32: two-pass).
//   2. per-CTA partial (S1', S2') per group -> scratch; sample-wide barrier (arrive counter + spin; the counters reset
//      themselves when the last CTA leaves); every CTA then sums the partials of its sample in chunk order: no
//      floating-point atomics anywhere, a forward is bit-reproducible run to run.
//   3. apply: the activation is read a second time (an L2 hit: producer -> GroupNorm -> consumer tensors of this
//      model fit the 126 MB L2) and written once.
// scratch layout (32-bit words): arrive[64] | depart[64] (uint, zero before the first use, self-resetting)
//                                | partial[B][chunks][G][2] fp32
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__global__ void gn_fused_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ y, long long ldy,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ scratch,
                                int B, int HW, int C, int groups, float eps, int silu, int rows_per_chunk) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float gn_sm[];          // [2][rpi][C] row-lane partials (>= 2 * 512 floats) | [C] pivots | mean[groups] | rstd[groups]
  const int vec = C >> 3;
  const int rpi = blockDim.x / vec;
  const int red = rpi * C > 512 ? rpi * C : 512;
  float* s_sum = gn_sm;
  float* s_sq = gn_sm + red;
  float* s_piv = gn_sm + 2 * red;
  float* s_mean = s_piv + C;
  float* s_rstd = s_mean + groups;
  const int cv = threadIdx.x % vec, rl = threadIdx.x / vec;
  const int b = blockIdx.y, chunks = gridDim.x;
  const int cpg = C / groups;
  unsigned int* arrive = reinterpret_cast<unsigned int*>(scratch) + b;
  unsigned int* depart = reinterpret_cast<unsigned int*>(scratch) + 64 + b;
  float* part_b = scratch + 128 + (long long)b * chunks * groups * 2;
  const bf16* xb = x + (long long)b * HW * ldx + cv * 8;

  // ---- 1. shifted moments of this CTA's rows (pivot = the channel's value in row 0 of the sample)
  float pv[8];
  load8(xb, pv);
  if (rl == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s_piv[cv * 8 + j] = pv[j];
  }
  const int r0 = blockIdx.x * rows_per_chunk;
  const int r1 = min(HW, r0 + rows_per_chunk);
  float a[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = 0.f; q[j] = 0.f; }
  int r = r0 + rl;
  for (; r + 3 * rpi < r1; r += 4 * rpi) {      // 4 independent 16-byte loads in flight per thread
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = __ldg(reinterpret_cast<const uint4*>(xb + (long long)(r + i * rpi) * ldx));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v[8];
      unpack8(u[i], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[j] - pv[j]; a[j] += d; q[j] = fmaf(d, d, q[j]); }
    }
  }
  for (; r < r1; r += rpi) {
    float v[8];
    load8(xb + (long long)r * ldx, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = v[j] - pv[j]; a[j] += d; q[j] = fmaf(d, d, q[j]); }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { s_sum[rl * C + cv * 8 + j] = a[j]; s_sq[rl * C + cv * 8 + j] = q[j]; }
  __syncthreads();
  // one thread per group: its channels' row-lane partials in fixed order, moved onto the group's common pivot
  const int rows_here = max(r1 - r0, 0);
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    const float P = s_piv[g * cpg];
    float s = 0.f, ss = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      float s1 = 0.f, s2 = 0.f;
      for (int k = 0; k < rpi; ++k) { s1 += s_sum[k * C + c]; s2 += s_sq[k * C + c]; }
      const float d = s_piv[c] - P;
      s += fmaf((float)rows_here, d, s1);
      ss += fmaf(d, fmaf((float)rows_here, d, 2.f * s1), s2);
    }
    *reinterpret_cast<float2*>(part_b + ((long long)blockIdx.x * groups + g) * 2) = make_float2(s, ss);
  }
  // ---- sample-wide barrier: every CTA of sample b has published its partials
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(arrive, 1u);
    unsigned int seen;
    unsigned long long t0 = 0;
    while (true) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(arrive) : "memory");
      if (seen >= (unsigned)chunks) break;
      __nanosleep(32);
      if (t0 == 0) t0 = globaltimer_ns();
      else if (globaltimer_ns() - t0 > 4000000000ull) {   // co-residency violated: trap instead of hanging the box
        printf("glg: groupnorm barrier timeout (sample %d chunk %d: %u of %d arrived)\n", b, blockIdx.x, seen, chunks);
        __trap();
      }
    }
  }
  __syncthreads();
  {
    // fixed order, spread over the block: thread (g, part) sums chunks part, part + P, ...; then the P sums in order
    const int P = blockDim.x / groups;
    const int g = threadIdx.x % groups, part = threadIdx.x / groups;
    if (part < P) {
      float s = 0.f, ss = 0.f;
      for (int k = part; k < chunks; k += P) {
        const float2 t = __ldcg(reinterpret_cast<const float2*>(part_b + ((long long)k * groups + g) * 2));
        s += t.x; ss += t.y;
      }
      s_sum[part * groups + g] = s;                    // P * groups <= blockDim <= 512 floats per half
      s_sq[part * groups + g] = ss;
    }
    __syncthreads();
    if (threadIdx.x < groups) {
      float ts = 0.f, tss = 0.f;
      for (int i = 0; i < P; ++i) { ts += s_sum[i * groups + threadIdx.x]; tss += s_sq[i * groups + threadIdx.x]; }
      const float inv_n = 1.f / ((float)HW * (float)cpg);
      const float m1 = ts * inv_n;                                       // mean of the shifted data
      const float var = fmaxf(tss * inv_n - m1 * m1, 0.f);
      s_mean[threadIdx.x] = s_piv[threadIdx.x * cpg] + m1;
      s_rstd[threadIdx.x] = rsqrtf(var + eps);
    }
    if (threadIdx.x == 0) {
      // everyone who reaches this point has left the spin loop: the last one out re-arms the counters
      if (atomicAdd(depart, 1u) == (unsigned)chunks - 1u) { *depart = 0u; *arrive = 0u; __threadfence(); }
    }
    __syncthreads();
  }

  // ---- 3. normalise + affine (+SiLU): 16-byte load / 8 FMA (+SiLU) / 16-byte store stream
  float sa[8], sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    const int g = c / cpg;
    const float ga = __ldg(gamma + c) * s_rstd[g];
    sa[j] = ga;
    sb[j] = __ldg(beta + c) - s_mean[g] * ga;
  }
  bf16* yb = y + (long long)b * HW * ldy + cv * 8;
  r = r0 + rl;
  for (; r + 3 * rpi < r1; r += 4 * rpi) {
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = __ldcg(reinterpret_cast<const uint4*>(xb + (long long)(r + i * rpi) * ldx));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v[8];
      unpack8(u[i], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = fmaf(v[j], sa[j], sb[j]);
        v[j] = silu ? silu_fast(t) : t;
      }
      store8(yb + (long long)(r + i * rpi) * ldy, v);
    }
  }
  for (; r < r1; r += rpi) {
    float v[8];
    const uint4 u = __ldcg(reinterpret_cast<const uint4*>(xb + (long long)r * ldx));
    unpack8(u, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = fmaf(v[j], sa[j], sb[j]);
      v[j] = silu ? silu_fast(t) : t;
    }
    store8(yb + (long long)r * ldy, v);
  }
}

// ---- small tensors: ONE kernel, one CTA per (group, batch): stats + apply, deterministic -------------
// Used when H*W <= 256 (the 16x16 / 8x8 levels): the three-launch path above is launch/latency bound there.
__global__ void gn_small_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ y, long long ldy,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                int HW, int C, int groups, float eps, int silu) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[2][32];
  __shared__ float s_pivot;
  const int g = blockIdx.x, b = blockIdx.y;
  const int cpg = C / groups, half = cpg >> 1;        // cpg is even: process bf16 pairs
  const int total = HW * half;
  const bf16* xb = x + (long long)b * HW * ldx + g * cpg;
  bf16* yb = y + (long long)b * HW * ldy + g * cpg;
  // pivot = mean of the group's channels in row 0; moments are taken on x - pivot (see gn_fused_kernel)
  if (threadIdx.x < 32) {
    float ps = 0.f;
    for (int c = threadIdx.x; c < cpg; c += 32) ps += __bfloat162float(xb[c]);
    ps = warp_sum(ps);
    if (threadIdx.x == 0) s_pivot = ps / (float)cpg;
  }
  __syncthreads();
  const float pivot = s_pivot;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = i / half, c2 = i - r * half;
    const float2 f = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(xb + (long long)r * ldx + c2 * 2)));
    const float d0 = f.x - pivot, d1 = f.y - pivot;
    s += d0 + d1;
    ss = fmaf(d0, d0, fmaf(d1, d1, ss));
  }
  s = warp_sum(s); ss = warp_sum(ss);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { red[0][warp] = s; red[1][warp] = ss; }
  __syncthreads();
  float ts = 0.f, tss = 0.f;
  for (int w = 0; w < nw; ++w) { ts += red[0][w]; tss += red[1][w]; }      // fixed order: deterministic
  const float inv_n = 1.f / (float)(HW * cpg);
  const float m1 = ts * inv_n;
  const float mean = pivot + m1;
  const float rstd = rsqrtf(fmaxf(tss * inv_n - m1 * m1, 0.f) + eps);
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = i / half, c2 = i - r * half;
    const int c = g * cpg + c2 * 2;
    const float2 f = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(xb + (long long)r * ldx + c2 * 2)));
    const float g0 = gamma[c] * rstd, g1 = gamma[c + 1] * rstd;
    float t0 = fmaf(f.x, g0, beta[c] - mean * g0);
    float t1 = fmaf(f.y, g1, beta[c + 1] - mean * g1);
    if (silu) { t0 = silu_fast(t0); t1 = silu_fast(t1); }
    *reinterpret_cast<uint32_t*>(yb + (long long)r * ldy + c2 * 2) = pack_bf16x2(t0, t1);
  }
}

// ---- small tensors, register-resident: one CTA per (group, sample) holds its whole [HW x cpg] slab in registers -----
// (8-byte units; H*W * cpg <= 256 threads * MAXU units * 4): ONE global read, shifted moments, ONE write.  The 16x16 and 8x8
// levels (C = 1280 / 1920 / 2560: cpg = 40 / 60 / 80) run 34 of the 61 norms of a UNet pass through this kernel.
template <int MAXU>
__global__ void __launch_bounds__(256) gn_reg_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ y, long long ldy,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     int HW, int C, int groups, float eps, int silu) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[2][8];
  const int g = blockIdx.x, b = blockIdx.y;
  const int cpg = C / groups, upr = cpg >> 2;              // 8-byte units per row
  const int total = HW * upr;
  const bf16* xb = x + (long long)b * HW * ldx + g * cpg;
  bf16* yb = y + (long long)b * HW * ldy + g * cpg;
  uint2 u[MAXU];
#pragma unroll
  for (int i = 0; i < MAXU; ++i) {
    const int idx = threadIdx.x + i * 256;
    if (idx < total) {
      const int r = idx / upr, c4 = idx - r * upr;
      u[i] = __ldg(reinterpret_cast<const uint2*>(xb + (long long)r * ldx + c4 * 4));
    } else {
      u[i] = make_uint2(0u, 0u);
    }
  }
  // pivot: element (row 0, first channel of the group); every thread reads it (L1 broadcast)
  const float pivot = __bfloat162float(xb[0]);
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXU; ++i) {
    if (threadIdx.x + i * 256 < total) {
      const float2 f0 = unpack_bf16x2(u[i].x), f1 = unpack_bf16x2(u[i].y);
      const float d0 = f0.x - pivot, d1 = f0.y - pivot, d2 = f1.x - pivot, d3 = f1.y - pivot;
      s += (d0 + d1) + (d2 + d3);
      ss = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, ss))));
    }
  }
  s = warp_sum(s); ss = warp_sum(ss);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s; red[1][warp] = ss; }
  __syncthreads();
  float ts = 0.f, tss = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { ts += red[0][w]; tss += red[1][w]; }      // fixed order: deterministic
  const float inv_n = 1.f / (float)(HW * cpg);
  const float m1 = ts * inv_n;
  const float mean = pivot + m1;
  const float rstd = rsqrtf(fmaxf(tss * inv_n - m1 * m1, 0.f) + eps);
#pragma unroll
  for (int i = 0; i < MAXU; ++i) {
    const int idx = threadIdx.x + i * 256;
    if (idx < total) {
      const int r = idx / upr, c4 = idx - r * upr;
      const int c = g * cpg + c4 * 4;
      const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float2 f0 = unpack_bf16x2(u[i].x), f1 = unpack_bf16x2(u[i].y);
      float t0 = fmaf((f0.x - mean) * rstd, ga.x, be.x), t1 = fmaf((f0.y - mean) * rstd, ga.y, be.y);
      float t2 = fmaf((f1.x - mean) * rstd, ga.z, be.z), t3 = fmaf((f1.y - mean) * rstd, ga.w, be.w);
      if (silu) { t0 = silu_fast(t0); t1 = silu_fast(t1); t2 = silu_fast(t2); t3 = silu_fast(t3); }
      *reinterpret_cast<uint2*>(yb + (long long)r * ldy + c4 * 4) = make_uint2(pack_bf16x2(t0, t1), pack_bf16x2(t2, t3));
    }
  }
}

// ---- LayerNorm: one warp per row, two-pass in registers ----------------------------------------
template <int MAXV>
__global__ void ln_kernel(const bf16* __restrict__ x, long long x_batch, bf16* __restrict__ y, long long y_batch,
                          const float* __restrict__ gamma, const float* __restrict__ beta,
                          int rows, int C, float eps, long long total_rows) {
  pdl_trigger();
  pdl_wait();
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= total_rows) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(gw / rows), r = (int)(gw % rows);
  const bf16* xr = x + (long long)b * x_batch + (long long)r * C;
  bf16* yr = y + (long long)b * y_batch + (long long)r * C;
  const int vec = C >> 3;
  float v[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int cv = lane + i * 32;
    if (cv < vec) {
      load8(xr + cv * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int cv = lane + i * 32;
    if (cv < vec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float dlt = v[i][j] - mean; ss = fmaf(dlt, dlt, ss); }
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int cv = lane + i * 32;
    if (cv < vec) {
      float o[8];
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + cv * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + cv * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + cv * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + cv * 8 + 4));
      o[0] = (v[i][0] - mean) * rstd * g0.x + b0.x; o[1] = (v[i][1] - mean) * rstd * g0.y + b0.y;
      o[2] = (v[i][2] - mean) * rstd * g0.z + b0.z; o[3] = (v[i][3] - mean) * rstd * g0.w + b0.w;
      o[4] = (v[i][4] - mean) * rstd * g1.x + b1.x; o[5] = (v[i][5] - mean) * rstd * g1.y + b1.y;
      o[6] = (v[i][6] - mean) * rstd * g1.z + b1.z; o[7] = (v[i][7] - mean) * rstd * g1.w + b1.w;
      store8(yr + cv * 8, o);
    }
  }
}

}  // namespace glg

using namespace glg;

extern "C" int glg_groupnorm(const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma, const float* beta,
                             float* stats, int32_t B, int32_t HW, int32_t C, int32_t groups, float eps, int32_t silu,
                             void* stream) {
  if (C % 8 || C % groups || (ldx % 8) || (ldy % 8)) return set_error("glg_groupnorm: C and leading dims must be multiples of 8, C % groups == 0");
  if (C > 4096) return set_error("glg_groupnorm: C too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static int small_mode = -1;        // GLG_GN_SMALL=0: the barrier kernel also for H*W <= 256 (measured slower there: 19 vs 15 us)
  if (small_mode < 0) { const char* e = getenv("GLG_GN_SMALL"); small_mode = e ? atoi(e) : 1; }
  {
    // register-resident path: the (group, sample) slab fits 256 threads x 20 eight-byte units, rows and group starts 8-byte aligned
    const int cpg = C / groups;
    const long long units = (long long)HW * (cpg / 4);
    if (small_mode && cpg % 4 == 0 && units <= 256 * 20 && !(((uintptr_t)x | (uintptr_t)y) & 7) && (C % 4 == 0)) {
      dim3 grid(groups, B);
      const bf16* xp = (const bf16*)x; bf16* yp = (bf16*)y;
      if (units <= 256 * 5) launch_k(gn_reg_kernel<5>, dim3(grid), dim3(256), 0, st, 1, xp, ldx, yp, ldy, gamma, beta, HW, C, groups, eps, silu);
      else if (units <= 256 * 10) launch_k(gn_reg_kernel<10>, dim3(grid), dim3(256), 0, st, 1, xp, ldx, yp, ldy, gamma, beta, HW, C, groups, eps, silu);
      else launch_k(gn_reg_kernel<20>, dim3(grid), dim3(256), 0, st, 1, xp, ldx, yp, ldy, gamma, beta, HW, C, groups, eps, silu);
      count_launch();
      return check_launch("gn_reg launch");
    }
  }
  if (small_mode && HW <= 256 && ((C / groups) % 2 == 0)) {
    dim3 grid(groups, B);
    launch_k(gn_small_kernel, dim3(grid), dim3(256), 0, st, 1, (const bf16*)x, ldx, (bf16*)y, ldy, gamma, beta, HW, C, groups, eps, silu);
    count_launch();
    return check_launch("gn_small launch");
  }
  const int vec = C / 8;
  int rpi = 256 / vec; if (rpi < 1) rpi = 1;
  const int threads = vec * rpi;          // <= 512 for C <= 4096
  if (threads > 1024) return set_error("glg_groupnorm: C too large for one block");
  if (B > 64) return set_error("glg_groupnorm: at most 64 samples per call");
  const size_t red = (size_t)rpi * C > 512 ? (size_t)rpi * C : 512;
  const size_t smem = (2 * red + (size_t)C + 2 * (size_t)groups) * sizeof(float);
  // every CTA of the grid must be resident at once (sample-wide barrier inside the kernel)
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return set_error(std::string("cudaFuncSetAttribute(gn_fused): ") + cudaGetErrorString(e));
    attr_set = true;
  }
  int occ = 0;
  {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_fused_kernel, threads, smem);
    if (e != cudaSuccess || occ < 1) return set_error(std::string("glg_groupnorm: occupancy query failed: ") + cudaGetErrorString(e));
    if (occ > 8) occ = 8;
  }
  const int capacity = num_sms() * occ;
  int chunks = capacity / B;
  const int max_chunks = (HW + rpi * 4 - 1) / (rpi * 4);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int rows_per_chunk = (HW + chunks - 1) / chunks;
  chunks = (HW + rows_per_chunk - 1) / rows_per_chunk;
  if (128 + (long long)B * chunks * groups * 2 > (long long)GLG_GN_SCRATCH_FLOATS(B, groups)) return set_error("glg_groupnorm: internal scratch sizing");
  dim3 grid(chunks, B);
  launch_k(gn_fused_kernel, dim3(grid), dim3(threads), smem, st, 1, (const bf16*)x, ldx, (bf16*)y, ldy, gamma, beta, stats, B, HW, C, groups, eps, silu, rows_per_chunk);
  count_launch();
  return check_launch("gn_fused launch");
}

extern "C" int glg_layernorm(const void* x, int64_t x_batch, void* y, int64_t y_batch, const float* gamma, const float* beta,
                             int32_t B, int32_t rows, int32_t C, float eps, void* stream) {
  if (C % 8 || C > 2048) return set_error("glg_layernorm: C must be a multiple of 8 and <= 2048");
  if ((x_batch % 8) || (y_batch % 8)) return set_error("glg_layernorm: batch strides must be multiples of 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)B * rows;
  const int wpb = 8;
  const unsigned grid = (unsigned)((total + wpb - 1) / wpb);
  const int vec = C / 8;
  if (vec <= 64) launch_k(ln_kernel<2>, dim3(grid), dim3(wpb * 32), 0, st, 1, (const bf16*)x, x_batch, (bf16*)y, y_batch, gamma, beta, rows, C, eps, total);
  else if (vec <= 160) launch_k(ln_kernel<5>, dim3(grid), dim3(wpb * 32), 0, st, 1, (const bf16*)x, x_batch, (bf16*)y, y_batch, gamma, beta, rows, C, eps, total);
  else launch_k(ln_kernel<8>, dim3(grid), dim3(wpb * 32), 0, st, 1, (const bf16*)x, x_batch, (bf16*)y, y_batch, gamma, beta, rows, C, eps, total);
  count_launch();
  return check_launch("layernorm launch");
}
