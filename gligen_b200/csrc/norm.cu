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

// ---- GroupNorm pass 1: per-(batch, group) mean / rstd, DETERMINISTIC ---------------------------------
// grid (chunks, B); blockDim = (C/8) * rpi, thread -> (fixed channel vector cv, row lane rl).
// Each CTA writes its per-group partial (sum, sumsq) to scratch; the last CTA of a batch (atomic ticket)
// reduces the partials in chunk order - no floating-point atomics on the result, so a forward is
// bit-reproducible run to run.
// scratch layout (floats): ticket[64] (uint, zero before the first use, reset by the last CTA) | final[B][G][2] (sum, sumsq)
//                          | partial[B][chunks][G][2]
__global__ void gn_stats_kernel(const bf16* __restrict__ x, long long ldx, float* __restrict__ scratch,
                                int B, int HW, int C, int groups, int rows_per_chunk) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float gn_sm[];          // [2][max(C, 512)]
  __shared__ bool is_last;
  float* s_sum = gn_sm;
  float* s_sq = gn_sm + (C > 512 ? C : 512);
  const int vec = C >> 3;
  const int rpi = blockDim.x / vec;
  const int cv = threadIdx.x % vec, rl = threadIdx.x / vec;
  const int b = blockIdx.y, chunks = gridDim.x;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch) + b;
  float* fin = scratch + 64 + (long long)b * groups * 2;
  float* part = scratch + 64 + (long long)B * groups * 2 + ((long long)b * chunks + blockIdx.x) * groups * 2;
  const int r0 = blockIdx.x * rows_per_chunk;
  const int r1 = min(HW, r0 + rows_per_chunk);
  float a[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = 0.f; q[j] = 0.f; }
  const bf16* xb = x + (long long)b * HW * ldx + cv * 8;
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
      for (int j = 0; j < 8; ++j) { a[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
    }
  }
  for (; r < r1; r += rpi) {
    float v[8];
    load8(xb + (long long)r * ldx, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
  }
  // fixed-order reduction over the rpi row lanes: lane rl == k adds in turn
  for (int k = 0; k < rpi; ++k) {
    if (rl == k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (k == 0) { s_sum[cv * 8 + j] = a[j]; s_sq[cv * 8 + j] = q[j]; }
        else { s_sum[cv * 8 + j] += a[j]; s_sq[cv * 8 + j] += q[j]; }
      }
    }
    __syncthreads();
  }
  const int cpg = C / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float s = 0.f, ss = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { s += s_sum[c]; ss += s_sq[c]; }
    part[g * 2 + 0] = s;
    part[g * 2 + 1] = ss;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == (unsigned)chunks - 1u);
  __syncthreads();
  if (is_last) {
    __threadfence();
    // last CTA of this sample: reduce the per-chunk partials.  Fixed order (deterministic) but spread over the whole
    // block: thread (g, part) sums chunks part, part + P, ... ; then the P partial sums are added in order.
    const float* pb = scratch + 64 + (long long)B * groups * 2 + (long long)b * chunks * groups * 2;
    const int P = blockDim.x / groups;                 // groups = 32: P = 7..10 row lanes
    const int g = threadIdx.x % groups, part = threadIdx.x / groups;
    float s = 0.f, ss = 0.f;
    if (part < P) {
      for (int k = part; k < chunks; k += P) {
        const float2 t = __ldcg(reinterpret_cast<const float2*>(pb + ((long long)k * groups + g) * 2));
        s += t.x; ss += t.y;
      }
      s_sum[part * groups + g] = s;                    // P * groups <= blockDim <= 512 floats per half
      s_sq[part * groups + g] = ss;
    }
    __syncthreads();
    if (threadIdx.x < groups) {
      float ts = 0.f, tss = 0.f;
      for (int i = 0; i < P; ++i) { ts += s_sum[i * groups + threadIdx.x]; tss += s_sq[i * groups + threadIdx.x]; }
      fin[threadIdx.x * 2] = ts;
      fin[threadIdx.x * 2 + 1] = tss;
    }
    if (threadIdx.x == 0) *ticket = 0u;       // ready for the next GroupNorm on this scratch (stream-ordered)
  }
}

// ---- GroupNorm pass 2: normalise + affine (+SiLU) -----------------------------------------------
// Same thread -> (channel vector, row lane) mapping as pass 1: the 8 scale/shift pairs of a thread live in
// registers, the row loop is a pure 16-byte load / 8 FMA (+SiLU) / 16-byte store stream.
__global__ void gn_apply_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ y, long long ldy,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ stats, int HW, int C, int groups, float eps, int silu,
                                int rows_per_chunk) {
  pdl_trigger();
  pdl_wait();
  const int vec = C >> 3;
  const int rpi = blockDim.x / vec;
  const int cv = threadIdx.x % vec, rl = threadIdx.x / vec;
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const float inv_n = 1.f / ((float)HW * (float)cpg);
  float sa[8], sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    const int g = c / cpg;
    const float mean = stats[64 + ((long long)b * groups + g) * 2] * inv_n;
    const float var = fmaxf(stats[64 + ((long long)b * groups + g) * 2 + 1] * inv_n - mean * mean, 0.f);
    const float ga = gamma[c] * rsqrtf(var + eps);
    sa[j] = ga;
    sb[j] = beta[c] - mean * ga;
  }
  const int r0 = blockIdx.x * rows_per_chunk;
  const int r1 = min(HW, r0 + rows_per_chunk);
  const bf16* xb = x + (long long)b * HW * ldx + cv * 8;
  bf16* yb = y + (long long)b * HW * ldy + cv * 8;
  int r = r0 + rl;
  for (; r + 3 * rpi < r1; r += 4 * rpi) {
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = __ldg(reinterpret_cast<const uint4*>(xb + (long long)(r + i * rpi) * ldx));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v[8];
      unpack8(u[i], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = fmaf(v[j], sa[j], sb[j]);
        v[j] = silu ? silu_f(t) : t;
      }
      store8(yb + (long long)(r + i * rpi) * ldy, v);
    }
  }
  for (; r < r1; r += rpi) {
    float v[8];
    load8(xb + (long long)r * ldx, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = fmaf(v[j], sa[j], sb[j]);
      v[j] = silu ? silu_f(t) : t;
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
  const int g = blockIdx.x, b = blockIdx.y;
  const int cpg = C / groups, half = cpg >> 1;        // cpg is even: process bf16 pairs
  const int total = HW * half;
  const bf16* xb = x + (long long)b * HW * ldx + g * cpg;
  bf16* yb = y + (long long)b * HW * ldy + g * cpg;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = i / half, c2 = i - r * half;
    const float2 f = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(xb + (long long)r * ldx + c2 * 2)));
    s += f.x + f.y;
    ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
  }
  s = warp_sum(s); ss = warp_sum(ss);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { red[0][warp] = s; red[1][warp] = ss; }
  __syncthreads();
  float ts = 0.f, tss = 0.f;
  for (int w = 0; w < nw; ++w) { ts += red[0][w]; tss += red[1][w]; }      // fixed order: deterministic
  const float inv_n = 1.f / (float)(HW * cpg);
  const float mean = ts * inv_n;
  const float rstd = rsqrtf(fmaxf(tss * inv_n - mean * mean, 0.f) + eps);
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = i / half, c2 = i - r * half;
    const int c = g * cpg + c2 * 2;
    const float2 f = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(xb + (long long)r * ldx + c2 * 2)));
    const float g0 = gamma[c] * rstd, g1 = gamma[c + 1] * rstd;
    float t0 = fmaf(f.x, g0, beta[c] - mean * g0);
    float t1 = fmaf(f.y, g1, beta[c + 1] - mean * g1);
    if (silu) { t0 = silu_f(t0); t1 = silu_f(t1); }
    *reinterpret_cast<uint32_t*>(yb + (long long)r * ldy + c2 * 2) = pack_bf16x2(t0, t1);
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
  if (HW <= 256 && ((C / groups) % 2 == 0)) {
    dim3 grid(groups, B);
    launch_k(gn_small_kernel, dim3(grid), dim3(256), 0, st, 1, (const bf16*)x, ldx, (bf16*)y, ldy, gamma, beta, HW, C, groups, eps, silu);
    count_launch();
    return check_launch("gn_small launch");
  }
  const int vec = C / 8;
  int rpi = 256 / vec; if (rpi < 1) rpi = 1;
  const int threads = vec * rpi;          // <= 512 for C <= 4096
  if (threads > 1024) return set_error("glg_groupnorm: C too large for one block");
  int chunks = (4 * 148 + B - 1) / B;
  int max_chunks = (HW + rpi * 4 - 1) / (rpi * 4);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int rows_per_chunk = (HW + chunks - 1) / chunks;
  chunks = (HW + rows_per_chunk - 1) / rows_per_chunk;
  dim3 grid(chunks, B);
  if (B > 64) return set_error("glg_groupnorm: at most 64 samples per call");
  if (64 + (long long)B * groups * 2 * (1 + chunks) > (long long)GLG_GN_SCRATCH_FLOATS(B, groups)) return set_error("glg_groupnorm: internal scratch sizing");
  launch_k(gn_stats_kernel, dim3(grid), dim3(threads), 2 * (C > 512 ? C : 512) * sizeof(float), st, 1, (const bf16*)x, ldx, stats, B, HW, C, groups, rows_per_chunk);
  count_launch();
  if (check_launch("gn_stats launch")) return -1;
  launch_k(gn_apply_kernel, dim3(grid), dim3(threads), 0, st, 1, (const bf16*)x, ldx, (bf16*)y, ldy, gamma, beta, stats, HW, C, groups, eps, silu, rows_per_chunk);
  count_launch();
  return check_launch("gn_apply launch");
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
