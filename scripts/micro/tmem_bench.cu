// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of warps per CTA and CTAs per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../gligen_b200/csrc/common.cuh"
using namespace glg;
// MODE 0: ld32 (4 KB / warp-instr), 1: st32, 2: ld16
template <int MODE>
__global__ void k(uint32_t* out, int iters, int ncols) {
  __shared__ uint32_t s_taddr;
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&s_taddr), ncols); tmem_relinquish(); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_taddr;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t colgrp = (uint32_t)(warp >> 2);
  uint32_t acc = 0;
  uint32_t r[32], r2[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = r2[i] = i + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    const uint32_t c0 = ((colgrp * 64 + (it & 1) * 32) % ncols);
    if (MODE == 0) {
      tmem_ld32(base + lane_base + c0, r);
      tmem_ld32(base + lane_base + ((c0 + 64) % ncols), r2);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= r[i] ^ r2[i];
    } else if (MODE == 1) {
      tmem_st32(base + lane_base + c0, r);
      tmem_st32(base + lane_base + ((c0 + 64) % ncols), r2);
      tmem_st_wait();
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(base, ncols);
}
template <int MODE>
void run(const char* name, int warps, int ctas_per_sm) {
  uint32_t* o; cudaMalloc(&o, 148 * 2 * 1024 * 4);
  const int iters = 20000;
  const int ncols = ctas_per_sm == 1 ? 512 : 256;
  k<MODE><<<148 * ctas_per_sm, warps * 32>>>(o, 100, ncols); cudaDeviceSynchronize();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); k<MODE><<<148 * ctas_per_sm, warps * 32>>>(o, iters, ncols); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  double bytes_per_sm = (double)ctas_per_sm * warps * 2.0 * 4096.0 * iters;
  printf("%-6s warps/CTA=%2d CTAs/SM=%d : %.3f ms  %.1f B/clk/SM (1.965 GHz)  %s\n", name, warps, ctas_per_sm, ms, bytes_per_sm / (ms * 1e-3) / 1.965e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(o);
}
int main() {
  for (int c = 1; c <= 2; ++c) for (int w : {4, 8, 16}) run<0>("ld32", w, c);
  for (int c = 1; c <= 2; ++c) for (int w : {4, 8, 16}) run<1>("st32", w, c);
  return 0;
}
