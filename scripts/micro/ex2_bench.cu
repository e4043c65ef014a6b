// Micro-benchmark: MUFU.EX2 throughput for f32 vs packed bf16x2 / f16x2 operands (sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 8 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { float f = __uint_as_float(x[i]); asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(f)); x[i] = __float_as_uint(f); }
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(x[i]) : "r"(x[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(x[i]) : "r"(x[i]));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name) {
  uint32_t* o; cudaMalloc(&o, 148 * 8 * 256 * 4);
  const int iters = 4096;
  k<MODE><<<148 * 8, 256>>>(o, 16, 1); cudaDeviceSynchronize();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); k<MODE><<<148 * 8, 256>>>(o, iters, 1); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ops = 148.0 * 8 * 256 * 8.0 * iters;   // instructions (per lane)
  printf("%-10s %.3f ms  %.2f lane-instr/clk/SM (at 1.965 GHz)  elements/clk/SM %.2f\n", name, ms, ops / (ms * 1e-3) / 1.965e9 / 148, (MODE ? 2 : 1) * ops / (ms * 1e-3) / 1.965e9 / 148);
  cudaFree(o);
}
int main() { run<0>("f32"); run<1>("bf16x2"); run<2>("f16x2"); return 0; }
