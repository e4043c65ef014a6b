// Micro-benchmark: per-SM throughput of the instructions in the attention softmax loop (sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
// MODE 0: FFMA  1: MUFU  2: pack (F2FP)  3: max3  4: softmax mix per 32 values: 32 FFMA + 32 MUFU + 32 FADD + 16 F2FP + 16 FMNMX3
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = seed + threadIdx.x * 0.001f + i;
  float acc = 0.f; uint32_t pacc = 0; float mx = -1e30f;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = fmaf(x[i], 1.0001f, -0.5f);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = ex2(x[i]);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { uint32_t p = pack(x[2 * i], x[2 * i + 1]); x[2 * i] = __uint_as_float(p); }
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) mx = fmaxf(fmaxf(mx, x[2 * i]), x[2 * i + 1]);
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] += mx * 1e-30f;
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) mx = fmaxf(fmaxf(mx, x[2 * i]), x[2 * i + 1]);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a0 = ex2(fmaf(x[2 * i], 0.01f, -mx));
        const float a1 = ex2(fmaf(x[2 * i + 1], 0.01f, -mx));
        sum += a0 + a1;
        pacc ^= pack(a0, a1);
        x[2 * i] = a0 + 1.0f; x[2 * i + 1] = a1 + 2.0f;
      }
      acc += sum;
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + mx + __uint_as_float(pacc);
}
template <int MODE>
void run(const char* name, int warps_per_smsp, double per_iter_elems) {
  float* o; cudaMalloc(&o, 148 * 1024 * 4);
  const int threads = warps_per_smsp * 4 * 32, iters = 2000;
  k<MODE><<<148, threads>>>(o, 10, 1.f); cudaDeviceSynchronize();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); k<MODE><<<148, threads>>>(o, iters, 1.f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double clk = ms * 1e-3 * 1.965e9;
  printf("%-8s warps/SMSP=%d: %.0f clk per warp-iteration (32 values/thread); per SMSP %.1f clk per warp-iteration-equivalent\n", name, warps_per_smsp, clk / iters, clk / iters / warps_per_smsp);
  cudaFree(o);
}
int main() {
  for (int w : {1, 4}) { run<0>("FFMA x32", w, 32); run<1>("MUFU x32", w, 32); run<2>("F2FP x16", w, 16); run<3>("max3 x16", w, 16); run<4>("mix", w, 32); }
  return 0;
}
