// Shared device-side helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM, small math.
// Everything here is inline PTX written for sm_100a (no CUTLASS dependency).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace glg {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// With GLG_PDL=1 every kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel
// in the stream may be scheduled (and run its prologue: barrier init, TMEM alloc, descriptor prefetch) while this
// one drains; pdl_wait() blocks until the previous grid has completed and its memory is visible, so it must precede
// the first access to global data.  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
// non-suspending poll (mbarrier.test_wait): the thread keeps its issue slot instead of being parked by the hardware
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Wait with a watchdog: a protocol bug traps (-> cudaErrorLaunchFailure at the next sync) instead of
// hanging the GPU box.  The watchdog only runs on the slow path.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (int i = 0; i < 256; ++i)           // common case: the phase completes within a few (HW-suspended) polls - no timer read on this path
    if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  while (true) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i)          // the timer read is slow: keep it off the wake-up path
      if (mbar_try_wait(bar, parity)) return;
    if (globaltimer_ns() - t0 > 4000000000ull) {  // 4 s
      printf("glg: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// Busy-polling wait (no hardware suspend between polls): lowest wake-up latency, at the price of issue slots.  Falls back to the
// suspending wait (with its watchdog) after 4096 polls.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (int i = 0; i < 4096; ++i)
    if (mbar_test_wait(bar, parity)) return;
  mbar_wait(bar, parity);
}

// One leader lane of a fully converged warp.  The single-thread roles (TMA producer, MMA issuer) run the whole warp
// through their loops and predicate the issue on this flag: a role entered as `if (lane == 0)` makes the compiler wrap
// every UTCHMMA / UTCBAR / UTMALDG in an ELECT + BRA.U.ANY loop over "the active lanes" (~6 slow uniform-datapath
// instructions each); with elect.sync the same instructions are emitted back to back, predicated.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ TMA (tiled tensor maps)
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the A operand read from TMEM (128 lanes x K bf16 packed two per 32-bit column).
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128B swizzle, rows of 64 bf16 (=128 B),
// 8-row core-matrix groups 1024 B apart (SBO).  (cute::UMMA::SmemDescriptor, version 1 = Blackwell.)
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= (uint64_t)1 << 16;                              // LBO (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                    // SBO = 1024 B  [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version = 1
  d |= (uint64_t)2 << 61;                              // layout type = SWIZZLE_128B
  return d;
}
// MN-major operand (e.g. V[key][d] used as B[K=key][N=d]): rows of 64 bf16 along MN (128 B), one row per K
// index, 8-row groups 1024 B apart (SBO); further 64-wide MN atoms follow at LBO bytes.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes = 16) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;    // LBO: byte offset between 64-element MN atoms
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16: BF16 x BF16 -> F32, both operands K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool b_mn_major = false) {
  return (b_mn_major ? (1u << 16) : 0u)   // B major: 0 = K, 1 = MN
       | (1u << 4)                      // C format = F32
       | (1u << 7)                      // A format = BF16
       | (1u << 10)                     // B format = BF16
       | ((uint32_t)(N >> 3) << 17)     // N / 8
       | ((uint32_t)(M >> 4) << 24);    // M / 16
}

__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {     // one 32-bit column of this thread's lane
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
// TMEM -> registers: 32 lanes x 32 columns of 32-bit; thread i of the warp receives lane (base+i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ math
// x * sigmoid(x); __fdividef: 2 ulp, no IEEE-division slow path (a CALL per element in the epilogues)
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// CLIP's activation (transformers "quick_gelu": x * sigmoid(1.702 x))
__device__ __forceinline__ float quick_gelu_f(float x) { return __fdividef(x, 1.0f + __expf(-1.702f * x)); }
// erf(z) ~= z P(z^2) / Q(z^2) on |z| <= 4 (clamped; |erf(4)-1| < 2e-8): own least-squares rational fit,
// max abs error 3.3e-7 in fp32 (checked against scipy.special.erf), branch-free: 11 FMA + 1 rcp.
__device__ __forceinline__ float erf_rational(float z) {
  z = fminf(fmaxf(z, -4.0f), 4.0f);
  const float z2 = z * z;
  float pn = 2.0269792457838776e-06f;
  pn = fmaf(pn, z2, 0.0002861879765987396f);
  pn = fmaf(pn, z2, 0.003845315193757415f);
  pn = fmaf(pn, z2, 0.05298357829451561f);
  pn = fmaf(pn, z2, 0.1923242062330246f);
  pn = fmaf(pn, z2, 1.128378987312317f);
  float qn = 3.7925383367110044e-05f;
  qn = fmaf(qn, z2, 0.0011811800068244338f);
  qn = fmaf(qn, z2, 0.015125652775168419f);
  qn = fmaf(qn, z2, 0.11488588154315948f);
  qn = fmaf(qn, z2, 0.5037747621536255f);
  qn = fmaf(qn, z2, 1.0f);
  return __fdividef(z * pn, qn);
}
// exact (erf) GELU of attention.py:44 (F.gelu default); abs error < 1e-6, far below bf16 resolution
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erf_rational(x * 0.70710678118654752f)); }
// Cheaper erf for the GEGLU epilogue (the FF1 GEMMs are bound by the epilogue's instruction issue, not by the tensor
// pipe: 128 x 128 GELUs per 2560 MMA cycles): z P3(z^2) / Q3(z^2) on |z| <= 3.2 (clamped; 1 - erf(3.2) = 6e-6), own
// least-squares fit: max abs error 3.4e-6 on erf, 1.5e-5 on gelu(x) over all x - three decimal orders below the bf16
// resolution of the stored product.  7 FMA + 1 rcp instead of 11 FMA + 1 rcp.
__device__ __forceinline__ float erf_rational3(float z) {
  z = fminf(fmaxf(z, -3.2f), 3.2f);
  const float z2 = z * z;
  float pn = 0.0007654472137801349f;
  pn = fmaf(pn, z2, 0.04346451908349991f);
  pn = fmaf(pn, z2, 0.15304264426231384f);
  pn = fmaf(pn, z2, 1.1283873319625854f);
  float qn = 0.009417801164090633f;
  qn = fmaf(qn, z2, 0.09465143829584122f);
  qn = fmaf(qn, z2, 0.4690375328063965f);
  qn = fmaf(qn, z2, 1.0f);
  return __fdividef(z * pn, qn);
}
// x * gelu_erf(g) for the GEGLU epilogue: 0.5 g (1 + erf(g / sqrt 2)) as one FMA on h = 0.5 g
__device__ __forceinline__ float geglu_f(float x, float g) {
  const float h = 0.5f * g;
  return x * fmaf(h, erf_rational3(g * 0.70710678118654752f), h);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace glg
