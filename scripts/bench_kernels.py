#!/usr/bin/env python
"""Micro-benchmarks of the individual C-ABI kernels at the shapes of the SD-1.4 GLIGEN forward (2B rows = 8).
CUDA-event timing, warm-up + N iterations back to back; prints one line per case and writes
gpurun_out/kernels_<tag>.json.   python scripts/bench_kernels.py [tag] [filter]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gligen_b200.ops import CudaOps, gn_scratch_floats  # noqa: E402

dev = "cuda:0"
ops = CudaOps(dev)
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
flt = sys.argv[2] if len(sys.argv) > 2 else ""
results = []


def rnd(*shape, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(dtype)


def timeit(name, fn, flops=0.0, nbytes=0.0, iters=20):
    """GPU time per call: `iters` calls captured into ONE CUDA graph and replayed - the python / ctypes issue time of a
    call (6-13 us, more than many of these kernels take) would otherwise be what a back-to-back loop measures."""
    if flt and flt not in name:
        return
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    r = dict(name=name, ms=ms, tflops=flops / ms / 1e9 if flops else None, gbs=nbytes / ms / 1e6 if nbytes else None)
    results.append(r)
    print(f"{name:58s} {ms*1e3:9.1f} us  " + (f"{r['tflops']:8.1f} TFLOP/s" if flops else "") + (f"  {r['gbs']:8.1f} GB/s" if nbytes else ""), flush=True)


Bt = 8
# ---------------- attention ----------------
for (d, T, G, heads) in ((40, 4096, 30, 8), (80, 1024, 30, 8), (160, 256, 30, 8), (160, 64, 30, 8)):
    C = heads * d
    qkv = rnd(Bt, T + G, 3 * C)
    out = torch.empty(Bt, T, C, device=dev, dtype=torch.bfloat16)
    kv = rnd(Bt, 77, 2 * C)
    qc = qkv[:, :T, :C].contiguous()
    for mode, mname in ((0, "auto"), (1, "mma_sync"), (2, "tcgen05"), (3, "short_tc"), (4, "tc2")):
        ops.lib.glg_debug_attn_mode(mode)
        if mode == 4 and d != 40:
            continue
        if mode != 3:
            timeit(f"attn self  d={d} T={T} [{mname}]", lambda: ops.attention(qkv[:, :T, :C], qkv[:, :T, C:2 * C], qkv[:, :T, 2 * C:], out, heads, d),
                   flops=4.0 * Bt * heads * T * T * d)
            timeit(f"attn fuser d={d} T={T}+{G} [{mname}]", lambda: ops.attention(qkv[:, :T, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], out, heads, d),
                   flops=4.0 * Bt * heads * T * (T + G) * d)
        if mode != 4:
          timeit(f"attn cross d={d} T={T}x77 [{mname}]", lambda: ops.attention(out, kv[:, :, :C], kv[:, :, C:], qc, heads, d),
               flops=4.0 * Bt * heads * T * 77 * d, nbytes=2.0 * Bt * T * C * 2)
    ops.lib.glg_debug_attn_mode(0)
    if d == 40:                      # share of the exp2 on the FMA pipe in the two-warpgroup kernel (pairs of 8)
        for poly in (0, 1, 2, 3):
            ops.lib.glg_debug_attn_poly_share(poly)
            timeit(f"attn self  d={d} T={T} [tc2 poly {poly}/8]", lambda: ops.attention(qkv[:, :T, :C], qkv[:, :T, C:2 * C], qkv[:, :T, 2 * C:], out, heads, d),
                   flops=4.0 * Bt * heads * T * T * d)
        ops.lib.glg_debug_attn_poly_share(0)

# ---------------- GEMMs (token GEMMs of the transformer blocks) ----------------
for cta2, bres, epi, cname in ((1, 1, 0, "stream-oldepi"), (1, 1, 1, "stream"), (1, 2, 1, "resident"), (2, 1, 1, "2cta"), (0, -1, -1, "auto")):
  ops.lib.glg_debug_gemm_cta2(cta2)
  ops.lib.glg_debug_gemm_bres(bres)
  ops.lib.glg_debug_gemm_epi(epi)
  for (T, C) in ((4096, 320), (1024, 640), (256, 1280), (64, 1280)):
      M = Bt * T
      x = rnd(M, C)
      res = rnd(M, C)
      bias = torch.randn(C, device=dev)
      xst = torch.zeros(C // 32, M, 2, device=dev)
      cs = torch.randn(3 * C, device=dev)
      for (nm, N, K, kw) in (("qkv", 3 * C, C, {}), ("qkv +lnfold", 3 * C, C, dict(ln=(xst, cs, 1e-5), bias=cs)), ("proj/out +bias+res", C, C, dict(bias=bias, residual=res)),
                             ("out +bias+res+stats", C, C, dict(bias=bias, residual=res, stats_out=xst)), ("ff2 +bias+res", C, 4 * C, dict(bias=bias, residual=res))):
          a = rnd(M, K)
          w = rnd(N, K, scale=K ** -0.5)
          o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
          timeit(f"gemm[{cname}] {nm:20s} M={M} N={N} K={K}", lambda: ops.gemm(a, w, o, **kw), flops=2.0 * M * N * K, nbytes=2.0 * (M * K + N * K + M * N))
      w1 = rnd(8 * C, C, scale=C ** -0.5)
      b1 = torch.randn(8 * C, device=dev)
      o1 = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
      timeit(f"gemm[{cname}] {'ff1 geglu':20s} M={M} N={8*C} K={C}", lambda: ops.gemm(x, w1, o1, bias=b1, geglu=True), flops=2.0 * M * 8 * C * C, nbytes=2.0 * (M * C + 8 * C * C + M * 4 * C))

ops.lib.glg_debug_gemm_cta2(0)
ops.lib.glg_debug_gemm_bres(0)
ops.lib.glg_debug_gemm_epi(-1)
for cta2, cname in ((1, "1cta"), (2, "2cta")):
    ops.lib.glg_debug_gemm_cta2(cta2)
    for (M, N, K) in ((4096, 4096, 4096), (8192, 8192, 8192)):
        a = rnd(M, K); w = rnd(N, K, scale=K ** -0.5); o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        timeit(f"gemm[{cname}] square M={M} N={N} K={K}", lambda: ops.gemm(a, w, o), flops=2.0 * M * N * K, nbytes=2.0 * (M * K + N * K + M * N), iters=5)
ops.lib.glg_debug_gemm_cta2(0)
# ---------------- conv3x3 ----------------
for (H, Cin, Cout) in ((64, 320, 320), (64, 960, 320), (64, 640, 320), (64, 640, 640), (32, 640, 640), (32, 1920, 640), (32, 1280, 1280), (16, 1280, 1280), (16, 2560, 1280), (8, 1280, 1280), (8, 2560, 1280)):
    a = rnd(Bt, H * H, Cin)
    w = rnd(9 * Cout, Cin, scale=(9 * Cin) ** -0.5)
    o = torch.empty(Bt, H * H, Cout, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(Cout, device=dev)
    res = rnd(Bt, H * H, Cout)
    for cta2, cname in ((1, "1cta"), (2, "2cta")):
        ops.lib.glg_debug_gemm_cta2(cta2)
        timeit(f"conv3x3[{cname}] {H}x{H} {Cin}->{Cout}", lambda: ops.gemm(a, w, o, bias=bias, residual=res, conv=(Bt, H, H)), flops=18.0 * Bt * H * H * Cin * Cout,
               nbytes=2.0 * (Bt * H * H * (Cin + 2 * Cout) + 9 * Cin * Cout))
    ops.lib.glg_debug_gemm_cta2(0)

# ---------------- norms ----------------
stats = torch.zeros(gn_scratch_floats(Bt), device=dev)
for (HW, C) in ((4096, 320), (4096, 960), (4096, 640), (1024, 640), (1024, 1920), (256, 1280), (256, 2560), (64, 2560)):
    x = rnd(Bt, HW, C)
    y = torch.empty_like(x)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    timeit(f"groupnorm+silu HW={HW} C={C}", lambda: ops.groupnorm(x, y, g, b, stats, 32, 1e-5, True), nbytes=2.0 * Bt * HW * C * 3)
for (T, C) in ((4096, 320), (1024, 640), (256, 1280)):
    x = rnd(Bt, T, C)
    y = torch.empty_like(x)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    timeit(f"layernorm T={T} C={C}", lambda: ops.layernorm(x, y, g, b), nbytes=2.0 * Bt * T * C * 2)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(results, open(os.path.join(ROOT, "gpurun_out", f"kernels_{tag}.json"), "w"), indent=1)
