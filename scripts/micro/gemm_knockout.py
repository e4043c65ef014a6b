#!/usr/bin/env python
"""Where does a short-K token GEMM spend its time?  Knock-outs of gemm_tc_kernel (time is real, results are not):
1 = no TMA operand loads, 2 = no MMAs, 4 = no epilogue work (barrier protocol only), 8 = no output stores."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gligen_b200.ops import CudaOps
dev = "cuda:0"; ops = CudaOps(dev)


def timeit(fn, iters=30):
    """GPU time: the calls are captured into one CUDA graph (a python-driven loop measures the ~10 us host issue time)."""
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
    return e0.elapsed_time(e1) / iters * 1e3


def cases(M, C):
    x = torch.randn(M, C, device=dev).to(torch.bfloat16)
    res = torch.randn(M, C, device=dev).to(torch.bfloat16)
    bias = torch.randn(C, device=dev)
    xst = torch.rand(C // 32, M, 2, device=dev)
    w = (torch.randn(C, C, device=dev) * C ** -0.5).to(torch.bfloat16); o = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    w3 = (torch.randn(3 * C, C, device=dev) * C ** -0.5).to(torch.bfloat16); o3 = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
    cs = torch.randn(3 * C, device=dev)
    w8 = (torch.randn(8 * C, C, device=dev) * C ** -0.5).to(torch.bfloat16); o4 = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
    b8 = torch.randn(8 * C, device=dev)
    a4 = torch.randn(M, 4 * C, device=dev).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, device=dev) * (4 * C) ** -0.5).to(torch.bfloat16)
    return {
        "plain  NxK=CxC": lambda: ops.gemm(x, w, o),
        "cxc +bias+res+stats": lambda: ops.gemm(x, w, o, bias=bias, residual=res, stats_out=xst),
        "qkv plain": lambda: ops.gemm(x, w3, o3),
        "qkv +lnfold": lambda: ops.gemm(x, w3, o3, bias=cs, ln=(xst, cs, 1e-5)),
        "ff1 geglu+ln": lambda: ops.gemm(x, w8, o4, bias=b8, geglu=True, ln=(xst, b8, 1e-5)),
        "ff2 +bias+res": lambda: ops.gemm(a4, w2, o, bias=bias, residual=res),
    }


for (M, C) in ((32768, 320), (8192, 640)):
    print(f"--- M={M} C={C}")
    cs_ = cases(M, C)
    print(f"{'case':24s}" + "".join(f"{n:>12s}" for n in ("full", "noTMA", "noMMA", "noEpi", "noStore", "noTMA+MMA", "noT+M+E", "empty-knl")))
    for name, fn in cs_.items():
        row = []
        for flags in (0, 1, 2, 4, 8, 3, 7):
            ops.lib.glg_debug_gemm_knockout(flags)
            row.append(timeit(fn))
        ops.lib.glg_debug_gemm_knockout(0)
        print(f"{name:24s}" + "".join(f"{t:12.1f}" for t in row), flush=True)
# launch-overhead yardstick: a trivial kernel through the same python path
t = torch.zeros(8, device=dev)
y = torch.zeros(8, device=dev, dtype=torch.bfloat16)
print("cast 8 elements (python + launch floor): %.1f us" % timeit(lambda: ops.cast(t, y)))
