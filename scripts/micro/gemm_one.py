#!/usr/bin/env python
"""A few launches of ONE token-GEMM shape of the level-0 transformer block (target for ncu).
argv[1] = case: cxc | qkv | qkvln | ff1 | ff2 ; argv[2] (optional) = GLG epilogue mode (0 old, 1 TMA)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gligen_b200.ops import CudaOps
dev = "cuda:0"; ops = CudaOps(dev)
case = sys.argv[1] if len(sys.argv) > 1 else "cxc"
if len(sys.argv) > 2:
    ops.lib.glg_debug_gemm_epi(int(sys.argv[2]))
M, C = 32768, 320
x = torch.randn(M, C, device=dev).to(torch.bfloat16)
res = torch.randn(M, C, device=dev).to(torch.bfloat16)
bias = torch.randn(C, device=dev)
xst = torch.rand(C // 32, M, 2, device=dev)
if case == "cxc":
    w = (torch.randn(C, C, device=dev) * C ** -0.5).to(torch.bfloat16); o = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(x, w, o, bias=bias, residual=res, stats_out=xst)
elif case in ("qkv", "qkvln"):
    w = (torch.randn(3 * C, C, device=dev) * C ** -0.5).to(torch.bfloat16); o = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
    cs = torch.randn(3 * C, device=dev)
    fn = (lambda: ops.gemm(x, w, o, bias=cs, ln=(xst, cs, 1e-5))) if case == "qkvln" else (lambda: ops.gemm(x, w, o))
elif case == "ff1":
    w = (torch.randn(8 * C, C, device=dev) * C ** -0.5).to(torch.bfloat16); o = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
    b1 = torch.randn(8 * C, device=dev)
    fn = lambda: ops.gemm(x, w, o, bias=b1, geglu=True, ln=(xst, b1, 1e-5))
else:
    a = torch.randn(M, 4 * C, device=dev).to(torch.bfloat16)
    w = (torch.randn(C, 4 * C, device=dev) * (4 * C) ** -0.5).to(torch.bfloat16); o = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, o, bias=bias, residual=res)
for _ in range(5):
    fn()
torch.cuda.synchronize()
