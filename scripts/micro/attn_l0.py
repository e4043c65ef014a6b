#!/usr/bin/env python
"""Level-0 self attention (d_head 40, 4096 tokens, 8 rows x 8 heads) through each tcgen05 kernel, timed by CUDA-graph replay.
argv[1:] = attention modes to time (default: 2 4 5); every mode is run with the FMA-pipe exp2 share 0, 1, 2."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gligen_b200.ops import CudaOps
dev = "cuda:0"; ops = CudaOps(dev)
modes = [int(a) for a in sys.argv[1:]] or [2, 4, 5]
Bt, heads, d, T, G = 8, 8, 40, 4096, 30
C = heads * d
qkv = (torch.randn(Bt, T + G, 3 * C, device=dev)).to(torch.bfloat16)
out = torch.empty(Bt, T, C, device=dev, dtype=torch.bfloat16)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


ref = None
for mode in modes:
    for poly, stag in ((0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (0, 600), (2, 600), (3, 600), (2, 1000), (3, 1000), (4, 600)):
        if mode != 5 and (poly > 2 or stag):
            continue
        if mode == 2 and poly:
            continue
        ops.lib.glg_debug_attn_tc3_stagger(stag)
        ops.lib.glg_debug_attn_mode(mode)
        ops.lib.glg_debug_attn_poly_share(poly)
        for name, Lk in (("self", T), ("fuser", T + G)):
            fn = lambda: ops.attention(qkv[:, :T, :C], qkv[:, :Lk, C:2 * C], qkv[:, :Lk, 2 * C:], out, heads, d)
            us = timeit(fn)
            fl = 4.0 * Bt * heads * T * Lk * d
            chk = ""
            if name == "self":
                if ref is None:
                    ref = out.float().clone()
                else:
                    chk = f"  rel-L2 vs first {((out.float() - ref).norm() / ref.norm()).item():.2e}"
            print(f"mode {mode} poly {poly}/8 stagger {stag:4d} {name:5s}: {us:7.1f} us  {fl / us / 1e6:6.1f} TFLOP/s{chk}", flush=True)
ops.lib.glg_debug_attn_mode(0); ops.lib.glg_debug_attn_poly_share(0)
if os.environ.get("VARS"):
    ops.lib.glg_debug_attn_mode(5); ops.lib.glg_debug_attn_poly_share(2)
    ref2 = None
    for var in [int(v) for v in os.environ.get("VARLIST", "0,1,2,3,5,7,0").split(",")]:
        ops.lib.glg_debug_attn_tc3_variant(var)
        for name, Lk in (("self", T), ("fuser", T + G)):
            us = timeit(lambda: ops.attention(qkv[:, :T, :C], qkv[:, :Lk, C:2 * C], qkv[:, :Lk, 2 * C:], out, heads, d))
            chk = ""
            if name == "self":
                ops.attention(qkv[:, :T, :C], qkv[:, :T, C:2 * C], qkv[:, :T, 2 * C:], out, heads, d); torch.cuda.synchronize()
                if ref2 is None:
                    ref2 = out.float().clone()
                chk = f"  rel-L2 vs variant 0 {((out.float() - ref2).norm() / ref2.norm()).item():.2e}"
            print(f"tc3 poly 2/8 variant {var} {name:5s}: {us:7.1f} us  {4.0 * Bt * heads * T * Lk * d / us / 1e6:6.1f} TFLOP/s{chk}", flush=True)
    ops.lib.glg_debug_attn_tc3_variant(2); ops.lib.glg_debug_attn_mode(0); ops.lib.glg_debug_attn_poly_share(0)
if os.environ.get("KO"):
    ops.lib.glg_debug_attn_mode(5)
    for ko in (0, 1, 2, 4, 8, 12, 14, 16, 17, 49):
        ops.lib.glg_debug_attn_tc3_knockout(ko)
        us = timeit(lambda: ops.attention(qkv[:, :T, :C], qkv[:, :T, C:2 * C], qkv[:, :T, 2 * C:], out, heads, d))
        print(f"tc3 knock-out {ko:2d}: {us:7.1f} us", flush=True)
    ops.lib.glg_debug_attn_tc3_knockout(0); ops.lib.glg_debug_attn_mode(0)
if os.environ.get("PROBE"):
    buf = torch.zeros(32 * 8, device=dev, dtype=torch.int64)
    ops.lib.glg_debug_attn_mode(5); ops.lib.glg_debug_attn_tc3_knockout(64); ops.lib.glg_debug_attn_probe(buf.data_ptr())
    for _ in range(3):
        buf.zero_()
        ops.attention(qkv[:, :T, :C], qkv[:, :T, C:2 * C], qkv[:, :T, 2 * C:], out, heads, d)
        torch.cuda.synchronize()
    ops.lib.glg_debug_attn_probe(None); ops.lib.glg_debug_attn_tc3_knockout(0); ops.lib.glg_debug_attn_mode(0)
    b = buf.view(32, 8).cpu().tolist()
    print("clock sums of CTA (0,0,0), 64 key tiles; softmax: wait_S | ld+free | max(+wait P free, rescale) | exp+st | st_wait+arrive | tail")
    print("  TMA+ones warp 0: wait_empty %d issue %d ones %d tail %d" % (b[0][0], b[0][1], b[0][2], b[0][5]))
    for w in (1, 2, 3):
        print("  issuer warp %d: wait_S_free %d QK(+wait K) %d wait_V+P %d PV %d tail %d" % (w, b[w][0], b[w][1], b[w][2], b[w][3], b[w][5]))
    for w in range(4, 16):
        print(f"  softmax warp {w} (wg {(w - 4) // 4}):", " | ".join(str(v) for v in b[w][:6]))
