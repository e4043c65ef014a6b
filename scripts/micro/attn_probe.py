#!/usr/bin/env python
"""Phase breakdown of the d_head=40 tcgen05 attention kernel (clock64 stamps of two softmax warps per CTA)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gligen_b200.ops import CudaOps
dev = "cuda:0"; ops = CudaOps(dev)
B, heads, d, T = 8, 8, 40, 4096
C = heads * d
qkv = (torch.randn(B, T, 3 * C, device=dev)).to(torch.bfloat16)
out = torch.empty(B, T, C, device=dev, dtype=torch.bfloat16)
run = lambda: ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], out, heads, d)
def timeit(n=10):
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print(f"s1 kernel: {timeit():.1f} us")
for kb, nc in ((0, 4), (20, 3), (60, 2), (120, 1)):
    ops.lib.glg_debug_attn_poly(kb * 1024 if kb else -1)
    t = timeit()
    print(f"  s1 with {nc} CTA/SM: {t:.1f} us -> {t * 1e-6 * 1.965e9 * nc * 148 / (2048 * 64):.0f} clk per CTA-iteration")
ops.lib.glg_debug_attn_poly(-1)
for flags, nm in ((32, "dbg build, nothing off"), (1, "no MUFU"), (2, "no QK mma"), (4, "no PV mma"), (6, "no mma"), (16, "no K/V TMA"), (7, "no MUFU, no mma"), (23, "no MUFU/mma/TMA")):
    ops.lib.glg_debug_attn_poly(flags)
    print(f"  s1 [{nm}]: {timeit():.1f} us")
ops.lib.glg_debug_attn_poly(0)
ops.lib.glg_debug_attn_tc_variant(1)
for flags, nm in ((32, "dbg build, nothing off"), (1, "no MUFU"), (2, "no QK mma"), (4, "no PV mma"), (16, "no K/V TMA"), (17, "no TMA, no MUFU"), (23, "no TMA/MUFU/mma")):
    ops.lib.glg_debug_attn_poly(flags)
    print(f"  pair [{nm}]: {timeit():.1f} us")
ops.lib.glg_debug_attn_poly(0)
print(f"pair kernel: {timeit():.1f} us")
ref = out.clone()
ops.lib.glg_debug_attn_tc_variant(0)
run(); torch.cuda.synchronize()
print("max |s1 - pair| =", (out.float() - ref.float()).abs().max().item())
nct = (T // 128) * heads * B
nkt = T // 64
def probe(variant, names, label):
    buf = torch.zeros(nct * 2 * 8, dtype=torch.int64, device=dev)
    ops.lib.glg_debug_attn_tc_variant(variant)
    ops.lib.glg_debug_attn_probe(buf.data_ptr())
    print(f"{label} probe build: {timeit():.1f} us")
    buf.zero_(); run(); torch.cuda.synchronize()
    ops.lib.glg_debug_attn_probe(None)
    ops.lib.glg_debug_attn_tc_variant(0)
    st = buf.view(nct, 2, 8).double()
    for w in range(2):
        m = st[:, w].mean(0) / nkt
        print(f"  warp {w}: " + "  ".join(f"{n}={v:.0f}" for n, v in zip(names, m.tolist())) + f"  | total/iter={m.sum():.0f} clk")
probe(1, ["wait s_full", "tmem ld S", "max+exchange", "wait pv_done", "rescale", "exp loop", "st P+arrive", "-"], "pair")
def trace(extra_kb, label):
    buf = torch.zeros(8 * 16, dtype=torch.int64, device=dev)
    ops.lib.glg_debug_attn_poly(extra_kb * 1024 if extra_kb else -1)
    ops.lib.glg_debug_attn_probe(buf.data_ptr())
    run(); torch.cuda.synchronize()
    t = timeit()
    buf.zero_(); run(); torch.cuda.synchronize()
    ops.lib.glg_debug_attn_probe(None); ops.lib.glg_debug_attn_poly(-1)
    ev = buf.view(8, 16).cpu()
    t0 = int(ev[0, 8])
    print(f"{label}: {t:.1f} us; event trace of CTA 0 (clk rel. to MMA kv_full wake of iteration 8)")
    names = {8: "M kv_full ok", 9: "M QK issued+commit", 0: "S s_full seen", 1: "S ldB+max", 2: "S ldA+max", 3: "S rescale chk", 4: "S exps A", 5: "S stA ldB expsB", 6: "S stB waited", 7: "S arrived", 10: "M p_full seen", 11: "M PV issued+commit"}
    for j in range(3):
        print("  iter", 8 + j, " ".join(f"[{names[k]} {int(ev[j, k]) - t0}]" for k in (8, 9, 0, 1, 2, 3, 4, 5, 6, 7, 10, 11)))

#probe(0, ["wait s_full", "ld B+max", "ld A+max", "rescale", "exps A", "st A, ld B, exps B", "st B+wait", "arrive"], "s1")
