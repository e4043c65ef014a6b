#!/usr/bin/env python
"""Pipeline probes of the d_head = 40 tcgen05 attention kernel at the L0 shape (8 x 8 heads x 4096 x 4096).

Uses the PROBE instantiation of attn_tc_kernel<48> (test hooks glg_debug_attn_probe / _poly / _tc_variant):
  * knock-outs: the kernel with MUFU / QK^T MMAs / P.V MMAs / K,V TMA loads removed (results are wrong, time is real)
  * phase stamps: clock64 deltas accumulated by two softmax warps of every CTA, averaged per key tile
  * event trace: absolute clock64 of the TMA, MMA and one softmax warp of CTA 0 for key tiles 8..15
Output of the B200 run is summarised in profiles/r1_attention_pipeline.md.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gligen_b200.ops import CudaOps  # noqa: E402

dev = "cuda:0"
ops = CudaOps(dev)
B, heads, d, T = 8, 8, 40, 4096
C = heads * d
qkv = (torch.randn(B, T, 3 * C, device=dev)).to(torch.bfloat16)
out = torch.empty(B, T, C, device=dev, dtype=torch.bfloat16)
nct = (T // 128) * heads * B
nkt = T // 64


def run():
    ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], out, heads, d)


def timeit(n=10):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print(f"production kernel (ones-column row sum): {timeit():.1f} us")
ops.lib.glg_debug_attn_tc_variant(3)
print(f"production kernel (softmax-side row sum): {timeit():.1f} us")
ops.lib.glg_debug_attn_tc_variant(1)
for flags, nm in ((32, "nothing off"), (1, "no MUFU"), (2, "no QK mma"), (4, "no PV mma"), (16, "no K/V TMA"), (23, "no MUFU/mma/TMA")):
    ops.lib.glg_debug_attn_poly(flags)
    print(f"  PROBE build [{nm}]: {timeit():.1f} us")
ops.lib.glg_debug_attn_poly(0)

PHASES = ["loop top", "mbar_wait s_full", "fence_after", "ld issue", "ld wait", "max+sts", "bar.sync", "lds+vote", "rescale",
          "exp loop", "st issue", "st wait", "fence_before", "syncwarp", "arrive", "-"]


def phases(flags, label):
    buf = torch.zeros(nct * 2 * 16, dtype=torch.int64, device=dev)
    ops.lib.glg_debug_attn_poly(flags)
    ops.lib.glg_debug_attn_probe(buf.data_ptr())
    t = timeit()
    buf.zero_()
    run()
    torch.cuda.synchronize()
    ops.lib.glg_debug_attn_probe(None)
    ops.lib.glg_debug_attn_poly(0)
    st = buf.view(nct, 2, 16).double()
    print(f"phase stamps [{label}] ({t:.1f} us with stamps; each stamp costs ~20 clk):")
    for w in range(2):
        m = st[:, w].mean(0) / nkt
        print(f"  warp {2 if w == 0 else 7}: " + "  ".join(f"{n}={v:.0f}" for n, v in zip(PHASES, m.tolist())) + f"  | total/tile={m.sum():.0f} clk")


def trace():
    buf = torch.zeros(8 * 16, dtype=torch.int64, device=dev)
    ops.lib.glg_debug_attn_poly(64)
    ops.lib.glg_debug_attn_probe(buf.data_ptr())
    run()
    torch.cuda.synchronize()
    buf.zero_()
    run()
    torch.cuda.synchronize()
    ops.lib.glg_debug_attn_probe(None)
    ops.lib.glg_debug_attn_poly(0)
    ev = buf.view(8, 16).cpu()
    t0 = int(ev[0, 5])
    names = ["TMA kv_empty ok->load", "MMA kv_full ok", "MMA QK issued", "MMA p_full seen", "MMA PV issued", "SM s_full seen", "SM arrived p_full"]
    print("event trace, CTA 0; columns = key tile 8..15; clk relative to softmax s_full(8) seen")
    for k, n in enumerate(names):
        print(f"  {n:24s}" + " ".join(f"{int(ev[j, k]) - t0:7d}" for j in range(8)))


phases(0, "all on")
phases(1, "no MUFU")
trace()
ops.lib.glg_debug_attn_tc_variant(0)
