#!/usr/bin/env python
"""Run the L0 self-attention shape a few times (target for ncu).  argv[1] = tc variant (1 pair, 0 auto)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gligen_b200.ops import CudaOps
dev = "cuda:0"; ops = CudaOps(dev)
ops.lib.glg_debug_attn_tc_variant(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
B, heads, d, T = 8, 8, 40, 4096
C = heads * d
qkv = (torch.randn(B, T, 3 * C, device=dev)).to(torch.bfloat16)
out = torch.empty(B, T, C, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], out, heads, d)
torch.cuda.synchronize()
