"""CPU fp32 ORACLE for the CLIP text encoder behind FrozenCLIPEmbedder.  TEST INFRASTRUCTURE ONLY.

The reference (ldm/modules/encoders/modules.py:144-173) calls `transformers.CLIPTextModel` (third-party, pinned
transformers==4.19.2 by env_docker/Dockerfile:3, not vendored under /root/reference).  This file restates that published algorithm
(transformers/models/clip/modeling_clip.py: CLIPTextEmbeddings, CLIPEncoderLayer, CLIPAttention with the causal mask,
CLIPMLP with quick_gelu, final_layer_norm, pooled = hidden state at argmax(input_ids)) in plain torch fp32, state-dict in / tensors
out.  Pinned by oracle/gen_golden_clip.py against the INSTALLED transformers' CLIPTextModel on the same seeded weights
(tests/golden/clip_text_*.pt).  Only tests/, __graft_entry__ and bench.py may import this file."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F


def clip_text_forward(cfg, sd: Dict[str, torch.Tensor], input_ids: torch.Tensor, prefix: str = "transformer.text_model.") -> Tuple[torch.Tensor, torch.Tensor]:
    B, L = input_ids.shape
    C, H = cfg.width, cfg.heads
    d = C // H
    x = sd[prefix + "embeddings.token_embedding.weight"][input_ids] + sd[prefix + "embeddings.position_embedding.weight"][:L][None]
    mask = torch.full((L, L), float("-inf")).triu(1)                      # causal: row i sees keys [0, i]
    for i in range(cfg.layers):
        l = f"{prefix}encoder.layers.{i}"
        h = F.layer_norm(x, (C,), sd[f"{l}.layer_norm1.weight"], sd[f"{l}.layer_norm1.bias"], cfg.eps)
        q = F.linear(h, sd[f"{l}.self_attn.q_proj.weight"], sd[f"{l}.self_attn.q_proj.bias"]) * d ** -0.5
        k = F.linear(h, sd[f"{l}.self_attn.k_proj.weight"], sd[f"{l}.self_attn.k_proj.bias"])
        v = F.linear(h, sd[f"{l}.self_attn.v_proj.weight"], sd[f"{l}.self_attn.v_proj.bias"])
        q, k, v = (t.view(B, L, H, d).transpose(1, 2) for t in (q, k, v))
        a = torch.softmax(q @ k.transpose(-1, -2) + mask, dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, L, C)
        x = x + F.linear(a, sd[f"{l}.self_attn.out_proj.weight"], sd[f"{l}.self_attn.out_proj.bias"])
        h = F.layer_norm(x, (C,), sd[f"{l}.layer_norm2.weight"], sd[f"{l}.layer_norm2.bias"], cfg.eps)
        h = F.linear(h, sd[f"{l}.mlp.fc1.weight"], sd[f"{l}.mlp.fc1.bias"])
        h = h * torch.sigmoid(1.702 * h)                                   # quick_gelu
        x = x + F.linear(h, sd[f"{l}.mlp.fc2.weight"], sd[f"{l}.mlp.fc2.bias"])
    z = F.layer_norm(x, (C,), sd[prefix + "final_layer_norm.weight"], sd[prefix + "final_layer_norm.bias"], cfg.eps)
    pooled = z[torch.arange(B), input_ids.argmax(dim=-1)]
    return z, pooled
