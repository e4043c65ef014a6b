"""CPU fp32 ORACLE for the spatial grounding modalities (hed / canny / depth / normal / sem).  TEST INFRASTRUCTURE ONLY.

Functional restatement (plain torch fp32, state-dict in / tensor out) of
  * the ConvNeXt-tiny grounding tokenizer  - ldm/modules/diffusionmodules/convnext.py:14-110 and
    {hed,canny,depth,normal,sem}_grounding_net.py (PositionNet.forward :38-63),
  * the grounding downsamplers              - {hed,canny,depth,normal,sem}_grounding_downsampler.py,
  * and how UNetModel.forward consumes them - openaimodel.py:436-443.
Pinned by executing the unmodified reference (oracle/gen_golden_spatial.py: strict state-dict load into the reference modules,
outputs compared, fixtures under tests/golden/spatial_*.pt).  Only tests/, __graft_entry__ and bench.py may import this file.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from gligen_b200.spec import CONVNEXT_TINY_DEPTHS, CONVNEXT_TINY_DIMS, SPATIAL_MAP_KEY, UNetConfig


def _ln_channels_first(x, w, b, eps=1e-6):
    """convnext.py:135-139: LayerNorm over dim 1 of an NCHW tensor (biased variance, eps inside the sqrt)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None, None] * x + b[:, None, None]


def convnext_block(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor) -> torch.Tensor:
    """convnext.py:38-52 Block.forward (drop_path = 0 -> identity)."""
    c = x.shape[1]
    h = F.conv2d(x, sd[f"{p}.dwconv.weight"], sd[f"{p}.dwconv.bias"], padding=3, groups=c)
    h = h.permute(0, 2, 3, 1)
    h = F.layer_norm(h, (c,), sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-6)
    h = F.linear(h, sd[f"{p}.pwconv1.weight"], sd[f"{p}.pwconv1.bias"])
    h = F.gelu(h)
    h = F.linear(h, sd[f"{p}.pwconv2.weight"], sd[f"{p}.pwconv2.bias"])
    h = sd[f"{p}.gamma"] * h
    return x + h.permute(0, 3, 1, 2)


def convnext_tiny(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor) -> torch.Tensor:
    """convnext.py:102-110 forward_features: [B, 3, H, W] -> [B, 768, H / 32, W / 32] (no final norm, no head)."""
    d = f"{p}.downsample_layers"
    for i in range(4):
        if i == 0:
            x = F.conv2d(x, sd[f"{d}.0.0.weight"], sd[f"{d}.0.0.bias"], stride=4)
            x = _ln_channels_first(x, sd[f"{d}.0.1.weight"], sd[f"{d}.0.1.bias"])
        else:
            x = _ln_channels_first(x, sd[f"{d}.{i}.0.weight"], sd[f"{d}.{i}.0.bias"])
            x = F.conv2d(x, sd[f"{d}.{i}.1.weight"], sd[f"{d}.{i}.1.bias"], stride=2)
        for j in range(CONVNEXT_TINY_DEPTHS[i]):
            x = convnext_block(sd, f"{p}.stages.{i}.{j}", x)
    return x


def position_net_spatial(cfg: UNetConfig, sd: Dict[str, torch.Tensor], g: Dict[str, torch.Tensor]) -> torch.Tensor:
    """hed_grounding_net.py:38-63 (canny / depth / normal identical; sem: nearest resize then in_conv, sem_grounding_net.py:40-47)
    -> objs [B, (resize / 32)^2, out_dim]."""
    pn = "position_net"
    x, mask = g[SPATIAL_MAP_KEY[cfg.tokenizer]].float(), g["mask"].float()
    B = x.shape[0]
    x = F.interpolate(x, cfg.tok_resize)                           # mode defaults to nearest
    if cfg.tokenizer == "sem":
        x = F.conv2d(x, sd[f"{pn}.in_conv.weight"], sd[f"{pn}.in_conv.bias"], padding=1)
    feat = convnext_tiny(sd, f"{pn}.convnext_tiny_backbone", x)
    n = cfg.spatial_tokens
    objs = feat.reshape(B, -1, n).permute(0, 2, 1)
    null = sd[f"{pn}.null_feature"].view(1, 1, -1).repeat(B, n, 1)
    m = mask.view(-1, 1, 1)
    objs = objs * m + null * (1 - m)
    objs = objs + sd[f"{pn}.pos_embedding"]
    h = F.silu(F.linear(objs, sd[f"{pn}.linears.0.weight"], sd[f"{pn}.linears.0.bias"]))
    h = F.silu(F.linear(h, sd[f"{pn}.linears.2.weight"], sd[f"{pn}.linears.2.bias"]))
    return F.linear(h, sd[f"{pn}.linears.4.weight"], sd[f"{pn}.linears.4.bias"])


def grounding_downsampler(cfg: UNetConfig, sd: Dict[str, torch.Tensor], extra: torch.Tensor, latent: int) -> torch.Tensor:
    """GroundingDownsampler.forward -> [B, ds_out_dim, latent, latent] (the reference hard-codes 64 = the SD latent size for hed,
    hed_grounding_downsampler.py:19; the conv stacks reach it as resize_input / 4).
    hed: channel 0, bicubic to the latent size.  canny / depth: channel 0, bicubic to resize_input, Conv(1,4,4,2,1)-SiLU-Conv(4,out,4,2,1).
    normal: 3 channels, same stack.  sem: nearest to resize_input, Conv(in_dim,16,4,2,1)-SiLU-Conv(16,out,4,2,1)."""
    x = extra.float()
    if cfg.tokenizer == "hed":
        return F.interpolate(x[:, 0].unsqueeze(1), (latent, latent), mode="bicubic")
    r = cfg.ds_resize
    if cfg.tokenizer in ("canny", "depth"):
        x = F.interpolate(x[:, 0].unsqueeze(1), (r, r), mode="bicubic")
    elif cfg.tokenizer == "normal":
        x = F.interpolate(x, (r, r), mode="bicubic")
    else:
        x = F.interpolate(x, (r, r), mode="nearest")
    p = "downsample_net.layers"
    x = F.silu(F.conv2d(x, sd[f"{p}.0.weight"], sd[f"{p}.0.bias"], stride=2, padding=1))
    return F.conv2d(x, sd[f"{p}.2.weight"], sd[f"{p}.2.bias"], stride=2, padding=1)
