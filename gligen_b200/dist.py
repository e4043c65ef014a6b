"""Multi-GPU plumbing: the sampling path shards by sample - one process per GPU, no collective
on the per-step path.  Collectives exist only at init (weight broadcast over NCCL/NVLink) and at the end
(optional latent gather).  Works with any torch.distributed backend (nccl on GPUs, gloo in CPU tests)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a global batch for `rank`; remainders go to the first ranks."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Slice every [B_total, ...] tensor of a dict (recursively) to this rank's samples.  Inputs (x_T, inpaint
    noise) must be drawn for the GLOBAL batch with one generator and then sliced, so an N-GPU run is row-wise
    identical to the 1-GPU run (shard-equivalence)."""
    out = {}
    for k, v in tensors.items():
        if isinstance(v, dict):
            out[k] = shard_batch(v, rank, world)
        elif isinstance(v, torch.Tensor) and v.dim() > 0:
            lo, hi = shard_range(v.shape[0], rank, world)
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def broadcast_module_weights(module: torch.nn.Module, src: int = 0, bucket_numel: int = 256 * 1024 * 1024) -> int:
    """Broadcast all parameters/buffers from `src` in flat buckets (one NCCL broadcast per <=1 GiB fp32
    bucket; 1.07 B parameters -> 4 calls).  Returns the number of elements sent."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    sent = 0
    i = 0
    while i < len(tensors):
        bucket, n = [], 0
        while i < len(tensors) and (n == 0 or n + tensors[i].numel() <= bucket_numel):
            bucket.append(tensors[i])
            n += tensors[i].numel()
            i += 1
        flat = torch.cat([t.reshape(-1).float() for t in bucket])
        dist.broadcast(flat, src=src)
        off = 0
        for t in bucket:
            t.copy_(flat[off: off + t.numel()].view_as(t))
            off += t.numel()
        sent += n
    return sent


def broadcast_tensors(tensors, src: int = 0, bucket_bytes: int = 512 << 20) -> int:
    """Broadcast a list of tensors in place, in list order, as flat per-dtype buckets of <= bucket_bytes (few large
    NCCL calls instead of one per tensor).  Every rank must pass tensors of identical shapes / dtypes.  Returns bytes sent."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    sent = 0
    by_dtype: Dict[torch.dtype, list] = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dtype, ts in by_dtype.items():
        i = 0
        while i < len(ts):
            bucket, nbytes = [], 0
            while i < len(ts) and (not bucket or nbytes + ts[i].numel() * ts[i].element_size() <= bucket_bytes):
                bucket.append(ts[i])
                nbytes += ts[i].numel() * ts[i].element_size()
                i += 1
            if len(bucket) == 1 and bucket[0].is_contiguous():
                dist.broadcast(bucket[0], src=src)
            else:
                flat = torch.cat([t.reshape(-1) for t in bucket])
                dist.broadcast(flat, src=src)
                off = 0
                for t in bucket:
                    t.copy_(flat[off: off + t.numel()].view_as(t))
                    off += t.numel()
            sent += nbytes
    return sent


def gather_latents(latent: torch.Tensor) -> torch.Tensor:
    """All-gather the per-rank [B/n, 4, H, W] latents into the global batch order (equal shards)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return latent
    parts = [torch.empty_like(latent) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, latent.contiguous())
    return torch.cat(parts, dim=0)
