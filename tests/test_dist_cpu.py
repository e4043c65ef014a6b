"""CPU, world_size 2, gloo: the multi-GPU path is sample sharding with no per-step collective.
Checks shard bookkeeping, the init-time weight broadcast, the latent gather, and shard-equivalence of the
engine plan (rank r's eps == rows of the single-process eps)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gligen_b200 import synth
    from gligen_b200.dist import broadcast_module_weights, gather_latents, shard_batch, shard_range
    from gligen_b200.engine import Engine
    from gligen_b200.spec import NAMED_CONFIGS, synthetic_state_dict
    from ref_ops import RefOps
    from test_abi_cpu import _tiny_model
    torch.set_num_threads(2)
    cfg = NAMED_CONFIGS["tiny"]
    model = _tiny_model()
    if rank == 0:
        model.load_state_dict(synthetic_state_dict(cfg, 0))
    sent = broadcast_module_weights(model, src=0, bucket_numel=8 * 1024 * 1024)
    assert sent == sum(p.numel() for p in model.parameters())
    sd = synthetic_state_dict(cfg, 0)
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), k
    eng = Engine(cfg, RefOps())
    eng.load_state_dict(model.state_dict())
    full = synth.make_inputs(cfg, 4, 6, seed=7)                 # the GLOBAL batch, same generator on every rank
    mine = shard_batch(full, rank, world)
    lo, hi = shard_range(4, rank, world)
    assert (lo, hi) == (2 * rank, 2 * rank + 2) and mine["x"].shape[0] == 2
    ts = torch.tensor([981, 501, 21, 1])
    e = eng.forward(mine["x"], ts[lo:hi], mine["context"], mine["grounding_input"])
    allv = gather_latents(e)
    if rank == 0:
        ref = eng.forward(full["x"], ts, full["context"], full["grounding_input"])
        torch.save({"gathered": allv, "single": ref}, os.path.join(tmp, "out.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_equivalence_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    out = torch.load(os.path.join(tmp_path, "out.pt"))
    assert out["gathered"].shape == out["single"].shape == (4, 4, 16, 16)
    assert (out["gathered"] - out["single"]).abs().max() < 2e-5


def _worker_spatial(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gligen_b200 import synth
    from gligen_b200.dist import gather_latents, shard_batch, shard_range
    from gligen_b200.engine import Engine
    from gligen_b200.spec import NAMED_CONFIGS, synthetic_state_dict
    from ref_ops import RefOps
    torch.set_num_threads(2)
    cfg = NAMED_CONFIGS["tiny_depth"]
    eng = Engine(cfg, RefOps())
    eng.load_state_dict(synthetic_state_dict(cfg, 0))
    full = synth.make_inputs(cfg, 4, seed=9)                    # maps, masks and the downsampler input are part of the batch
    mine = shard_batch(full, rank, world)
    lo, hi = shard_range(4, rank, world)
    assert mine["grounding_input"]["depth"].shape[0] == 2 and mine["grounding_extra_input"].shape[0] == 2
    ts = torch.tensor([981, 501, 21, 1])
    e_c, e_u = eng.forward_cfg(mine["x"], ts[lo:hi], mine["context"], mine["uc"], mine["grounding_input"], None, mine["grounding_extra_input"])
    allc, allu = gather_latents(e_c.clone()), gather_latents(e_u.clone())
    if rank == 0:
        c, u = eng.forward_cfg(full["x"], ts, full["context"], full["uc"], full["grounding_input"], None, full["grounding_extra_input"])
        torch.save({"gathered": torch.cat([allc, allu]), "single": torch.cat([c, u]).clone()}, os.path.join(tmp, "out_spatial.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_equivalence_spatial_gloo(tmp_path):
    """A spatial-map model (ConvNeXt tokenizer + grounding downsampler in the plan's static part) shards the same way: the maps and the
    downsampler input are rows of the batch."""
    port = 31500 + (os.getpid() % 2000)
    mp.start_processes(_worker_spatial, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    out = torch.load(os.path.join(tmp_path, "out_spatial.pt"))
    assert out["gathered"].shape == out["single"].shape == (8, 4, 16, 16)
    assert (out["gathered"] - out["single"]).abs().max() < 5e-5


def test_shard_range_remainders():
    from gligen_b200.dist import shard_range
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard_range(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
