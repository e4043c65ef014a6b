"""Only the two helpers gligen_inference.py imports from the reference's trainer.py (:64-92); training
itself is outside this repo's scope."""
import torch


def read_official_ckpt(ckpt_path):
    """Split an official SD checkpoint by key prefix (reference trainer.py:64-85): UNet keys -> "model", CLIP ->
    "text_encoder", VAE -> "autoencoder", the two EMA counters -> "unexpected", everything else (schedule buffers) ->
    "diffusion"."""
    state_dict = torch.load(ckpt_path, map_location="cpu")["state_dict"]
    out = {"model": {}, "text_encoder": {}, "autoencoder": {}, "unexpected": {}, "diffusion": {}}
    prefixes = (("model.diffusion_model", "model"), ("cond_stage_model", "text_encoder"), ("first_stage_model", "autoencoder"))
    for k, v in state_dict.items():
        for pre, name in prefixes:
            if k.startswith(pre):
                out[name][k.replace(pre + ".", "")] = v
                break
        else:
            out["unexpected" if k in ("model_ema.decay", "model_ema.num_updates") else "diffusion"][k] = v
    return out


def batch_to_device(batch, device):
    for k in batch:
        if isinstance(batch[k], torch.Tensor):
            batch[k] = batch[k].to(device)
    return batch
