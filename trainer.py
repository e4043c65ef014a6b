"""Only the two helpers gligen_inference.py imports from the reference's trainer.py (:64-92); training
itself is outside this repo's scope."""
import torch


def read_official_ckpt(ckpt_path):
    """Split an official SD checkpoint by key prefix into model / text_encoder / autoencoder / unexpected."""
    state_dict = torch.load(ckpt_path, map_location="cpu")["state_dict"]
    out = {"model": {}, "text_encoder": {}, "autoencoder": {}, "unexpected": {}, "diffusion": {}}
    prefixes = (("model.diffusion_model.", "model"), ("cond_stage_model.", "text_encoder"), ("first_stage_model.", "autoencoder"))
    for k, v in state_dict.items():
        for pre, name in prefixes:
            if k.startswith(pre):
                out[name][k[len(pre):]] = v
                break
        else:
            out["unexpected"][k] = v
    out["diffusion"] = None
    return out


def batch_to_device(batch, device):
    for k in batch:
        if isinstance(batch[k], torch.Tensor):
            batch[k] = batch[k].to(device)
    return batch
