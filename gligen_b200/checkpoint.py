"""Real-checkpoint path: what `gligen_inference.load_ckpt` (gligen_inference.py:70-86) and
`convert_ckpt.py:5-16` do, for a GLIGEN checkpoint file in the reference's format

    {"model": unet state dict (966 keys), "autoencoder": ..., "text_encoder": ..., "diffusion": ...,
     "config_dict": {"_content": {"model": {target, params}, "autoencoder": {...}, "text_encoder": {...}, "diffusion": {...},
                                  "grounding_tokenizer_input": {...}, ...}}}

Every `target` is a dotted class path resolved by `ldm.util.instantiate_from_config`; with this repo on the path the UNet,
diffusion wrapper and VAE resolve to the drop-in classes, the text encoder to the reference's (when a checkout is behind
this repo; it needs the `clip` / `transformers` weights the reference needs).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch


def add_additional_channels(state_dict: Dict[str, torch.Tensor], num_additional_channels: int) -> None:
    """convert_ckpt.py:5-16: widen the first conv of a UNet state dict IN PLACE by `num_additional_channels` zero input
    channels (inpainting: 4 latent + 4 masked-latent + 1 mask = +5; openaimodel.py:299-305).  Sizes come from the tensor
    itself (the reference hard-codes 320 x 4)."""
    if num_additional_channels == 0:
        return
    key = "input_blocks.0.0.weight"
    w = state_dict[key]
    wide = torch.zeros(w.shape[0], w.shape[1] + num_additional_channels, w.shape[2], w.shape[3], dtype=w.dtype, device=w.device)
    wide[:, : w.shape[1]] = w
    state_dict[key] = wide


def save_ckpt(path: str, config: Dict, model_sd: Dict, autoencoder_sd: Optional[Dict] = None, text_encoder_sd: Optional[Dict] = None,
              diffusion_sd: Optional[Dict] = None) -> None:
    """Write a checkpoint in the layout `load_ckpt` reads (Trainer.save_ckpt_and_result, trainer.py:441-463, stores the
    OmegaConf config as `config_dict`; its plain-dict form sits under `_content`)."""
    out = {"model": model_sd, "config_dict": {"_content": config}}
    if autoencoder_sd is not None:
        out["autoencoder"] = autoencoder_sd
    if text_encoder_sd is not None:
        out["text_encoder"] = text_encoder_sd
    if diffusion_sd is not None:
        out["diffusion"] = diffusion_sd
    torch.save(out, path)


def load_ckpt(ckpt_path: str, device="cuda", with_text_encoder: bool = True) -> Tuple[object, object, object, object, Dict]:
    """(model, autoencoder, text_encoder, diffusion, config) exactly like gligen_inference.load_ckpt.  `with_text_encoder=False`
    skips the CLIP text encoder (returns None) for hosts without the `clip` package."""
    from ldm.util import instantiate_from_config
    saved = torch.load(ckpt_path, map_location="cpu")
    config = saved["config_dict"]["_content"]
    model = instantiate_from_config(config["model"]).to(device).eval()
    model.load_state_dict(saved["model"])
    if "grounding_tokenizer_input" in config:
        model.grounding_tokenizer_input = instantiate_from_config(config["grounding_tokenizer_input"])      # gligen_inference.py:348-349
    autoencoder = None
    if "autoencoder" in config:
        autoencoder = instantiate_from_config(config["autoencoder"]).to(device).eval()
        own = set(autoencoder.state_dict())
        sd = saved["autoencoder"]
        # the decoder-only drop-in (no reference checkout behind this repo) holds just decoder.* / post_quant_conv.*
        autoencoder.load_state_dict(sd if own >= set(sd) else {k: v for k, v in sd.items() if k in own}, strict=own >= set(sd))
    text_encoder = None
    if with_text_encoder and "text_encoder" in config:
        text_encoder = instantiate_from_config(config["text_encoder"]).to(device).eval()
        text_encoder.load_state_dict(saved["text_encoder"])
    diffusion = instantiate_from_config(config["diffusion"]).to(device)
    if "diffusion" in saved:
        diffusion.load_state_dict(saved["diffusion"])
    return model, autoencoder, text_encoder, diffusion, config
