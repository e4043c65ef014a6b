"""CPU: the CLIP text encoder row.  Oracle against the fixture written from the installed transformers' CLIPTextModel
(oracle/gen_golden_clip.py, bit-identical at generation time); the engine's wiring (fused QKV packing, causal attention call,
residual / quick_gelu epilogues, pooling) executed with the torch-fp32 checker ops; the drop-in FrozenCLIPEmbedder surface."""
import os

import pytest
import torch

from conftest import GOLD
from gligen_b200.clip_text import NAMED_CLIP_CONFIGS, ClipTextEngine, clip_text_param_shapes, synthetic_clip_state_dict
from oracle.clip_oracle import clip_text_forward
from ref_ops import RefOps


@pytest.mark.parametrize("name", ["tiny_clip_text", "sd14_clip_text"])
def test_oracle_matches_library_fixture(name):
    g = torch.load(os.path.join(GOLD, f"clip_text_{name}.pt"))
    cfg = NAMED_CLIP_CONFIGS[name]
    z, pooled = clip_text_forward(cfg, synthetic_clip_state_dict(cfg, 0), g["input_ids"])
    assert (z - g["last_hidden_state"]).abs().max() <= 2e-5 and (pooled - g["pooler_output"]).abs().max() <= 2e-5


@pytest.mark.parametrize("name", ["tiny_clip_text", "sd14_clip_text"])
def test_engine_wiring(name):
    g = torch.load(os.path.join(GOLD, f"clip_text_{name}.pt"))
    cfg = NAMED_CLIP_CONFIGS[name]
    eng = ClipTextEngine(cfg, RefOps())
    eng.load_state_dict(synthetic_clip_state_dict(cfg, 0))
    z, pooled = eng.forward(g["input_ids"])
    assert (z - g["last_hidden_state"]).abs().max() <= 1e-4 and (pooled - g["pooler_output"]).abs().max() <= 1e-4
    z2, _ = eng.forward(g["input_ids"][:1, :40])          # shorter sequences / other batch sizes get their own workspace
    assert (z2 - clip_text_forward(cfg, synthetic_clip_state_dict(cfg, 0), g["input_ids"][:1, :40])[0]).abs().max() <= 1e-4


def test_drop_in_surface():
    from ldm.util import instantiate_from_config
    m = instantiate_from_config(dict(target="ldm.modules.encoders.modules.FrozenCLIPEmbedder")).eval()
    assert set(m.state_dict()) == set(clip_text_param_shapes(m.cfg, "transformer."))
    sd = synthetic_clip_state_dict(m.cfg, 0)
    sd["transformer.text_model.embeddings.position_ids"] = torch.arange(77)[None]     # what transformers 4.19.2 checkpoints carry
    m.load_state_dict(sd)                                                              # strict
    k = "transformer.text_model.encoder.layers.3.mlp.fc1.weight"
    assert torch.equal(m.state_dict()[k], sd[k])
    with pytest.raises(RuntimeError):
        m.encode_tokens(torch.zeros(1, 77, dtype=torch.int64))                         # CUDA only
