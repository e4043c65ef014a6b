"""Pin oracle/clip_oracle.py against the installed `transformers` CLIPTextModel (the library FrozenCLIPEmbedder wraps) and write
tests/golden/clip_text_*.pt.          python oracle/gen_golden_clip.py

The reference pins transformers==4.19.2 (env_docker/Dockerfile:3); this container has another release, whose CLIPTextModel computes
the same function (eos_token_id = 2, the shipped openai/clip-vit-large-patch14 config value, selects the argmax(input_ids) pooling
of 4.19.2).  Weights: seeded synthetic (no checkpoint can be downloaded here)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from gligen_b200.clip_text import NAMED_CLIP_CONFIGS, synthetic_clip_state_dict, synthetic_token_ids  # noqa: E402
from oracle.clip_oracle import clip_text_forward  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def run(name, B, seed):
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel
    cfg = NAMED_CLIP_CONFIGS[name]
    hf = CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.width, intermediate_size=cfg.ffn, num_hidden_layers=cfg.layers,
                        num_attention_heads=cfg.heads, max_position_embeddings=cfg.max_length, hidden_act="quick_gelu", layer_norm_eps=cfg.eps,
                        eos_token_id=2, attn_implementation="eager")
    model = CLIPTextModel(hf).eval()
    sd = synthetic_clip_state_dict(cfg, 0, prefix="")
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("position_ids") for k in missing), (missing, unexpected)
    ids = synthetic_token_ids(cfg, B, seed)
    with torch.no_grad():
        out = model(input_ids=ids)
    z, pooled = clip_text_forward(cfg, sd, ids, prefix="text_model.")
    e1, e2 = (z - out.last_hidden_state).abs().max().item(), (pooled - out.pooler_output).abs().max().item()
    print(f"{name}: B={B} oracle vs transformers {transformers.__version__} max-abs last_hidden_state {e1:.2e} pooler_output {e2:.2e}; |z| max {z.abs().max():.2f}")
    assert e1 <= 2e-4 and e2 <= 2e-4
    torch.save({"config": name, "B": B, "seed": seed, "transformers": transformers.__version__, "input_ids": ids,
                "last_hidden_state": out.last_hidden_state.clone(), "pooler_output": out.pooler_output.clone(),
                "oracle_vs_library_max_abs": (e1, e2)}, os.path.join(GOLD, f"clip_text_{name}.pt"))


if __name__ == "__main__":
    run("tiny_clip_text", 3, 5)
    run("sd14_clip_text", 2, 6)
