"""VAE decoder engine wiring on the CPU: gligen_b200.vae.VAEDecoderEngine executed with the torch-fp32 checker ops
(tests/ref_ops.py) against the REFERENCE decode goldens (tests/golden/*_vae*.pt, oracle/gen_golden.py --vae).  Validates the
restructurings that are exact in real arithmetic - post_quant_conv + 1/scale_factor folded into conv_in through a ones
channel, V^T produced directly, v's bias added after P.V, channels-last layouts, packed 3x3 weights - before any kernel runs."""
import os

import pytest
import torch

from conftest import GOLD, assert_close
from gligen_b200.spec import NAMED_VAE_CONFIGS, synthetic_vae_state_dict
from gligen_b200.vae import VAEDecoderEngine
from ref_ops import RefOps


@pytest.mark.parametrize("name,B", [("tiny_vae64", 2), ("small_vae", 1)])
def test_vae_engine_wiring_matches_reference(name, B):
    cfg = NAMED_VAE_CONFIGS[name]
    gold = torch.load(os.path.join(GOLD, f"{name}_B{B}.pt"))
    eng = VAEDecoderEngine(cfg, RefOps("cpu", torch.float32))
    eng.load_state_dict(synthetic_vae_state_dict(cfg, 0))
    img = eng.decode(gold["z"])
    tol = 2e-3 if gold["image"].dtype == torch.float16 else 2e-5          # fp16-stored fixture vs fp32
    assert_close(img, gold["image"].float(), rel=tol, max_rel=tol * 5, what=f"{name} decode (checker ops)")


def test_dropin_autoencoder_surface():
    """decoder-only drop-in (no reference behind this repo in the test process): reference parameter names, strict=False load."""
    from ldm.models.autoencoder import AutoencoderKL
    from gligen_b200.spec import vae_decoder_param_shapes
    cfg = NAMED_VAE_CONFIGS["tiny_vae64"]
    dd = dict(double_z=True, z_channels=cfg.z_channels, resolution=16, in_channels=3, out_ch=cfg.out_ch, ch=cfg.ch, ch_mult=list(cfg.ch_mult),
              num_res_blocks=cfg.num_res_blocks, attn_resolutions=[], dropout=0.0)
    m = AutoencoderKL(ddconfig=dd, embed_dim=cfg.embed_dim, scale_factor=cfg.scale_factor).eval()
    want = vae_decoder_param_shapes(cfg)
    have = {k: tuple(v.shape) for k, v in m.state_dict().items() if k.startswith(("decoder.", "post_quant_conv."))}
    assert have == {k: tuple(s) for k, s in want.items()}
    sd = synthetic_vae_state_dict(cfg, 0)
    sd["encoder.conv_in.weight"] = torch.zeros(1)                         # a full checkpoint also carries the encoder half
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing if k.startswith(("decoder.", "post_quant_conv."))]
    with pytest.raises(RuntimeError, match="CUDA"):
        m.decode(torch.zeros(1, 4, 8, 8))
