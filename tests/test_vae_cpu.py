"""VAE decoder engine wiring on the CPU: gligen_b200.vae.VAEDecoderEngine executed with the torch-fp32 checker ops
(tests/ref_ops.py) against the REFERENCE decode goldens (tests/golden/*_vae*.pt, oracle/gen_golden.py --vae).  Validates the
restructurings that are exact in real arithmetic - post_quant_conv + 1/scale_factor folded into conv_in through a ones
channel, V^T produced directly, v's bias added after P.V, channels-last layouts, packed 3x3 weights - before any kernel runs."""
import os

import pytest
import torch

from conftest import GOLD, assert_close
from gligen_b200.spec import NAMED_VAE_CONFIGS, synthetic_vae_encoder_state_dict, synthetic_vae_state_dict
from gligen_b200.vae import VAEDecoderEngine, VAEEncoderEngine
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


@pytest.mark.parametrize("name,B", [("tiny_vae64", 2), ("small_vae", 1)])
def test_vae_encoder_wiring_matches_reference(name, B):
    """Encoder half: conv_out o quant_conv folded into one 3x3, asymmetric-pad stride-2 gather, against the
    REFERENCE moments (oracle/gen_golden.py --vae-enc)."""
    cfg = NAMED_VAE_CONFIGS[name]
    gold = torch.load(os.path.join(GOLD, f"{name}_enc_B{B}.pt"))
    eng = VAEEncoderEngine(cfg, RefOps("cpu", torch.float32))
    eng.load_state_dict(synthetic_vae_encoder_state_dict(cfg, 1))
    mom = eng.encode_moments(gold["x"].float())
    assert_close(mom, gold["moments"], rel=2e-5, max_rel=1e-4, what=f"{name} encode moments (checker ops)")


def test_vae_oracle_encode_matches_reference_fixture():
    """oracle/vae_oracle.py encode against the committed reference outputs, including the posterior sample drawn from
    torch's global CPU generator (distributions.py:36)."""
    from oracle import vae_oracle as VO
    for name, B in (("tiny_vae", 2), ("tiny_vae64", 2)):
        cfg = NAMED_VAE_CONFIGS[name]
        gold = torch.load(os.path.join(GOLD, f"{name}_enc_B{B}.pt"))
        sd = synthetic_vae_encoder_state_dict(cfg, 1)
        assert torch.equal(VO.vae_encode_moments(cfg, sd, gold["x"].float()), gold["moments"])
        torch.manual_seed(gold["noise_seed"])
        assert torch.equal(VO.vae_encode(cfg, sd, gold["x"].float()), gold["z0"])


def test_dropin_autoencoder_surface():
    """stand-alone drop-in (no reference behind this repo in the test process): reference parameter names, strict load."""
    from ldm.models.autoencoder import AutoencoderKL
    from gligen_b200.spec import vae_decoder_param_shapes, vae_encoder_param_shapes
    cfg = NAMED_VAE_CONFIGS["tiny_vae64"]
    dd = dict(double_z=True, z_channels=cfg.z_channels, resolution=16, in_channels=3, out_ch=cfg.out_ch, ch=cfg.ch, ch_mult=list(cfg.ch_mult),
              num_res_blocks=cfg.num_res_blocks, attn_resolutions=[], dropout=0.0)
    m = AutoencoderKL(ddconfig=dd, embed_dim=cfg.embed_dim, scale_factor=cfg.scale_factor).eval()
    want = vae_decoder_param_shapes(cfg)
    have = {k: tuple(v.shape) for k, v in m.state_dict().items() if k.startswith(("decoder.", "post_quant_conv."))}
    assert have == {k: tuple(s) for k, s in want.items()}
    want_e = vae_encoder_param_shapes(cfg)
    have_e = {k: tuple(v.shape) for k, v in m.state_dict().items() if k.startswith(("encoder.", "quant_conv."))}
    assert have_e == {k: tuple(s) for k, s in want_e.items()}
    sd = dict(synthetic_vae_state_dict(cfg, 0))
    sd.update(synthetic_vae_encoder_state_dict(cfg, 1))                    # a full checkpoint: both halves, strict
    m.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.decode(torch.zeros(1, 4, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.encode(torch.zeros(1, 3, 16, 16))
