"""GPU: the rows chained the way gligen_inference.run() chains them (:377-441) - text encoder -> PLMS sampling loop with CFG
and scheduled sampling -> VAE decode - every stage on this repo's kernels through the drop-in classes, against the same chain of
CPU oracles (each pinned separately against the reference / the library) on the same seeded weights, token ids and noise.
Small models (tiny CLIP tower, tiny UNet, tiny VAE), B = 2, PLMS S = 4.  Tolerances: the short-loop latent tolerance of DESIGN 2
(rel-L2 <= 6e-2 - the text context now carries the encoder's bf16 error too) and rel-L2 <= 8e-2 on the decoded image."""
import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_text_to_image_chain():
    from functools import partial
    from gligen_b200 import synth
    from gligen_b200.clip_text import TINY_CLIP_TEXT, synthetic_clip_state_dict, synthetic_token_ids
    from gligen_b200.pipeline import alpha_generator, build_model, sampler_inputs, set_alpha_scale, to_device
    from gligen_b200.spec import NAMED_VAE_CONFIGS, synthetic_state_dict, synthetic_vae_encoder_state_dict, synthetic_vae_state_dict
    from ldm.models.autoencoder import AutoencoderKL
    from ldm.models.diffusion.ldm import LatentDiffusion
    from ldm.models.diffusion.plms import PLMSSampler
    from ldm.util import instantiate_from_config
    from oracle import sampler_oracle as SO, unet_oracle as UO
    from oracle.clip_oracle import clip_text_forward
    from oracle.vae_oracle import vae_decode
    B, S, G, guidance, alpha_type = 2, 4, 6, 5.0, [1.0, 0.0, 0.0]
    # ---- text encoder: prompt and negative prompt -----------------------------------------------------------------------------
    text = instantiate_from_config(dict(target="ldm.modules.encoders.modules.FrozenCLIPEmbedder", params=dict(text_config="tiny_clip_text"))).to(DEV).eval()
    csd = synthetic_clip_state_dict(TINY_CLIP_TEXT, 0)
    text.load_state_dict(csd)
    ids, neg = synthetic_token_ids(TINY_CLIP_TEXT, B, 3), synthetic_token_ids(TINY_CLIP_TEXT, B, 4)
    context, uc = text.encode_tokens(ids), text.encode_tokens(neg)
    ctx_ref, uc_ref = clip_text_forward(TINY_CLIP_TEXT, csd, ids)[0], clip_text_forward(TINY_CLIP_TEXT, csd, neg)[0]
    assert_close(context, ctx_ref, rel=1.5e-2, max_rel=6e-2, what="context")
    # ---- sampling loop ------------------------------------------------------------------------------------------------------------
    cfg, model = build_model("tiny", device=DEV)
    sd = synthetic_state_dict(cfg, 0)
    inp = synth.make_inputs(cfg, B, G, seed=7)
    dinp = dict(x=inp["x"].to(DEV), context=context, uc=uc)
    input, _, _ = sampler_inputs(cfg, model, dinp, to_device(inp["batch"], DEV))
    diffusion = LatentDiffusion(linear_start=0.00085, linear_end=0.012, timesteps=1000).to(DEV)
    sampler = PLMSSampler(diffusion, model, alpha_generator_func=partial(alpha_generator, type=alpha_type), set_alpha_scale=set_alpha_scale)
    torch.manual_seed(1234)
    shape = (B, cfg.in_channels, cfg.image_size, cfg.image_size)
    lat = sampler.sample(S=S, shape=shape, input=input, uc=uc, guidance_scale=guidance)
    null = UO.null_grounding(cfg, inp["grounding_input"])

    def eps_fn(x, t, cond):
        return UO.unet_forward(cfg, sd, x, t, ctx_ref if cond else uc_ref, inp["grounding_input"] if cond else null, 1.0)

    torch.manual_seed(1234)
    lat_ref = SO.plms_sample(eps_fn, S, shape, x_T=inp["x"].clone(), use_cfg=True, guidance_scale=guidance)
    r = assert_close(lat, lat_ref, rel=6e-2, max_rel=0.12, what="PLMS latent")
    # ---- VAE decode -------------------------------------------------------------------------------------------------------------------
    v = NAMED_VAE_CONFIGS["tiny_vae64"]
    dd = dict(double_z=True, z_channels=v.z_channels, resolution=v.image_size, in_channels=3, out_ch=v.out_ch, ch=v.ch, ch_mult=list(v.ch_mult),
              num_res_blocks=v.num_res_blocks, attn_resolutions=[], dropout=0.0)
    vae = AutoencoderKL(ddconfig=dd, embed_dim=v.embed_dim, scale_factor=v.scale_factor)
    vsd = dict(synthetic_vae_state_dict(v, 0)); vsd.update(synthetic_vae_encoder_state_dict(v, 1))
    vae.load_state_dict(vsd, strict=False)
    vae = vae.to(DEV).eval()
    img = vae.decode(lat)
    img_ref = vae_decode(v, vsd, lat_ref)
    torch.cuda.synchronize()
    ri = assert_close(img, img_ref, rel=8e-2, max_rel=0.2, what="decoded image")
    print(f"\\nchain: context rel-L2 ok; PLMS S={S} latent rel-L2 {r[0]:.3e} max-rel {r[1]:.3e}; image rel-L2 {ri[0]:.3e} max-rel {ri[1]:.3e}")
