"""CPU: the HOST logic of the drop-in samplers and UNetModel surface (RNG consumption order, DDIM/PLMS schedules,
alpha scheduling + first-conv swap, inpainting blend, cond+uncond batching, sub-stepping of PLMS) against the latents the
REAL reference samplers produced (tests/golden).  The CUDA kernels are replaced by their torch-fp32 statements
(tests/ref_ops.py) through test-only injection - the product classes themselves refuse to run without CUDA."""
import os
from functools import partial

import pytest
import torch

import test_engine_gpu as teg          # shared config glue (build_model, alpha_generator, set_alpha_scale, make_case)
from conftest import GOLD, rel_l2
from gligen_b200.engine import Engine
from ref_ops import RefOps


@pytest.fixture
def cpu_backend(monkeypatch):
    from ldm.models.diffusion import _sampling
    from ldm.modules.diffusionmodules.openaimodel import UNetModel

    def engine(self):
        if self._engine is None:
            self._engine = Engine(self.cfg, RefOps())
            self._engine_stale = True
        if self._engine_stale:
            self._engine.load_state_dict(self.state_dict())
            self._engine_stale = False
        return self._engine

    def update(self, x, e_c, e_u, guidance_scale, olds, coefs, index, want_e):
        torch.randn_like(x)                      # same generator draw as the product path (sigma_t == 0 noise)
        x_prev = torch.empty_like(x)
        e_out = torch.empty_like(x) if want_e else None
        RefOps().sampler_update(x, e_c, e_u, float(guidance_scale), olds, [float(c) for c in coefs],
                                float(self.ddim_alphas[index]), float(self.ddim_alphas_prev[index]), e_out, x_prev)
        return x_prev, e_out

    monkeypatch.setattr(UNetModel, "engine", engine)
    monkeypatch.setattr(_sampling.SamplerBase, "_update", update)
    monkeypatch.setattr(teg, "DEV", "cpu")


@pytest.mark.parametrize("name,gold_file", [("tiny", "tiny_B2_G6.pt"), ("tiny_text_image", "tiny_text_image_B2_G5.pt"),
                                            ("tiny_keypoint", "tiny_keypoint_B2_G34.pt"), ("tiny_inpaint", "tiny_inpaint_B2_G6.pt")])
def test_dropin_samplers_match_reference_latents(cpu_backend, name, gold_file):
    from ldm.models.diffusion.ddim import DDIMSampler
    from ldm.models.diffusion.ldm import LatentDiffusion
    from ldm.models.diffusion.plms import PLMSSampler
    gold = torch.load(os.path.join(GOLD, gold_file))
    diffusion = LatentDiffusion(linear_start=0.00085, linear_end=0.012, timesteps=1000)
    cwd = os.getcwd()
    os.chdir(GOLD)                               # restore_first_conv_from_SD reads a CWD-relative file, like the reference
    try:
        for kind in ("plms", "ddim"):
            g = gold[kind]
            cfg, model = teg.build_model(name)
            inp, extra, mask, z0 = teg.make_case(cfg, gold)
            grounding = model.grounding_tokenizer_input.prepare(inp["batch"])
            cls = PLMSSampler if kind == "plms" else DDIMSampler
            sampler = cls(diffusion, model, alpha_generator_func=partial(teg.alpha_generator, type=g["alpha_type"]),
                          set_alpha_scale=teg.set_alpha_scale)
            input = dict(x=inp["x"].clone(), timesteps=None, context=inp["context"], grounding_input=grounding,
                         inpainting_extra_input=extra, grounding_extra_input=None)
            shape = (gold["B"], cfg.in_channels, cfg.image_size, cfg.image_size)
            torch.manual_seed(1234)
            lat = sampler.sample(S=g["S"], shape=shape, input=input, uc=inp["uc"], guidance_scale=g["guidance"], mask=mask, x0=z0)
            r = rel_l2(lat, g["latent"])
            assert r < 2e-4, f"{name} {kind}: latent rel_l2 {r:.3e}"
    finally:
        os.chdir(cwd)
