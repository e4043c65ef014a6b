"""LatentDiffusion: the schedule + q_sample the samplers need (reference ldm/models/diffusion/ldm.py:11-22)."""
import torch

from .ddpm import DDPM


class LatentDiffusion(DDPM):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.clip_denoised = False

    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = torch.randn_like(x_start)
        shape = (t.shape[0],) + (1,) * (x_start.dim() - 1)
        a = self.sqrt_alphas_cumprod.gather(-1, t).reshape(shape)
        s = self.sqrt_one_minus_alphas_cumprod.gather(-1, t).reshape(shape)
        return a * x_start + s * noise


# names this drop-in does not define resolve to the reference module of the same name when a reference checkout
# follows this repo on sys.path (gligen_b200/_overlay.py)
from gligen_b200._overlay import fallback as _fallback  # noqa: E402

__getattr__ = _fallback(__name__, __file__)
