"""CPU ORACLE for the sampling loops around the UNet (PLMS / DDIM).  TEST INFRASTRUCTURE ONLY.

Functional restatement of ldm/models/diffusion/plms.py:25-162 and ddim.py:27-134, the schedule
helpers (ldm/modules/diffusionmodules/util.py:30-83, ldm/models/diffusion/ddpm.py:19-54) and the
scheduled-sampling helpers of gligen_inference.py:24-66.  RNG consumption order is preserved
(randn(shape) once, randn_like per x_prev even though sigma = 0, randn_like per q_sample).

Pinned by oracle/gen_golden.py against the reference classes executed in-process.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import numpy as np
import torch


# ---- schedule ---------------------------------------------------------------------------------
def make_schedule(linear_start=0.00085, linear_end=0.012, timesteps=1000) -> Dict[str, torch.Tensor]:
    """util.py:30-35 ("linear" = linspace of sqrt(beta), squared; float64) + ddpm.py:19-43 (-> fp32)."""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
    ac = np.cumprod(1.0 - betas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return dict(betas=f32(betas), alphas_cumprod=f32(ac), alphas_cumprod_prev=f32(ac_prev),
                sqrt_alphas_cumprod=f32(np.sqrt(ac)), sqrt_one_minus_alphas_cumprod=f32(np.sqrt(1.0 - ac)))


def ddim_timesteps(S: int, T: int = 1000) -> np.ndarray:
    """util.py:55-69, 'uniform': range(0, T, T//S) + 1."""
    c = T // S
    return np.asarray(list(range(0, T, c))) + 1


def ddim_parameters(alphas_cumprod: torch.Tensor, steps: np.ndarray, eta: float = 0.0):
    """util.py:72-83 (called with alphacums on the CPU, plms.py:46)."""
    alphas = alphas_cumprod[steps]
    alphas_prev = np.asarray([alphas_cumprod[0]] + alphas_cumprod[steps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def alpha_generator(length: int, type: Optional[List[float]] = None) -> List[float]:
    """gligen_inference.py:31-66."""
    if type is None:
        type = [1, 0, 0]
    assert len(type) == 3 and type[0] + type[1] + type[2] == 1
    s0 = int(type[0] * length)
    s1 = int(type[1] * length)
    s2 = length - s0 - s1
    decay = list(np.arange(start=0, stop=1, step=1 / s1)[::-1]) if s1 != 0 else []
    alphas = [1] * s0 + decay + [0] * s2
    assert len(alphas) == length
    return alphas


def q_sample(sched, x_start, t, noise=None):
    """ldm.py:19-22."""
    if noise is None:
        noise = torch.randn_like(x_start)
    a = sched["sqrt_alphas_cumprod"].gather(-1, t).reshape(-1, 1, 1, 1)
    b = sched["sqrt_one_minus_alphas_cumprod"].gather(-1, t).reshape(-1, 1, 1, 1)
    return a * x_start + b * noise


# ---- samplers -------------------------------------------------------------------------------------
# eps_fn(x, t, cond: bool) -> eps.   on_alpha(alpha) is called once per step BEFORE the model runs
# (set_alpha_scale + restore_first_conv_from_SD when alpha == 0, plms.py:85-89).
EpsFn = Callable[[torch.Tensor, torch.Tensor, bool], torch.Tensor]


def _cfg_eps(eps_fn: EpsFn, x, t, use_cfg: bool, g: float):
    e = eps_fn(x, t, True)
    if use_cfg:
        eu = eps_fn(x, t, False)
        e = eu + g * (e - eu)                                   # plms.py:121
    return e


@torch.no_grad()
def plms_sample(eps_fn: EpsFn, S: int, shape, sched=None, x_T=None, use_cfg=True, guidance_scale=7.5,
                alphas: Optional[List[float]] = None, on_alpha=None, mask=None, x0=None):
    """plms.py:59-162."""
    sched = sched or make_schedule()
    steps = ddim_timesteps(S)
    sig, al, alp = ddim_parameters(sched["alphas_cumprod"], steps)
    sq1m = np.sqrt(1.0 - al)
    b = shape[0]
    img = torch.randn(shape) if x_T is None else x_T
    time_range = np.flip(steps)
    total = steps.shape[0]
    old_eps: List[torch.Tensor] = []
    use_cfg = use_cfg and guidance_scale != 1

    def x_prev_of(x, e, index):                                   # plms.py:125-139
        a_t = torch.full((b, 1, 1, 1), al[index])
        a_prev = torch.full((b, 1, 1, 1), alp[index])
        s_t = torch.full((b, 1, 1, 1), sig[index])
        s1 = torch.full((b, 1, 1, 1), sq1m[index])
        pred_x0 = (x - s1 * e) / a_t.sqrt()
        dir_xt = (1.0 - a_prev - s_t ** 2).sqrt() * e
        noise = s_t * torch.randn_like(x)
        return a_prev.sqrt() * pred_x0 + dir_xt + noise

    for i, step in enumerate(time_range):
        if alphas is not None and on_alpha is not None:
            on_alpha(alphas[i])
        index = total - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        ts_next = torch.full((b,), int(time_range[min(i + 1, len(time_range) - 1)]), dtype=torch.long)
        if mask is not None:
            img = q_sample(sched, x0, ts) * mask + (1.0 - mask) * img      # plms.py:96-100
        x = img
        e_t = _cfg_eps(eps_fn, x, ts, use_cfg, guidance_scale)
        if len(old_eps) == 0:
            x_p = x_prev_of(x, e_t, index)
            e_next = _cfg_eps(eps_fn, x_p, ts_next, use_cfg, guidance_scale)
            e_prime = (e_t + e_next) / 2
        elif len(old_eps) == 1:
            e_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        img = x_prev_of(x, e_prime, index)
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
    return img


@torch.no_grad()
def ddim_sample(eps_fn: EpsFn, S: int, shape, sched=None, x_T=None, use_cfg=True, guidance_scale=7.5,
                alphas: Optional[List[float]] = None, on_alpha=None, mask=None, x0=None):
    """ddim.py:59-134 (eta = 0)."""
    sched = sched or make_schedule()
    steps = ddim_timesteps(S)
    sig, al, alp = ddim_parameters(sched["alphas_cumprod"], steps)
    sq1m = np.sqrt(1.0 - al)
    b = shape[0]
    img = torch.randn(shape) if x_T is None else x_T
    time_range = np.flip(steps)
    total = steps.shape[0]
    use_cfg = use_cfg and guidance_scale != 1
    for i, step in enumerate(time_range):
        if alphas is not None and on_alpha is not None:
            on_alpha(alphas[i])
        index = total - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        if mask is not None:
            img = q_sample(sched, x0, ts) * mask + (1.0 - mask) * img
        e = _cfg_eps(eps_fn, img, ts, use_cfg, guidance_scale)
        a_t = torch.full((b, 1, 1, 1), al[index])
        a_prev = torch.full((b, 1, 1, 1), alp[index])
        s_t = torch.full((b, 1, 1, 1), sig[index])
        s1 = torch.full((b, 1, 1, 1), sq1m[index])
        pred_x0 = (img - s1 * e) / a_t.sqrt()
        dir_xt = (1.0 - a_prev - s_t ** 2).sqrt() * e
        noise = s_t * torch.randn_like(img)
        img = a_prev.sqrt() * pred_x0 + dir_xt + noise
    return img


def draw_masks_from_boxes(boxes: torch.Tensor, size: int) -> torch.Tensor:
    """inpaint_mask_func.py:16-41 with randomize_fg_mask=False, random_add_bg_mask=False:
    mask [B,1,size,size] = 1 everywhere, 0 inside each box (int-truncated pixel coords)."""
    image_masks = []
    for box in boxes:
        m = torch.ones(size, size)
        for bx in box:
            x0, x1 = bx[0] * size, bx[2] * size
            y0, y1 = bx[1] * size, bx[3] * size
            m[int(y0):int(y1), int(x0):int(x1)] = 0
        image_masks.append(m)
    return torch.stack(image_masks).unsqueeze(1)
