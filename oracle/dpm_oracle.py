"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never import this from the product path.

CPU restatement of what the pipeline does between two transformer calls:
classifier-free-guidance combine with the dynamic scale (PIPE:991-997), `CogVideoXDPMScheduler.step` (DPM:330-439 with
get_variables DPM:306-317 and get_mult DPM:319-328) and the `replace_gt` re-noise / blend (PIPE:1017-1034, add_noise
DPM:442-466), including the reference's dtype behaviour: the scheduler's alpha table is float64 (`scaled_linear` betas,
DPM:203), the per-step coefficients are 0-dim float64 tensors, products of such a coefficient with the bf16 latents stay
bf16 — the coefficient itself is first rounded to bf16, then the product once more — everything touching the fp32 model
output is fp32.

PIPE = pipelines/cogvideo/pipeline_cogvideox_inpainting_i2v_branch_anyl.py, DPM = schedulers/scheduling_dpm_cogvideox.py
(under /root/reference/diffusers/src/diffusers).  Pinned by tests/golden/dpm_steps.pt, which oracle/make_golden_dpm.py
produces by running the real scheduler.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


def alphas_cumprod_table(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.0120, snr_shift_scale=1.0,
                         rescale_betas_zero_snr=True) -> torch.Tensor:
    """DPM:199-221 for beta_schedule="scaled_linear" (float64)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
    ac = torch.cumprod(1.0 - betas, dim=0)
    ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)
    if rescale_betas_zero_snr:                       # rescale_zero_terminal_snr, DPM:85-112
        s = ac.sqrt()
        s0, sT = s[0].clone(), s[-1].clone()
        s = s - sT
        s = s * (s0 / (s0 - sT))
        ac = s ** 2
    return ac


def trailing_timesteps(num_inference_steps: int, num_train_timesteps: int = 1000) -> np.ndarray:
    """DPM:293-298."""
    ratio = num_train_timesteps / num_inference_steps
    return np.round(np.arange(num_train_timesteps, 0, -ratio)).astype(np.int64) - 1


def dynamic_guidance_scale(guidance_scale: float, t: int, num_inference_steps: int) -> float:
    """PIPE:991-994 (python floats)."""
    return 1 + guidance_scale * ((1 - math.cos(math.pi * ((num_inference_steps - t) / num_inference_steps) ** 5.0)) / 2)


@dataclass
class StepCoefficients:
    """The scalars of one scheduler step, float64 as the reference computes them."""
    sqrt_alpha_t: float
    sqrt_beta_t: float
    mult0: float
    mult1: float
    mult2: Optional[float]      # second order: 1 + 1 / (2 r)
    mult3: Optional[float]      # second order: 1 / (2 r)
    mult_noise: float
    second_order: bool


def step_coefficients(table: torch.Tensor, t: int, t_back: Optional[int], num_inference_steps: int, have_old: bool,
                      set_alpha_to_one: bool = True) -> StepCoefficients:
    """DPM:386-431 scalar part.  torch float64 0-dim arithmetic, exactly the reference's expressions."""
    n_train = table.shape[0]
    prev_t = t - n_train // num_inference_steps
    a_t = table[t]
    a_prev = table[prev_t] if prev_t >= 0 else (torch.tensor(1.0) if set_alpha_to_one else table[0])
    a_back = table[t_back] if t_back is not None else None
    lamb = ((a_t / (1 - a_t)) ** 0.5).log()
    lamb_next = ((a_prev / (1 - a_prev)) ** 0.5).log()
    h = lamb_next - lamb
    r = None
    if a_back is not None:
        lamb_prev = ((a_back / (1 - a_back)) ** 0.5).log()
        r = (lamb - lamb_prev) / h
    mult0 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * (-h).exp()
    mult1 = (-2 * h).expm1() * a_prev ** 0.5
    mult_noise = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
    second = have_old and prev_t >= 0
    m2 = m3 = None
    if a_back is not None:
        m2, m3 = float(1 + 1 / (2 * r)), float(1 / (2 * r))
    return StepCoefficients(float(a_t ** 0.5), float((1 - a_t) ** 0.5), float(mult0), float(mult1), m2, m3, float(mult_noise), second)


def _f32(x: float) -> torch.Tensor:
    return torch.tensor(x, dtype=torch.float64).to(torch.float32)


def cfg_combine(noise_pred: torch.Tensor, g: float) -> torch.Tensor:
    """PIPE:981, 995-997: fp32 of the transformer output, uncond + g * (text - uncond)."""
    u, c = noise_pred.float().chunk(2)
    return u + g * (c - u)


def dpm_step(model_output: torch.Tensor, old_pred: Optional[torch.Tensor], sample: torch.Tensor, noise1: torch.Tensor,
             noise2: Optional[torch.Tensor], co: StepCoefficients, prediction_type: str = "v_prediction"):
    """DPM:402-436.  model_output fp32; sample / noise in the latent dtype (bf16 in the pipeline); returns
    (prev_sample fp32, pred_original_sample fp32).  A 0-dim coefficient times a tensor keeps the tensor's dtype."""
    lat = sample.dtype
    # coefficient * tensor: the 0-dim float64 coefficient is cast to the TENSOR's dtype first (bf16 for the latents!), the
    # product is formed in fp32 and rounded once
    k = lambda c, x: (x.float() * torch.tensor(c, dtype=torch.float64).to(x.dtype).float()).to(x.dtype)   # noqa: E731
    if prediction_type == "v_prediction":
        pred = k(co.sqrt_alpha_t, sample).float() - k(co.sqrt_beta_t, model_output)
    elif prediction_type == "epsilon":
        pred = (sample.float() - k(co.sqrt_beta_t, model_output)) / _f32(co.sqrt_alpha_t)
    elif prediction_type == "sample":
        pred = model_output
    else:
        raise ValueError(prediction_type)
    if not co.second_order:
        prev = (k(co.mult0, sample).float() - k(co.mult1, pred)) + k(co.mult_noise, noise1).float()
        return prev, pred
    den = k(co.mult2, pred) - k(co.mult3, old_pred)
    prev = (k(co.mult0, sample).float() - k(co.mult1, den)) + k(co.mult_noise, noise2).float()
    assert lat == sample.dtype
    return prev, pred


def add_noise_coefficients(table: torch.Tensor, t: int, dtype: torch.dtype):
    """DPM:451-463: the table is cast to the latent dtype FIRST, then indexed, then ** 0.5 in that dtype."""
    a = table.to(dtype)[t]
    return float(a ** 0.5), float((1 - a) ** 0.5)


def replace_gt_blend(latents: torch.Tensor, gt_latents: torch.Tensor, noise0: torch.Tensor, mask: torch.Tensor,
                     sa: Optional[float], sb: Optional[float], mask_background: bool = False) -> torch.Tensor:
    """PIPE:1017-1034 in the latent dtype (every op rounds): proper = sa * gt + sb * noise0 (skipped on the last step),
    latents = (1 - mask) * proper + mask * latents (or the mirrored form)."""
    dt = latents.dtype
    proper = gt_latents
    if sa is not None:
        proper = torch.tensor(sa, dtype=dt) * gt_latents + torch.tensor(sb, dtype=dt) * noise0
    if mask_background:
        return mask * proper + (1 - mask) * latents
    return (1 - mask) * proper + mask * latents
