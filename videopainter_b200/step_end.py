"""Host side of the fused step end: the per-step scalars of `CogVideoXDPMScheduler.step`
(DPM:306-328, 386-422), of `add_noise` (DPM:451-463) and the dynamic guidance scale (PIPE:991-994), computed exactly as the
reference computes them (torch float64 0-dim arithmetic on the scheduler's own alpha table), and one call per denoise step
into `vp_step_end`.  Nothing here synchronises with the device: the timesteps are known on the host before the loop starts,
which also removes the two `t.item()` round trips of PIPE:983 / 993.

    se = StepEnd(scheduler.alphas_cumprod, scheduler.timesteps.tolist(), guidance_scale=6.0, use_dynamic_cfg=True)
    for i, t in enumerate(timesteps):
        noise_pred = transformer(...)[0]                                   # [2, F, C, H, W] bf16
        n1, n2 = randn_tensor(...), randn_tensor(...)                      # as the reference draws them (DPM:423, 431)
        latents, old = se(i, noise_pred, latents, old, n1, n2, gt=video_latents, noise0=noise, mask=init_mask)
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

from . import ops

BF16 = torch.bfloat16


def _bf(x: torch.Tensor) -> float:
    """0-dim float64 coefficient -> the value the reference multiplies a bf16 tensor with."""
    return float(x.to(BF16).float())


def _f32(x: torch.Tensor) -> float:
    return float(x.to(torch.float32))


class StepEnd:
    def __init__(self, alphas_cumprod: torch.Tensor, timesteps: Sequence[int], guidance_scale: float = 6.0,
                 use_dynamic_cfg: bool = True, set_alpha_to_one: bool = True, mask_background: bool = False,
                 prediction_type: str = "v_prediction"):
        if prediction_type != "v_prediction":
            raise NotImplementedError("the fused step end implements v-prediction (the CogVideoX-5B scheduler configuration)")
        self.table = alphas_cumprod.detach().to("cpu")
        self.timesteps: List[int] = [int(t) for t in timesteps]
        self.n = len(self.timesteps)
        self.guidance_scale = float(guidance_scale)
        self.use_dynamic_cfg = use_dynamic_cfg
        self.final_alpha = torch.tensor(1.0) if set_alpha_to_one else self.table[0]
        self.mask_background = mask_background
        # every scalar of the whole schedule is computed once, here (float64 torch arithmetic is ~100 us per step on the host)
        self._co = {(i, h): self._coefficients(i, h) for i in range(self.n) for h in (False, True)}
        self._rn = [self._renoise_coefficients(i) for i in range(self.n)]
        self._g = [self._guidance(i) for i in range(self.n)]

    def guidance(self, i: int) -> float:
        return self._g[i]

    def coefficients(self, i: int, have_old: bool):
        return self._co[(i, bool(have_old))]

    def renoise_coefficients(self, i: int):
        return self._rn[i]

    def _guidance(self, i: int) -> float:
        t = self.timesteps[i]
        if not self.use_dynamic_cfg:
            return self.guidance_scale
        return 1 + self.guidance_scale * ((1 - math.cos(math.pi * ((self.n - t) / self.n) ** 5.0)) / 2)     # PIPE:991-994

    def _coefficients(self, i: int, have_old: bool):
        """DPM:386-422 (scalars).  Returns the argument tuple of vp_step_end between `noise` and `pred_out`."""
        tab = self.table
        t = self.timesteps[i]
        t_back = self.timesteps[i - 1] if i > 0 else None
        prev_t = t - tab.shape[0] // self.n
        a_t = tab[t]
        a_prev = tab[prev_t] if prev_t >= 0 else self.final_alpha
        lamb = ((a_t / (1 - a_t)) ** 0.5).log()
        lamb_next = ((a_prev / (1 - a_prev)) ** 0.5).log()
        h = lamb_next - lamb
        m0 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * (-h).exp()
        m1 = (-2 * h).expm1() * a_prev ** 0.5
        mn = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
        second = have_old and prev_t >= 0 and t_back is not None
        m2 = m3 = 0.0
        if second:
            a_back = tab[t_back]
            r = (lamb - ((a_back / (1 - a_back)) ** 0.5).log()) / h
            m2, m3 = _f32(1 + 1 / (2 * r)), _f32(1 / (2 * r))
        return (_bf(a_t ** 0.5), _f32((1 - a_t) ** 0.5), _bf(m0), _f32(m1), m2, m3, _bf(mn), int(second))

    def _renoise_coefficients(self, i: int):
        """add_noise to the NEXT timestep (PIPE:1026-1030, DPM:451-463: table cast to bf16, indexed, ** 0.5 in bf16)."""
        if i >= self.n - 1:
            return 0.0, 0.0, 0
        a = self.table.to(BF16)[self.timesteps[i + 1]]
        return float(a ** 0.5), float((1 - a) ** 0.5), 1

    @torch.no_grad()
    def __call__(self, i: int, noise_pred: torch.Tensor, latents: torch.Tensor, old_pred: Optional[torch.Tensor],
                 noise_first: torch.Tensor, noise_second: Optional[torch.Tensor], gt: Optional[torch.Tensor] = None,
                 noise0: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None, want_prev_fp32: bool = False):
        """One step for ONE sample: noise_pred [2, F, C, H, W] bf16 (uncond, text), latents / noises / gt [1, F, C, H, W] bf16,
        old_pred fp32 or None, mask [1, F, 1, H, W] bf16.  Returns (latents bf16, pred_original_sample fp32[, prev fp32])."""
        if noise_pred.shape[0] != 2 or latents.shape[0] != 1:
            raise ValueError("step end runs the CFG pair of one sample: noise_pred [2, ...], latents [1, ...]")
        co = self.coefficients(i, old_pred is not None)
        second = co[-1]
        noise = noise_second if second else noise_first
        n = latents.numel()
        F_, C, H, W = latents.shape[1:]
        dev = latents.device
        pred = torch.empty(latents.shape, dtype=torch.float32, device=dev)
        prev = torch.empty(latents.shape, dtype=torch.float32, device=dev) if want_prev_fp32 else None
        out = torch.empty_like(latents)
        sa, sb, renoise = self.renoise_coefficients(i)
        p = ops._p
        ops._launch("step_end", p(noise_pred, BF16, "noise_pred"), float(self.guidance(i)), p(latents, BF16, "latents"),
                    p(old_pred, torch.float32, "old_pred") if second else None, p(noise, BF16, "noise"),
                    *[float(c) for c in co[:-1]], int(second), p(pred), p(prev), p(out), p(gt, BF16, "gt"),
                    p(noise0, BF16, "noise0") if gt is not None else None, p(mask, BF16, "mask") if gt is not None else None,
                    int(C), int(H * W), float(sa), float(sb), int(renoise), int(self.mask_background), int(n))
        ops.launch_count += 1
        return (out, pred, prev) if want_prev_fp32 else (out, pred)
