"""ORACLE TOOLING — generates tests/golden/dpm_steps.pt by running the REAL CogVideoXDPMScheduler (build container only):

    python oracle/make_golden_dpm.py

A 6-step trailing schedule on a small latent, bf16 latents, fp32 model outputs, the pipeline's call pattern (PIPE:981-1034):
CFG combine with the dynamic scale, scheduler.step with the carried old_pred_original_sample, replace_gt blend.  The
generator-drawn noises are stored, so that the checker does not depend on the RNG implementation."""
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/diffusers/src")


def main():
    from diffusers import CogVideoXDPMScheduler  # type: ignore
    import diffusers.schedulers.scheduling_dpm_cogvideox as mod  # type: ignore
    sch = CogVideoXDPMScheduler(snr_shift_scale=1.0, prediction_type="v_prediction", rescale_betas_zero_snr=True,
                                timestep_spacing="trailing", clip_sample=False, beta_schedule="scaled_linear",
                                beta_start=0.00085, beta_end=0.012, set_alpha_to_one=True)
    runs = []
    for n_steps in (6, 4):            # 6: every later step is second order; 4: the last step falls off the table (final alpha)
        runs.append(one_run(sch, mod, n_steps))
    out = os.path.join(ROOT, "tests", "golden", "dpm_steps.pt")
    torch.save({"runs": runs}, out)
    print("wrote", out, [[s["t"] for s in r["steps"]] for r in runs], [[len(s["noises"]) for s in r["steps"]] for r in runs])


def one_run(sch, mod, n_steps):
    sch.set_timesteps(n_steps)
    timesteps = sch.timesteps
    g = torch.Generator().manual_seed(7)
    shape = (1, 4, 16, 8, 8)
    bf16 = torch.bfloat16
    latents = torch.randn(shape, generator=g).to(bf16)
    gt = torch.randn(shape, generator=g).to(bf16)
    noise0 = torch.randn(shape, generator=g).to(bf16)
    mask = (torch.rand((1, 4, 1, 8, 8), generator=g) > 0.5).to(bf16)
    guidance_scale = 6.0
    # capture the noises the scheduler draws
    drawn = []
    real_randn = mod.randn_tensor

    def spy(*a, **k):
        t = real_randn(*a, **k)
        drawn.append(t.clone())
        return t
    mod.randn_tensor = spy
    gen = torch.Generator().manual_seed(42)
    rec = {"num_inference_steps": n_steps, "table": sch.alphas_cumprod.clone(), "timesteps": timesteps.clone(), "latents0": latents.clone(), "gt": gt, "noise0": noise0,
           "mask": mask, "guidance_scale": guidance_scale, "steps": []}
    old = None
    for i, t in enumerate(timesteps):
        noise_pred_bf16 = torch.randn((2,) + shape[1:], generator=g).to(bf16)       # what the transformer returns
        noise_pred = noise_pred_bf16.float()
        gs = 1 + guidance_scale * ((1 - math.cos(math.pi * ((n_steps - t.item()) / n_steps) ** 5.0)) / 2)
        u, c = noise_pred.chunk(2)
        mo = u + gs * (c - u)
        drawn.clear()
        lat_in = latents.clone()
        latents_f, new_old = sch.step(mo, old, t, timesteps[i - 1] if i > 0 else None, latents, generator=gen, return_dict=False)
        stepped = latents_f.to(bf16)
        lat = stepped
        proper = gt
        if i < len(timesteps) - 1:
            proper = sch.add_noise(gt, noise0, torch.tensor([timesteps[i + 1]]))
        lat = (1 - mask) * proper + mask * lat
        rec["steps"].append({"t": int(t), "noise_pred": noise_pred_bf16, "guidance": gs, "model_output": mo, "latents_in": lat_in,
                             "old_in": None if old is None else old.clone(), "noises": [d.clone() for d in drawn],
                             "prev_sample": latents_f.clone(), "pred_original": new_old.clone(), "stepped_bf16": stepped.clone(),
                             "latents_out": lat.clone()})
        old = new_old
        latents = lat
    mod.randn_tensor = real_randn
    return rec


if __name__ == "__main__":
    main()
