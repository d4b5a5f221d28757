"""GPU parity of the fused step end (vp_step_end through videopainter_b200.step_end.StepEnd) against the vectors produced
by the REAL CogVideoXDPMScheduler / pipeline arithmetic (tests/golden/dpm_steps.pt): bit-exact, chained over whole
schedules (the output latents and pred_original_sample of step i feed step i + 1)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "dpm_steps.pt")


def test_step_end_bit_exact_over_two_schedules():
    from videopainter_b200.step_end import StepEnd
    rec = torch.load(GOLD)
    dev = "cuda"
    for run in rec["runs"]:
        ts = run["timesteps"].tolist()
        se = StepEnd(run["table"], ts, guidance_scale=run["guidance_scale"], use_dynamic_cfg=True)
        gt, noise0, mask = run["gt"].to(dev), run["noise0"].to(dev), run["mask"].to(dev)
        latents = run["latents0"].to(dev)
        old = None
        for i, st in enumerate(run["steps"]):
            assert torch.equal(latents.cpu(), st["latents_in"])
            n1 = st["noises"][0].to(dev)
            n2 = st["noises"][1].to(dev) if len(st["noises"]) > 1 else None
            latents, old, prev = se(i, st["noise_pred"].to(dev), latents, old, n1, n2, gt=gt, noise0=noise0, mask=mask,
                                    want_prev_fp32=True)
            torch.cuda.synchronize()
            assert torch.equal(old.cpu(), st["pred_original"]), (run["num_inference_steps"], i)
            assert torch.equal(prev.cpu(), st["prev_sample"]), (run["num_inference_steps"], i)
            assert torch.equal(latents.cpu(), st["latents_out"]), (run["num_inference_steps"], i)


def test_step_end_without_replace_gt_and_full_size():
    """No blend: latents = bf16(prev_sample); and the production latent size [1, 13, 16, 60, 90] against the oracle."""
    from oracle import dpm_oracle as D
    from videopainter_b200.step_end import StepEnd
    rec = torch.load(GOLD)
    run = rec["runs"][0]
    ts = run["timesteps"].tolist()
    se = StepEnd(run["table"], ts, guidance_scale=6.0, use_dynamic_cfg=True)
    st = run["steps"][2]
    out, pred = se(2, st["noise_pred"].cuda(), st["latents_in"].cuda(), st["old_in"].cuda(), st["noises"][0].cuda(), st["noises"][1].cuda())
    assert torch.equal(out.cpu(), st["stepped_bf16"]) and torch.equal(pred.cpu(), st["pred_original"])
    g = torch.Generator().manual_seed(3)
    shape = (1, 13, 16, 60, 90)
    bf = torch.bfloat16
    lat, gt, n0, n1, n2 = (torch.randn(shape, generator=g).to(bf) for _ in range(5))
    npred = torch.randn((2,) + shape[1:], generator=g).to(bf)
    old = torch.randn(shape, generator=g)
    mask = (torch.rand((1, 13, 1, 60, 90), generator=g) > 0.4).to(bf)
    i = 3
    out, pred = se(i, npred.cuda(), lat.cuda(), old.cuda(), n1.cuda(), n2.cuda(), gt=gt.cuda(), noise0=n0.cuda(), mask=mask.cuda())
    co = D.step_coefficients(run["table"], ts[i], ts[i - 1], len(ts), have_old=True)
    mo = D.cfg_combine(npred, D.dynamic_guidance_scale(6.0, ts[i], len(ts)))
    prev_ref, pred_ref = D.dpm_step(mo, old, lat, n1, n2, co)
    sa, sb = D.add_noise_coefficients(run["table"], ts[i + 1], bf)
    ref = D.replace_gt_blend(prev_ref.to(bf), gt, n0, mask, sa, sb)
    assert torch.equal(pred.cpu(), pred_ref) and torch.equal(out.cpu(), ref)


def test_step_end_against_the_real_scheduler_running_on_the_gpu():
    """The pipeline runs the scheduler with its alpha table on the CPU and the latents on the GPU (PIPE:1003-1034); the golden
    vectors of the test above come from the scheduler executing on CPU tensors.  Here the real `CogVideoXDPMScheduler`
    (baseline/_ref or /root/reference) steps CUDA tensors through the pipeline's own arithmetic and `StepEnd` must reproduce
    it bit for bit, chained over whole schedules.  (First run of this test: the CUDA execution rounds a 0-dim CPU coefficient
    to bf16 before it multiplies a bf16 tensor, exactly like the CPU execution — an fp32 multiply with the unrounded
    coefficient mismatched on 8 160 elements — so one set of coefficients serves both.)"""
    import math
    import pipeline_acceptance as PA
    if PA.reference_path() is None:
        pytest.skip("reference scheduler not available (no /root/reference, no baseline/_ref)")
    PA.load_reference()
    from diffusers import CogVideoXDPMScheduler  # type: ignore
    import diffusers.schedulers.scheduling_dpm_cogvideox as mod  # type: ignore
    from videopainter_b200.step_end import StepEnd
    dev, bf16 = "cuda", torch.bfloat16
    sch = CogVideoXDPMScheduler(snr_shift_scale=1.0, prediction_type="v_prediction", rescale_betas_zero_snr=True,
                                timestep_spacing="trailing", clip_sample=False, beta_schedule="scaled_linear",
                                beta_start=0.00085, beta_end=0.012, set_alpha_to_one=True)
    for n_steps in (6, 4):
        sch.set_timesteps(n_steps)
        timesteps = sch.timesteps
        g = torch.Generator().manual_seed(7)
        shape = (1, 4, 16, 8, 8)
        latents = torch.randn(shape, generator=g).to(bf16).to(dev)
        gt = torch.randn(shape, generator=g).to(bf16).to(dev)
        noise0 = torch.randn(shape, generator=g).to(bf16).to(dev)
        mask = (torch.rand((1, 4, 1, 8, 8), generator=g) > 0.5).to(bf16).to(dev)
        drawn = []
        real_randn = mod.randn_tensor

        def spy(*a, **k):
            t = real_randn(*a, **k)
            drawn.append(t.clone())
            return t
        mod.randn_tensor = spy
        try:
            gen = torch.Generator().manual_seed(42)
            se = StepEnd(sch.alphas_cumprod, timesteps.tolist(), guidance_scale=6.0, use_dynamic_cfg=True)
            ours_lat, ours_old, old = latents.clone(), None, None
            for i, t in enumerate(timesteps):
                noise_pred_bf16 = torch.randn((2,) + shape[1:], generator=g).to(bf16).to(dev)
                noise_pred = noise_pred_bf16.float()
                gs = 1 + 6.0 * ((1 - math.cos(math.pi * ((n_steps - t.item()) / n_steps) ** 5.0)) / 2)            # PIPE:991-994
                u, c = noise_pred.chunk(2)
                mo = u + gs * (c - u)
                drawn.clear()
                latents_f, old = sch.step(mo, old, t, timesteps[i - 1] if i > 0 else None, latents, generator=gen,
                                          return_dict=False)
                lat = latents_f.to(bf16)
                proper = gt
                if i < len(timesteps) - 1:
                    proper = sch.add_noise(gt, noise0, torch.tensor([timesteps[i + 1]]))
                latents = (1 - mask) * proper + mask * lat                                                        # PIPE:1031-1034
                n1 = drawn[0].to(dev)
                n2 = drawn[1].to(dev) if len(drawn) > 1 else None
                ours_lat, ours_old = se(i, noise_pred_bf16, ours_lat, ours_old, n1, n2, gt=gt, noise0=noise0, mask=mask)
                torch.cuda.synchronize()
                assert torch.equal(ours_old, old), (n_steps, i, float((ours_old - old).abs().max()))
                assert torch.equal(ours_lat, latents), (n_steps, i, float((ours_lat.float() - latents.float()).abs().max()))
        finally:
            mod.randn_tensor = real_randn
