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
