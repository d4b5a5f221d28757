"""The CPU oracle of the step-end arithmetic (oracle/dpm_oracle.py: CFG combine, CogVideoXDPMScheduler.step, replace_gt
blend) against vectors produced by the REAL scheduler (oracle/make_golden_dpm.py).  Everything here is bit-exact: the
oracle mirrors the reference's dtype promotions and op order."""
import os

import torch

from oracle import dpm_oracle as D

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dpm_steps.pt")


def test_alpha_table_and_timesteps_match_the_scheduler():
    rec = torch.load(GOLD)
    for run in rec["runs"]:
        assert torch.equal(D.alphas_cumprod_table(), run["table"])
        assert run["table"].dtype == torch.float64 and float(run["table"][-1]) == 0.0      # zero terminal SNR
        assert D.trailing_timesteps(run["num_inference_steps"]).tolist() == run["timesteps"].tolist()


def test_step_end_oracle_is_bit_exact():
    rec = torch.load(GOLD)
    for run in rec["runs"]:
        n = run["num_inference_steps"]
        ts = run["timesteps"].tolist()
        for i, st in enumerate(run["steps"]):
            g = D.dynamic_guidance_scale(run["guidance_scale"], st["t"], n)
            assert g == st["guidance"]
            mo = D.cfg_combine(st["noise_pred"], g)
            assert torch.equal(mo, st["model_output"])
            co = D.step_coefficients(run["table"], st["t"], ts[i - 1] if i > 0 else None, n, have_old=st["old_in"] is not None)
            assert co.second_order == (len(st["noises"]) == 2)
            prev, pred = D.dpm_step(mo, st["old_in"], st["latents_in"], st["noises"][0], st["noises"][1] if co.second_order else None, co)
            assert torch.equal(pred, st["pred_original"]), (n, i)
            assert torch.equal(prev, st["prev_sample"]), (n, i)
            sa = sb = None
            if i < len(ts) - 1:
                sa, sb = D.add_noise_coefficients(run["table"], ts[i + 1], torch.bfloat16)
            out = D.replace_gt_blend(prev.to(torch.bfloat16), run["gt"], run["noise0"], run["mask"], sa, sb)
            assert torch.equal(out, st["latents_out"]), (n, i)


def test_host_coefficients_equal_the_oracle():
    """videopainter_b200.step_end computes the same scalars (host logic, no GPU) — product code and oracle are separate
    restatements of DPM:306-328 / 386-422 / 451-463 and PIPE:991-994."""
    from videopainter_b200.step_end import StepEnd
    rec = torch.load(GOLD)
    bf = lambda x: float(torch.tensor(x, dtype=torch.float64).to(torch.bfloat16).float())     # noqa: E731
    f32 = lambda x: float(torch.tensor(x, dtype=torch.float64).to(torch.float32))              # noqa: E731
    for run in rec["runs"]:
        n, ts = run["num_inference_steps"], run["timesteps"].tolist()
        se = StepEnd(run["table"], ts, guidance_scale=run["guidance_scale"], use_dynamic_cfg=True)
        for i, st in enumerate(run["steps"]):
            assert se.guidance(i) == st["guidance"]
            have_old = st["old_in"] is not None
            co = D.step_coefficients(run["table"], st["t"], ts[i - 1] if i > 0 else None, n, have_old)
            got = se.coefficients(i, have_old)
            want = (bf(co.sqrt_alpha_t), f32(co.sqrt_beta_t), bf(co.mult0), f32(co.mult1),
                    f32(co.mult2) if co.second_order else 0.0, f32(co.mult3) if co.second_order else 0.0, bf(co.mult_noise),
                    int(co.second_order))
            assert got == want, (n, i, got, want)
            sa, sb, renoise = se.renoise_coefficients(i)
            if i < len(ts) - 1:
                assert (sa, sb, renoise) == (*D.add_noise_coefficients(run["table"], ts[i + 1], torch.bfloat16), 1)
            else:
                assert renoise == 0
