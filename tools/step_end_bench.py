"""SURVEY §8f N1 measurement: the fused step end (one launch) against the reference's eager sequence of the same arithmetic
(PIPE:981-1034 + CogVideoXDPMScheduler.step; written out below with torch ops as the pipeline executes them, coefficients
precomputed so that neither side synchronises) at the production latent size.  `python tools/step_end_bench.py`"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BF = torch.bfloat16


def main():
    from videopainter_b200.step_end import StepEnd
    dev = "cuda"
    n_steps = 50
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    ac = torch.cumprod(1.0 - betas, dim=0)
    s = ac.sqrt(); s0, sT = s[0].clone(), s[-1].clone(); ac = ((s - sT) * (s0 / (s0 - sT))) ** 2
    import numpy as np
    ts = (np.round(np.arange(1000, 0, -1000 / n_steps)).astype(np.int64) - 1).tolist()
    se = StepEnd(ac, ts, guidance_scale=6.0, use_dynamic_cfg=True)
    g = torch.Generator(device=dev).manual_seed(0)
    shape = (1, 13, 16, 60, 90)
    lat, gt, n0, n1, n2 = (torch.randn(shape, device=dev, generator=g).to(BF) for _ in range(5))
    npred = torch.randn((2,) + shape[1:], device=dev, generator=g).to(BF)
    old = torch.randn(shape, device=dev, generator=g)
    mask = (torch.rand((1, 13, 1, 60, 90), device=dev, generator=g) > 0.4).to(BF)
    i = 10
    co = se.coefficients(i, True)
    sa, sb, _ = se.renoise_coefficients(i)
    gsc = se.guidance(i)
    c = [torch.tensor(v, dtype=torch.float64) for v in co[:-1]]      # 0-dim CPU tensors, as in the reference

    def eager():
        npf = npred.float()
        u, cc = npf.chunk(2)
        mo = u + gsc * (cc - u)
        pred = c[0] * lat - c[1] * mo
        den = c[4] * pred - c[5] * old
        prev = c[2] * lat - c[3] * den + c[6] * n2
        x = prev.to(BF)
        proper = torch.tensor(sa, dtype=BF) * gt + torch.tensor(sb, dtype=BF) * n0
        return (1 - mask) * proper + mask * x, pred

    def fused():
        return se(i, npred, lat, old, n1, n2, gt=gt, noise0=n0, mask=mask)

    a, b = eager(), fused()
    torch.cuda.synchronize()
    same = torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    res = {"identical": bool(same)}
    for name, fn in (("eager_reference_sequence", eager), ("fused_vp_step_end", fused)):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name + "_us"] = e0.elapsed_time(e1) * 1000 / 50
    n = lat.numel()
    res["algorithmic_bytes"] = n * (2 * 2 + 2 + 4 + 2 + 4 + 2 + 2 + 2 + 2) + mask.numel() * 2   # npred x2, lat, old, noise, pred, out, gt, noise0
    res["fused_gbs"] = res["algorithmic_bytes"] / res["fused_vp_step_end_us"] / 1e3
    print(json.dumps(res))


if __name__ == "__main__":
    main()
