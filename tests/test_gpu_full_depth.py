"""Parity of the BENCHMARKED configuration (BASELINE.json configs[1], north-star acceptance): CogVideoX-5B-I2V depth and size —
42 backbone layers + 2 branch layers, 49x480x720 -> 17 776 tokens, CFG batch 2 — one denoise step at t in {999, 499, 19}
against the fp32 oracle executed on the same GPU (bf16-rounded weights, TF32 off).

Tolerance (stated by BASELINE.json north_star / SURVEY.md §8c): noise-pred (and branch-sample) cosine >= 0.9995 and max-abs
error <= 3 % of the fp32 result's max-abs.  The same oracle executed in bf16 (the arithmetic the reference runs: every op
rounded to bf16) is printed beside our numbers so the bound can be read against the reference's own bf16-vs-fp32 error.
The last hidden state (the 42-layer residual stream, a few large outlier channels) is held to cosine >= 0.9995 and to a
max-abs error no larger than max(3 %, what the reference's own bf16 arithmetic shows on the same inputs) — first measured
run: ours 3.5 %, reference-in-bf16 4.3 %; noise-pred ours 1.7 %, reference-in-bf16 2.4 %.

The run writes its numbers to gpurun_out/full_depth_parity.json when that directory exists (copied into profiles/)."""
import json
import os

import pytest
import torch

from _util import cos_sim

pytestmark = [pytest.mark.gpu, pytest.mark.slow]
BF16 = torch.bfloat16
COS_MIN = 0.9995
REL_MAX = 3e-2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(got, ref):
    g, r = got.float(), ref.float()
    return float((g - r).abs().max() / (r.abs().max() + 1e-30))


@pytest.fixture(scope="module")
def stack():
    import videopainter_b200 as vp
    from oracle import cogvideox_oracle as O
    free, _ = torch.cuda.mem_get_info()
    if free < 100e9:
        pytest.skip("needs ~90 GB of device memory (fp32 oracle weights + bf16 model + fp32 score chunks)")
    cfg, cfg_b = O.full_config(), O.full_config(num_layers=2)
    sd_t = O.init_state_dict(cfg, 101, device="cuda")
    sd_b = O.init_state_dict(cfg_b, 102, branch=True, device="cuda")
    kw = cfg.to_kwargs(); kw.pop("norm_eps")
    kwb = cfg_b.to_kwargs(); kwb.pop("norm_eps")
    tr = vp.CogVideoXTransformer3DModel(**kw, device="cuda", dtype=BF16)
    br = vp.CogvideoXBranchModel(**kwb, device="cuda", dtype=BF16)
    tr.load_state_dict({k: v.to(BF16) for k, v in sd_t.items()}, strict=True)
    br.load_state_dict({k: v.to(BF16) for k, v in sd_b.items()}, strict=True)
    for sd in (sd_t, sd_b):                       # the checkpoint IS bf16: the fp32 oracle gets the bf16-rounded values
        for v in sd.values():
            v.copy_(v.to(BF16).float())
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    records = []
    yield O, cfg, cfg_b, sd_t, sd_b, tr, br, records
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir) and records:
        with open(os.path.join(out_dir, "full_depth_parity.json"), "w") as f:
            json.dump(records, f, indent=1)


@pytest.mark.parametrize("t", [999, 499, 19])
def test_full_depth_step_against_fp32_oracle(stack, t):
    O, cfg, cfg_b, sd_t, sd_b, tr, br, records = stack
    inp = O.make_inputs(cfg, 7, device="cuda", rect_mask=True)
    inp["timestep"] = torch.full((2,), t, dtype=torch.int64, device="cuda")
    lat_in = torch.cat([inp["latents"], inp["image_latents"]], dim=2).to(BF16)
    cond = torch.cat([inp["masked_latents"], inp["mask"]], dim=2).to(BF16)
    text = inp["text"].to(BF16)
    with torch.no_grad():
        samples = br(hidden_states=inp["latents"].to(BF16), encoder_hidden_states=text, branch_cond=cond,
                     timestep=inp["timestep"], image_rotary_emb=inp["rope"], return_dict=False)[0]
        out, hs, rmask = tr(hidden_states=lat_in, encoder_hidden_states=text, timestep=inp["timestep"],
                            image_rotary_emb=inp["rope"], branch_block_samples=samples,
                            branch_block_masks=inp["mask"][:, :, :1].to(BF16), return_hidden_states=True,
                            return_resample_mask=True, return_dict=False)
        hs_last = hs[-1].clone()
        del hs
        # ---- checker 1: the oracle in fp32 ----
        rs, (rout, rhs, rrm) = O.denoise_step(sd_t, sd_b, cfg, cfg_b, inp, head_chunk=4)
        rhs_last = rhs[-1]
        del rhs
        # ---- for scale: the same oracle executed in bf16 (what the reference's eager path computes) ----
        sdt16 = dict(tr.state_dict())
        sdb16 = dict(br.state_dict())
        bs, (bout, bhs, _) = O.denoise_step(sdt16, sdb16, cfg, cfg_b, inp, dtype=BF16, head_chunk=8)
        bhs_last = bhs[-1]
        del bhs
    assert out.shape == (2, 13, 16, 60, 90) and hs_last.shape == (2, 17776, 3072)
    assert torch.equal(rmask, rrm)                                   # index / mask work: bit-exact
    rec = {"t": t, "layers": "42+2", "tokens": 17776, "cfg_batch": 2}
    for name, got, ref, b16 in (("noise_pred", out, rout, bout), ("hs_last", hs_last, rhs_last, bhs_last),
                                ("branch_sample_0", samples[0], rs[0], bs[0]), ("branch_sample_1", samples[1], rs[1], bs[1])):
        rec[name] = {"cos": cos_sim(got, ref), "rel_max": _rel(got, ref),
                     "reference_bf16_cos": cos_sim(b16, ref), "reference_bf16_rel_max": _rel(b16, ref),
                     "nan": int(torch.isnan(got.float()).sum())}
        print(f"[full-depth t={t}] {name}: ours cos={rec[name]['cos']:.7f} rel_max={rec[name]['rel_max']:.4e} | "
              f"oracle-in-bf16 cos={rec[name]['reference_bf16_cos']:.7f} rel_max={rec[name]['reference_bf16_rel_max']:.4e}")
    records.append(rec)
    for name in ("noise_pred", "hs_last", "branch_sample_0", "branch_sample_1"):
        assert rec[name]["nan"] == 0, rec
        assert rec[name]["cos"] >= COS_MIN, rec
        bound = REL_MAX if name != "hs_last" else max(REL_MAX, rec[name]["reference_bf16_rel_max"])
        assert rec[name]["rel_max"] <= bound, rec
