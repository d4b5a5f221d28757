"""GPU parity of the corners of the two forwards that the first goldens do not reach, against outputs of the REAL reference
(tests/golden/tiny_variants.pt, oracle/make_golden.py run_variants) and against the fp32 oracle:
  - second window with prev_clip_weight == 0.0 (T3D:141-146 normalises the previous states, AP:2156 / 2247 then ignore them)
  - add_first=True together with masks (T3D:600-609)
  - fuse_qkv_projections() on a model built with the ID-resample processor (T3D:433-456, AP:2378-2436)
  - the wo_text branch (BR:407-412, T3D:186-216, AP:2316-2366)
  - a LoRA-carrying backbone (PEFT-style wrapped projections, alpha / r scaling, attention_kwargs["scale"]: T3D:490-498)
Tolerance as everywhere: cosine >= 0.9995, max-abs error <= 3 % of the reference's max-abs (bf16 path vs fp32 reference)."""
import os

import pytest
import torch

from _util import assert_close_bf16

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
GOLD = os.path.join(os.path.dirname(__file__), "golden")
COS_MIN, REL_MAX = 0.9995, 3e-2


def _mirror(cfg, sd, branch=False, **extra):
    import videopainter_b200 as vp
    kw = cfg.to_kwargs(); kw.pop("norm_eps")
    kw.update(extra)
    cls = vp.CogvideoXBranchModel if branch else vp.CogVideoXTransformer3DModel
    if branch:
        kw.pop("id_pool_resample_learnable", None)
    m = cls(**kw, device="cuda", dtype=BF16)
    m.load_state_dict({k: v.to(BF16) for k, v in sd.items()}, strict=True)
    return m


def _step(tr, br, inp, kw=None, add_first=False):
    lat = torch.cat([inp["latents"], inp["image_latents"]], dim=2).to(BF16)
    cond = torch.cat([inp["masked_latents"], inp["mask"]], dim=2).to(BF16)
    text = inp["text"].to(BF16)
    smp = br(hidden_states=inp["latents"].to(BF16), encoder_hidden_states=text, branch_cond=cond, timestep=inp["timestep"],
             image_rotary_emb=inp["rope"], return_dict=False)[0]
    return tr(hidden_states=lat, encoder_hidden_states=text, timestep=inp["timestep"], image_rotary_emb=inp["rope"],
              branch_block_samples=smp, attention_kwargs=kw, branch_block_masks=inp["mask"][:, :, :1].to(BF16), add_first=add_first,
              return_hidden_states=True, return_resample_mask=True, return_dict=False)


@pytest.mark.parametrize("resample", [False, True])
def test_variants_against_reference_golden(resample):
    from oracle import cogvideox_oracle as O
    rec = torch.load(os.path.join(GOLD, "tiny_variants.pt"))
    cfg, cfg_b = O.tiny_config(id_pool_resample_learnable=resample), O.tiny_config(num_layers=1)
    tr = _mirror(cfg, O.init_state_dict(cfg, rec["seed_t"]))
    br = _mirror(cfg_b, O.init_state_dict(cfg_b, rec["seed_b"], branch=True), branch=True)
    inp, inp2 = O.make_inputs(cfg, rec["seed_in"], device="cuda"), O.make_inputs(cfg, 2, device="cuda")
    tag = "resample" if resample else "plain"
    out, hs, rmask = _step(tr, br, inp)
    kw = dict(prev_hidden_states={i: h for i, h in enumerate(hs)}, prev_clip_weight=0.0, prev_resample_mask=rmask)
    out2 = _step(tr, br, inp2, kw)[0]
    assert_close_bf16(f"w2 prev_clip_weight=0 ({tag})", out2.cpu(), rec[f"w2_prev0_{tag}"], COS_MIN, REL_MAX)
    assert torch.equal(out2, _step(tr, br, inp2)[0])          # bit-identical to a window without previous states
    outa = _step(tr, br, inp, add_first=True)[0]
    assert_close_bf16(f"add_first with masks ({tag})", outa.cpu(), rec[f"add_first_masked_{tag}"], COS_MIN, REL_MAX)
    if resample:
        tr.fuse_qkv_projections()
        kw = dict(prev_hidden_states={i: h for i, h in enumerate(hs)}, prev_clip_weight=0.5, prev_resample_mask=rmask)
        outf = _step(tr, br, inp2, kw)[0]
        assert_close_bf16("fused QKV, window 2", outf.cpu(), rec["fused_qkv_w2"], COS_MIN, REL_MAX)
        tr.unfuse_qkv_projections()
        assert not torch.equal(_step(tr, br, inp2, kw)[0], outf)        # back on the resample processor


def test_wo_text_branch_against_reference_golden():
    from oracle import cogvideox_oracle as O
    rec = torch.load(os.path.join(GOLD, "tiny_variants.pt"))
    cfg_b = O.tiny_config(num_layers=2)
    br = _mirror(cfg_b, O.init_state_dict(cfg_b, rec["wo_text_seed_b"], branch=True), branch=True, wo_text=True)
    inp = O.make_inputs(cfg_b, rec["seed_in"], device="cuda")
    cond = torch.cat([inp["masked_latents"], inp["mask"]], dim=2).to(BF16)
    smp = br(hidden_states=inp["latents"].to(BF16), encoder_hidden_states=inp["text"].to(BF16), branch_cond=cond,
             timestep=inp["timestep"], image_rotary_emb=inp["rope"], conditioning_scale=0.7, wo_text=True, return_dict=False)[0]
    for i, (a, b) in enumerate(zip(smp, rec["wo_text_samples"])):
        assert_close_bf16(f"wo_text branch sample {i}", a.cpu(), b, COS_MIN, REL_MAX)
    with pytest.raises(ValueError):                              # the reference cannot run this combination either
        br(hidden_states=inp["latents"].to(BF16), encoder_hidden_states=inp["text"].to(BF16), branch_cond=cond,
           timestep=inp["timestep"], image_rotary_emb=inp["rope"], wo_text=False, return_dict=False)


@pytest.mark.parametrize("resample", [False, True])
def test_lora_backbone_against_oracle(resample):
    """PEFT-style LoRA on to_q / to_k / to_v / to_out.0 (TRAINID:1520-1526) of a production-width model: two adapters, one
    inactive, scaling = alpha / r = 0.5, attention_kwargs['scale'] = 0.5 -> the oracle runs with W + 0.25 B A."""
    from _fake_peft import inject
    from oracle import cogvideox_oracle as O
    cfg = O.full_config(num_layers=2, sample_height=16, sample_width=24, id_pool_resample_learnable=resample)
    cfg_b = O.full_config(num_layers=1, sample_height=16, sample_width=24)
    sd_t = O.init_state_dict(cfg, 51, device="cuda")
    sd_b = O.init_state_dict(cfg_b, 52, branch=True, device="cuda")
    tr, br = _mirror(cfg, sd_t), _mirror(cfg_b, sd_b, branch=True)
    base = _step(tr, br, O.make_inputs(cfg, 6, device="cuda", rect_mask=True))[0]
    lora = inject(tr, r=16, lora_alpha=8, seed=3, adapters=("id", "unused"), inactive=("unused",))
    inp = O.make_inputs(cfg, 6, device="cuda", rect_mask=True)
    out, hs, rmask = _step(tr, br, inp, kw={"scale": 0.5})
    assert not torch.equal(out, base)
    torch.backends.cuda.matmul.allow_tf32 = False
    r32 = lambda sd: {k: v.to(BF16).float() for k, v in sd.items()}   # noqa: E731
    merged = O.lora_merge(r32(sd_t), {k: tuple(t.cuda() for t in v["id"]) for k, v in lora.items()}, scale=0.25)
    with torch.no_grad():
        _, (rout, rhs, rrm) = O.denoise_step(merged, r32(sd_b), cfg, cfg_b, inp, head_chunk=8)
    assert torch.equal(rmask, rrm)
    assert_close_bf16("lora noise_pred", out, rout, COS_MIN, REL_MAX)
    assert_close_bf16("lora hs_last", hs[-1], rhs[-1], COS_MIN, REL_MAX)
    # scale 1.0 is a different model (and a different packed-weight cache entry); disabling the adapters restores the base
    out1 = _step(tr, br, inp)[0]
    assert not torch.equal(out1, out)
    for m in tr.modules():
        if hasattr(m, "lora_A"):
            m.disable_adapters = True
    assert torch.equal(_step(tr, br, inp)[0], base)
