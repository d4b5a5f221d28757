"""Pins the oracle (oracle/cogvideox_oracle.py) to the reference: against the committed golden vectors
that oracle/make_golden.py produced by running the real reference modules, and — when /root/reference is
mounted (build container only) — against the live modules."""
import os

import pytest
import torch

from oracle import cogvideox_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = dict(rtol=2e-4, atol=2e-5)   # fp32 vs fp32, different op grouping (explicit softmax vs SDPA)


def _ck(sd):
    s = a = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        s += float(v.sum())
        a += float(v.abs().sum())
    return torch.tensor([s, a], dtype=torch.float64)


@pytest.mark.parametrize("name", ["tiny_step", "tiny_step_resample"])
def test_oracle_matches_reference_golden(name):
    rec = torch.load(os.path.join(GOLD, name + ".pt"))
    cfg = O.tiny_config(id_pool_resample_learnable=rec["resample"])
    cfg_b = O.tiny_config(num_layers=1)
    sd_t = O.init_state_dict(cfg, rec["seed_t"])
    sd_b = O.init_state_dict(cfg_b, rec["seed_b"], branch=True)
    # the regenerated weights must be the very weights the reference ran with
    assert torch.equal(_ck(sd_t), rec["ck_t"]) and torch.equal(_ck(sd_b), rec["ck_b"])
    inp = O.make_inputs(cfg, rec["seed_in"])
    with torch.no_grad():
        samples, (out, hs, rmask) = O.denoise_step(sd_t, sd_b, cfg, cfg_b, inp)
    for a, b in zip(samples, rec["branch_samples"]):
        torch.testing.assert_close(a, b, **TOL)
    torch.testing.assert_close(out, rec["noise_pred"], **TOL)
    assert torch.equal(rmask, rec["resample_mask"])
    torch.testing.assert_close(hs[-1], rec["hs_last"], **TOL)
    sums = torch.stack([h.double().sum() for h in hs])
    torch.testing.assert_close(sums, rec["hs_sums"], rtol=1e-4, atol=1e-2)
    # second window: prev_hidden_states / prev_clip_weight / prev_resample_mask (T3D:574-582, AP:2156-2189, 2247-2252)
    inp2 = O.make_inputs(cfg, 2)
    kw = dict(prev_hidden_states={i: h for i, h in enumerate(hs)}, prev_clip_weight=rec["prev_w"],
              prev_resample_mask=rmask)
    with torch.no_grad():
        _, (out2, hs2, rmask2) = O.denoise_step(sd_t, sd_b, cfg, cfg_b, inp2, attention_kwargs=kw)
    torch.testing.assert_close(out2, rec["noise_pred_w2"], **TOL)
    torch.testing.assert_close(hs2[-1], rec["hs_last_w2"], **TOL)
    assert torch.equal(rmask2, rec["resample_mask_w2"])


def test_oracle_block_full_width_golden():
    rec = torch.load(os.path.join(GOLD, "block_full_width.pt"))
    cfg = O.full_config(num_layers=1, sample_height=8, sample_width=8, max_text_seq_length=24)
    sd = O.init_state_dict(cfg, rec["seed"])
    g = torch.Generator().manual_seed(rec["seed_in"])
    Sv = cfg.latent_frames * 16
    h = torch.randn(1, Sv, 3072, generator=g)
    e = torch.randn(1, 24, 3072, generator=g)
    temb = torch.randn(1, 512, generator=g)
    rope = O.rope_3d(64, ((0, 0), (4, 4)), (4, 4), cfg.latent_frames)
    with torch.no_grad():
        ho, eo = O.block(sd, "transformer_blocks.0.", cfg, h, e, temb, rope)
    torch.testing.assert_close(ho[:, ::7, ::5], rec["h_out"], **TOL)
    torch.testing.assert_close(eo[:, ::3, ::5], rec["e_out"], **TOL)


def test_oracle_rope_and_pos_tables_shape():
    cfg = O.full_config()
    cos, sin = O.pipeline_rope(cfg, 480, 720, 13)
    assert cos.shape == (17550, 64) and sin.shape == (17550, 64)
    # repeat-interleaved pairs (EMB:641-642)
    assert torch.equal(cos[:, 0::2], cos[:, 1::2])
    # first token: position 0 on every axis
    assert torch.allclose(cos[0], torch.ones(64)) and torch.allclose(sin[0], torch.zeros(64))


@pytest.mark.skipif(not os.path.isdir("/root/reference/diffusers/src"), reason="reference not mounted")
def test_oracle_matches_live_reference_pos_tables():
    import sys
    sys.path.insert(0, "/root/reference/diffusers/src")
    from diffusers.models.embeddings import get_3d_rotary_pos_embed, CogVideoXPatchEmbed  # type: ignore
    cos, sin = get_3d_rotary_pos_embed(64, ((0, 0), (30, 45)), (30, 45), 13)
    c2, s2 = O.pipeline_rope(O.full_config(), 480, 720, 13)
    assert torch.equal(cos, c2) and torch.equal(sin, s2)
    cfg = O.tiny_config()
    pe = CogVideoXPatchEmbed(patch_size=2, in_channels=32, embed_dim=128, text_embed_dim=64, sample_width=8,
                             sample_height=8, sample_frames=49, max_text_seq_length=16,
                             use_positional_embeddings=False, use_learned_positional_embeddings=True)
    assert torch.equal(pe.pos_embedding, O.sincos_pos_embedding(cfg))


def test_oracle_matches_reference_variant_goldens():
    """tests/golden/tiny_variants.pt (oracle/make_golden.py run_variants, outputs of the real reference): window 2 with
    prev_clip_weight == 0.0, add_first with masks, fused QKV projections, the wo_text branch."""
    rec = torch.load(os.path.join(GOLD, "tiny_variants.pt"))
    for resample in (False, True):
        cfg, cfg_b = O.tiny_config(id_pool_resample_learnable=resample), O.tiny_config(num_layers=1)
        sd_t = O.init_state_dict(cfg, rec["seed_t"])
        sd_b = O.init_state_dict(cfg_b, rec["seed_b"], branch=True)
        inp, inp2 = O.make_inputs(cfg, rec["seed_in"]), O.make_inputs(cfg, 2)
        tag = "resample" if resample else "plain"
        with torch.no_grad():
            _, (out, hs, rmask) = O.denoise_step(sd_t, sd_b, cfg, cfg_b, inp)
            kw = dict(prev_hidden_states={i: h for i, h in enumerate(hs)}, prev_clip_weight=0.0, prev_resample_mask=rmask)
            _, (out2, _, _) = O.denoise_step(sd_t, sd_b, cfg, cfg_b, inp2, attention_kwargs=kw)
            torch.testing.assert_close(out2, rec[f"w2_prev0_{tag}"], **TOL)
            # prev_clip_weight == 0.0 must equal running window 2 with no previous states at all
            _, (out2n, _, _) = O.denoise_step(sd_t, sd_b, cfg, cfg_b, inp2)
            torch.testing.assert_close(out2, out2n, rtol=0, atol=0)
            lat = torch.cat([inp["latents"], inp["image_latents"]], dim=2)
            cond = torch.cat([inp["masked_latents"], inp["mask"]], dim=2)
            smp = O.branch_forward(sd_b, cfg_b, inp["latents"], inp["text"], cond, inp["timestep"], inp["rope"])
            outa = O.transformer_forward(sd_t, cfg, lat, inp["text"], inp["timestep"], inp["rope"], smp, inp["mask"][:, :, :1],
                                         add_first=True, return_hidden_states=True, return_resample_mask=True)[0]
            torch.testing.assert_close(outa, rec[f"add_first_masked_{tag}"], **TOL)
            if resample:
                lat2 = torch.cat([inp2["latents"], inp2["image_latents"]], dim=2)
                cond2 = torch.cat([inp2["masked_latents"], inp2["mask"]], dim=2)
                smp2 = O.branch_forward(sd_b, cfg_b, inp2["latents"], inp2["text"], cond2, inp2["timestep"], inp2["rope"])
                kw = dict(prev_hidden_states={i: h for i, h in enumerate(hs)}, prev_clip_weight=0.5, prev_resample_mask=rmask)
                outf = O.transformer_forward(sd_t, cfg, lat2, inp2["text"], inp2["timestep"], inp2["rope"], smp2,
                                             inp2["mask"][:, :, :1], attention_kwargs=kw, return_hidden_states=True,
                                             return_resample_mask=True, fused_qkv=True)[0]
                torch.testing.assert_close(outf, rec["fused_qkv_w2"], **TOL)
    cfg_b = O.tiny_config(num_layers=2)
    sd_b = O.init_state_dict(cfg_b, rec["wo_text_seed_b"], branch=True)
    inp = O.make_inputs(cfg_b, rec["seed_in"])
    cond = torch.cat([inp["masked_latents"], inp["mask"]], dim=2)
    with torch.no_grad():
        smp = O.branch_forward(sd_b, cfg_b, inp["latents"], inp["text"], cond, inp["timestep"], inp["rope"], conditioning_scale=0.7,
                               wo_text=True)
    for a, b in zip(smp, rec["wo_text_samples"]):
        torch.testing.assert_close(a, b, **TOL)
