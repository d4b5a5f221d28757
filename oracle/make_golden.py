"""ORACLE TOOLING — generates tests/golden/*.pt by running the REAL reference modules.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

For every case it (1) builds the reference `CogVideoXTransformer3DModel` / `CogvideoXBranchModel`
from /root/reference/diffusers/src, (2) loads the oracle's seeded state-dict with strict=True (which
also pins the parameter names), (3) runs the reference forward in fp32 on seeded inputs and (4)
stores outputs (+ weight checksums, so a regenerated state-dict can be verified bit-for-bit).
Weights are *not* stored: they are regenerated from the seed by `init_state_dict`.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF_SRC = "/root/reference/diffusers/src"

from oracle import cogvideox_oracle as O  # noqa: E402


def import_reference():
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from diffusers import CogVideoXTransformer3DModel, CogvideoXBranchModel  # type: ignore
    return CogVideoXTransformer3DModel, CogvideoXBranchModel


def ref_kwargs(cfg: O.OracleConfig) -> dict:
    kw = cfg.to_kwargs()
    kw.pop("norm_eps")
    return kw


def build_reference(cfg: O.OracleConfig, cfg_b: O.OracleConfig, seed_t: int, seed_b: int):
    T3D, BR = import_reference()
    sd_t = O.init_state_dict(cfg, seed_t)
    sd_b = O.init_state_dict(cfg_b, seed_b, branch=True)
    tr = T3D(**ref_kwargs(cfg)).eval()
    tr.load_state_dict(sd_t, strict=True)
    kwb = ref_kwargs(cfg_b)
    br = BR(**kwb).eval()
    br.load_state_dict(sd_b, strict=True)
    return tr, br, sd_t, sd_b


def checksum(sd):
    s = 0.0
    a = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        s += float(v.sum())
        a += float(v.abs().sum())
    return torch.tensor([s, a], dtype=torch.float64)


@torch.no_grad()
def run_case(name, resample: bool, two_windows: bool, prev_w: float):
    cfg = O.tiny_config(id_pool_resample_learnable=resample)
    cfg_b = O.tiny_config(num_layers=1)
    tr, br, sd_t, sd_b = build_reference(cfg, cfg_b, 11, 12)
    inp = O.make_inputs(cfg, seed=1)
    lat_in = torch.cat([inp["latents"], inp["image_latents"]], dim=2)
    cond = torch.cat([inp["masked_latents"], inp["mask"]], dim=2)
    samples = br(hidden_states=inp["latents"], encoder_hidden_states=inp["text"], branch_cond=cond,
                 timestep=inp["timestep"], image_rotary_emb=inp["rope"], return_dict=False)[0]
    out, hs, rmask = tr(hidden_states=lat_in, encoder_hidden_states=inp["text"], timestep=inp["timestep"],
                        image_rotary_emb=inp["rope"], branch_block_samples=samples,
                        branch_block_masks=inp["mask"][:, :, :1], return_hidden_states=True,
                        return_resample_mask=True, return_dict=False)
    rec = dict(case=name, resample=resample, prev_w=prev_w, seed_t=11, seed_b=12, seed_in=1,
               ck_t=checksum(sd_t), ck_b=checksum(sd_b),
               branch_samples=[s.clone() for s in samples], noise_pred=out.clone(),
               resample_mask=rmask.clone(), hs_last=hs[-1].clone(),
               hs_sums=torch.stack([h.double().sum() for h in hs]))
    if two_windows:
        inp2 = O.make_inputs(cfg, seed=2)
        lat2 = torch.cat([inp2["latents"], inp2["image_latents"]], dim=2)
        cond2 = torch.cat([inp2["masked_latents"], inp2["mask"]], dim=2)
        samples2 = br(hidden_states=inp2["latents"], encoder_hidden_states=inp2["text"], branch_cond=cond2,
                      timestep=inp2["timestep"], image_rotary_emb=inp2["rope"], return_dict=False)[0]
        kw = dict(prev_hidden_states={i: h for i, h in enumerate(hs)}, prev_clip_weight=prev_w,
                  prev_resample_mask=rmask)
        out2, hs2, rmask2 = tr(hidden_states=lat2, encoder_hidden_states=inp2["text"], timestep=inp2["timestep"],
                               image_rotary_emb=inp2["rope"], branch_block_samples=samples2,
                               attention_kwargs=kw, branch_block_masks=inp2["mask"][:, :, :1],
                               return_hidden_states=True, return_resample_mask=True, return_dict=False)
        rec.update(noise_pred_w2=out2.clone(), resample_mask_w2=rmask2.clone(), hs_last_w2=hs2[-1].clone())
    path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
    torch.save(rec, path)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB")


@torch.no_grad()
def run_block_case():
    """One full-width CogVideoXBlock (D=3072, 48 heads) at a short sequence: pins the block maths at the
    production width.  Stores a strided subsample of the outputs."""
    sys.path.insert(0, REF_SRC)
    from diffusers.models.transformers.cogvideox_transformer_3d import CogVideoXBlock  # type: ignore
    cfg = O.full_config(num_layers=1, sample_height=8, sample_width=8, max_text_seq_length=24)
    sd = O.init_state_dict(cfg, 21)
    blk = CogVideoXBlock(dim=3072, num_attention_heads=48, attention_head_dim=64, time_embed_dim=512,
                         attention_bias=True).eval()
    bsd = {k[len("transformer_blocks.0."):]: v for k, v in sd.items() if k.startswith("transformer_blocks.0.")}
    blk.load_state_dict(bsd, strict=True)
    g = torch.Generator().manual_seed(5)
    Sv = cfg.latent_frames * 4 * 4
    h = torch.randn(1, Sv, 3072, generator=g)
    e = torch.randn(1, 24, 3072, generator=g)
    temb = torch.randn(1, 512, generator=g)
    rope = O.rope_3d(64, ((0, 0), (4, 4)), (4, 4), cfg.latent_frames)
    ho, eo = blk(h, e, temb, image_rotary_emb=rope)
    rec = dict(case="block_full_width", seed=21, seed_in=5, h_out=ho[:, ::7, ::5].clone(), e_out=eo[:, ::3, ::5].clone(),
               h_sum=ho.double().sum(), e_sum=eo.double().sum())
    path = os.path.join(ROOT, "tests", "golden", "block_full_width.pt")
    torch.save(rec, path)
    print("block_full_width ->", path, os.path.getsize(path) // 1024, "KiB")


@torch.no_grad()
def run_variants():
    """Corners of the reference the first two cases do not reach: window 2 with prev_clip_weight == 0.0 (T3D:141-146
    still normalises the previous states, both processors then ignore them: AP:2156, 2247), add_first=True WITH masks
    (T3D:600-609), the wo_text branch (BR:407-412 on a branch built with wo_text=True) and fused QKV projections (T3D:433-456)
    on a model built with the resample processor.  Outputs of the real reference modules."""
    rec = dict(case="tiny_variants", seed_t=11, seed_b=12, seed_in=1)
    for resample in (False, True):
        cfg, cfg_b = O.tiny_config(id_pool_resample_learnable=resample), O.tiny_config(num_layers=1)
        tr, br, sd_t, sd_b = build_reference(cfg, cfg_b, 11, 12)
        inp, inp2 = O.make_inputs(cfg, seed=1), O.make_inputs(cfg, seed=2)

        def step(i, kw=None, add_first=False, model=tr):
            lat = torch.cat([i["latents"], i["image_latents"]], dim=2)
            cond = torch.cat([i["masked_latents"], i["mask"]], dim=2)
            smp = br(hidden_states=i["latents"], encoder_hidden_states=i["text"], branch_cond=cond, timestep=i["timestep"],
                     image_rotary_emb=i["rope"], return_dict=False)[0]
            return model(hidden_states=lat, encoder_hidden_states=i["text"], timestep=i["timestep"], image_rotary_emb=i["rope"],
                         branch_block_samples=smp, attention_kwargs=kw, branch_block_masks=i["mask"][:, :, :1], add_first=add_first,
                         return_hidden_states=True, return_resample_mask=True, return_dict=False)
        out, hs, rmask = step(inp)
        kw = dict(prev_hidden_states={i: h for i, h in enumerate(hs)}, prev_clip_weight=0.0, prev_resample_mask=rmask)
        tag = "resample" if resample else "plain"
        rec[f"w2_prev0_{tag}"] = step(inp2, kw)[0].clone()
        rec[f"add_first_masked_{tag}"] = step(inp, add_first=True)[0].clone()
        if resample:
            tr.fuse_qkv_projections()
            kw = dict(prev_hidden_states={i: h for i, h in enumerate(hs)}, prev_clip_weight=0.5, prev_resample_mask=rmask)
            rec["fused_qkv_w2"] = step(inp2, kw)[0].clone()
            tr.unfuse_qkv_projections()
    # wo_text branch
    _, BR = import_reference()
    cfg_b = O.tiny_config(num_layers=2)
    sd_b = O.init_state_dict(cfg_b, 13, branch=True)
    brw = BR(**ref_kwargs(cfg_b), wo_text=True).eval()
    brw.load_state_dict(sd_b, strict=True)
    inp = O.make_inputs(cfg_b, seed=1)
    cond = torch.cat([inp["masked_latents"], inp["mask"]], dim=2)
    smp = brw(hidden_states=inp["latents"], encoder_hidden_states=inp["text"], branch_cond=cond, timestep=inp["timestep"],
              image_rotary_emb=inp["rope"], conditioning_scale=0.7, wo_text=True, return_dict=False)[0]
    rec["wo_text_samples"] = [s.clone() for s in smp]
    rec["wo_text_seed_b"] = 13
    path = os.path.join(ROOT, "tests", "golden", "tiny_variants.pt")
    torch.save(rec, path)
    print("tiny_variants ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    run_case("tiny_step", resample=False, two_windows=True, prev_w=0.5)
    run_case("tiny_step_resample", resample=True, two_windows=True, prev_w=0.5)
    run_block_case()
    run_variants()
