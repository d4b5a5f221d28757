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


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    run_case("tiny_step", resample=False, two_windows=True, prev_w=0.5)
    run_case("tiny_step_resample", resample=True, two_windows=True, prev_w=0.5)
    run_block_case()
