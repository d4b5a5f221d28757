"""CPU checks of the drop-in against the REAL reference classes (only where /root/reference is mounted, i.e. in the build
container): install() on the diffusers fork, shape / processor introspection and weight packing from the live reference
modules (plain, ID-resample, from_transformer, fused QKV, LoRA-wrapped) must equal what the mirror classes give for the same
state-dict; the patched forward refuses to run on CPU (no fallback) and uninstall() restores the reference."""
import os
import sys

import pytest
import torch

REF = "/root/reference/diffusers/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from diffusers import CogVideoXTransformer3DModel, CogvideoXBranchModel  # type: ignore
    return CogVideoXTransformer3DModel, CogvideoXBranchModel


def _packed_equal(a, b):
    import dataclasses
    for f in dataclasses.fields(a):
        if f.name in ("workspace", "dims", "blocks"):
            continue
        x, y = getattr(a, f.name), getattr(b, f.name)
        if torch.is_tensor(x):
            assert torch.equal(x, y), f.name
        elif isinstance(x, list):
            assert len(x) == len(y) and all(torch.equal(p, q) for p, q in zip(x, y)), f.name
        else:
            assert x == y, f.name
    assert len(a.blocks) == len(b.blocks)
    for ba, bb in zip(a.blocks, b.blocks):
        for f in dataclasses.fields(ba):
            assert torch.equal(getattr(ba, f.name), getattr(bb, f.name)), f.name


@pytest.mark.parametrize("resample", [False, True])
def test_packing_from_live_reference_modules_equals_mirror(ref, resample):
    import videopainter_b200 as vp
    from oracle import cogvideox_oracle as O
    from videopainter_b200.models import dims_from_module, packed_for
    RT, RB = ref
    cfg = O.tiny_config(id_pool_resample_learnable=resample)
    kw = cfg.to_kwargs(); kw.pop("norm_eps")
    sd = O.init_state_dict(cfg, 3)
    rt = RT(**kw).eval(); rt.load_state_dict(sd, strict=True)
    mt = vp.CogVideoXTransformer3DModel(**kw); mt.load_state_dict(sd, strict=True)
    assert dims_from_module(rt, False) == dims_from_module(mt, False)
    assert dims_from_module(rt, False).resample == resample
    _packed_equal(packed_for(rt, False, "cpu"), packed_for(mt, False, "cpu"))
    # the branch built the way infer/inpaint.py:327-333 does without a checkpoint: from_transformer (BR:255-293)
    rb = RB.from_transformer(rt, num_layers=1, attention_head_dim=64, num_attention_heads=2, load_weights_from_transformer=True)
    db = dims_from_module(rb, True)
    assert (db.num_layers, db.patch_in_channels, db.is_branch, db.resample, db.wo_text) == (1, 33, True, False, False)
    kwb = O.tiny_config(num_layers=1).to_kwargs(); kwb.pop("norm_eps")
    mb = vp.CogvideoXBranchModel(**kwb); mb.load_state_dict(rb.state_dict(), strict=True)
    _packed_equal(packed_for(rb, True, "cpu"), packed_for(mb, True, "cpu"))
    # steady state: the second call is a cache hit on the sentinel key (same object back)
    assert packed_for(rt, False, "cpu") is packed_for(rt, False, "cpu")
    # load_state_dict / .to() / optimiser steps touch every parameter, sentinels included -> repacked
    sd2 = {k: (v + 1.0 if k == "proj_out.weight" else v) for k, v in sd.items()}
    rt.load_state_dict(sd2, strict=True)
    assert packed_for(rt, False, "cpu").proj_w[0, 0] != packed_for(mt, False, "cpu").proj_w[0, 0]
    # an in-place edit of ONE parameter is the documented blind spot of the sentinel check: invalidate() covers it
    before = packed_for(rt, False, "cpu")
    with torch.no_grad():
        rt.transformer_blocks[0].ff.net[2].weight.mul_(2.0)
    from videopainter_b200.models import invalidate
    invalidate(rt)
    assert not torch.equal(packed_for(rt, False, "cpu").blocks[0].ff2_w, before.blocks[0].ff2_w)


def test_fused_qkv_projections_of_the_reference(ref):
    import videopainter_b200 as vp
    from oracle import cogvideox_oracle as O
    from videopainter_b200.models import dims_from_module, packed_for
    RT, _ = ref
    cfg = O.tiny_config(id_pool_resample_learnable=True)
    kw = cfg.to_kwargs(); kw.pop("norm_eps")
    sd = O.init_state_dict(cfg, 4)
    rt = RT(**kw).eval(); rt.load_state_dict(sd, strict=True)
    plain = packed_for(rt, False, "cpu")
    assert plain.dims.resample and not plain.dims.fused_qkv
    rt.fuse_qkv_projections()                              # T3D:433-456: to_qkv + FusedCogVideoXAttnProcessor2_0
    fused = packed_for(rt, False, "cpu")
    assert fused is not plain and fused.dims.fused_qkv and not fused.dims.resample      # the fused processor has no resample path
    for a, b in zip(plain.blocks, fused.blocks):
        assert torch.equal(a.qkv_w, b.qkv_w) and torch.equal(a.qkv_b, b.qkv_b)
    rt.unfuse_qkv_projections()
    assert dims_from_module(rt, False).resample
    # the mirror offers the same switch
    mt = vp.CogVideoXTransformer3DModel(**kw); mt.load_state_dict(sd, strict=True)
    mt.fuse_qkv_projections()
    assert dims_from_module(mt, False).fused_qkv
    for a, b in zip(plain.blocks, packed_for(mt, False, "cpu").blocks):
        assert torch.equal(a.qkv_w, b.qkv_w)


def test_wo_text_branch_of_the_reference(ref):
    from oracle import cogvideox_oracle as O
    from videopainter_b200.models import dims_from_module
    import videopainter_b200 as vp
    _, RB = ref
    kwb = O.tiny_config(num_layers=1).to_kwargs(); kwb.pop("norm_eps")
    rb = RB(**kwb, wo_text=True)
    assert dims_from_module(rb, True).wo_text
    assert dims_from_module(vp.CogvideoXBranchModel(**kwb, wo_text=True), True).wo_text


def test_lora_wrapped_reference_module_packs_like_the_oracle_merge(ref):
    """PEFT-style wrapping of the live reference module (test stand-in for peft, tests/_fake_peft.py): active / inactive
    adapters, alpha / r scaling and attention_kwargs['scale'] (T3D:490-498) all reach the packed weights."""
    from _fake_peft import inject
    from oracle import cogvideox_oracle as O
    from videopainter_b200.models import packed_for
    RT, _ = ref
    cfg = O.tiny_config()
    kw = cfg.to_kwargs(); kw.pop("norm_eps")
    sd = O.init_state_dict(cfg, 6)
    rt = RT(**kw).eval(); rt.load_state_dict(sd, strict=True)
    lora = inject(rt, r=4, lora_alpha=2, seed=1, adapters=("id", "unused"), inactive=("unused",))    # scaling 0.5
    for scale in (1.0, 0.25):
        pm = packed_for(rt, False, "cpu", lora_scale=scale)
        merged = O.lora_merge(sd, {k: v["id"] for k, v in lora.items()}, scale=0.5 * scale)
        for i, blk in enumerate(pm.blocks):
            p = f"transformer_blocks.{i}.attn1."
            want = torch.cat([merged[p + n + ".weight"] for n in ("to_q", "to_k", "to_v")]).bfloat16()
            assert torch.equal(blk.qkv_w, want), (scale, i)
            assert torch.equal(blk.out_w, merged[p + "to_out.0.weight"].bfloat16())
    for m in rt.modules():                                  # disable_adapters: base weights only
        if hasattr(m, "lora_A"):
            m.disable_adapters = True
    pm = packed_for(rt, False, "cpu")
    assert torch.equal(pm.blocks[0].out_w, sd["transformer_blocks.0.attn1.to_out.0.weight"].bfloat16())


def test_install_patches_the_reference_classes_and_refuses_cpu(ref):
    import videopainter_b200 as vp
    from oracle import cogvideox_oracle as O
    RT, RB = ref
    cfg = O.tiny_config()
    kw = cfg.to_kwargs(); kw.pop("norm_eps")
    rt = RT(**kw).eval()
    inp = O.make_inputs(cfg, 1)
    lat = torch.cat([inp["latents"], inp["image_latents"]], dim=2)
    with torch.no_grad():
        want = rt(lat, inp["text"], inp["timestep"], image_rotary_emb=inp["rope"], return_dict=False)[0]
    vp.install()
    try:
        assert RT.forward is vp.models.transformer_forward and RB.forward is vp.models.branch_forward
        with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
            rt(lat, inp["text"], inp["timestep"], image_rotary_emb=inp["rope"], return_dict=False)
    finally:
        vp.uninstall()
    with torch.no_grad():
        again = rt(lat, inp["text"], inp["timestep"], image_rotary_emb=inp["rope"], return_dict=False)[0]
    assert torch.equal(want, again)
