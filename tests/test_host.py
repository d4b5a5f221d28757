"""CPU-side checks: the C-ABI library loads and exports every symbol of include/vp_b200.h, the host mirror keeps the
reference's signatures / parameter names, weight packing (fused QKV, LoRA merge, conv padding) and that the product path
fails loudly without CUDA (no fallback)."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/diffusers/src"


def test_capi_exports_every_declared_symbol():
    from videopainter_b200.build import build, LIB
    build()
    hdr = open(os.path.join(ROOT, "include", "vp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 14
    L = ctypes.CDLL(LIB)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in vp_b200.h but not exported"
    L.vp_version.restype = ctypes.c_int
    assert L.vp_version() >= 100
    from videopainter_b200._lib import SIGNATURES
    assert set(SIGNATURES) | {"vp_version", "vp_last_error", "vp_last_cuda_error"} == declared


def test_ctypes_signatures_match_header_arity():
    from videopainter_b200._lib import SIGNATURES
    hdr = open(os.path.join(ROOT, "include", "vp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for name, args in SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", hdr, flags=re.S)
        assert m, name
        assert len([a for a in m.group(1).split(",") if a.strip()]) == len(args), name


def test_no_cpu_fallback():
    import videopainter_b200 as vp
    from oracle import cogvideox_oracle as O
    cfg = O.tiny_config()
    kw = cfg.to_kwargs(); kw.pop("norm_eps")
    tr = vp.CogVideoXTransformer3DModel(**kw)
    inp = O.make_inputs(cfg, 1)
    lat = torch.cat([inp["latents"], inp["image_latents"]], dim=2)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        tr(lat, inp["text"], inp["timestep"], image_rotary_emb=inp["rope"], return_dict=False)


def test_mirror_state_dict_names_match_oracle_and_reference_names():
    import videopainter_b200 as vp
    from oracle import cogvideox_oracle as O
    cfg = O.tiny_config(id_pool_resample_learnable=True)
    kw = cfg.to_kwargs(); kw.pop("norm_eps")
    tr = vp.CogVideoXTransformer3DModel(**kw)
    tr.load_state_dict(O.init_state_dict(cfg, 1), strict=True)          # names pinned to the reference by make_golden.py
    kwb = O.tiny_config(num_layers=1).to_kwargs(); kwb.pop("norm_eps")
    br = vp.CogvideoXBranchModel(**kwb)
    br.load_state_dict(O.init_state_dict(O.tiny_config(num_layers=1), 2, branch=True), strict=True)
    from videopainter_b200.models import dims_from_module
    d = dims_from_module(tr, False)
    assert (d.heads, d.head_dim, d.num_layers, d.resample, d.patch_in_channels, d.max_text) == (2, 64, 2, True, 32, 16)
    db = dims_from_module(br, True)
    assert db.patch_in_channels == 33 and db.is_branch and not db.resample


def test_pack_fuses_qkv_pads_conv_and_merges_lora():
    from oracle import cogvideox_oracle as O
    from videopainter_b200.engine import Dims, pack_state_dict
    cfg = O.tiny_config(num_layers=1)
    sd = O.init_state_dict(cfg, 5, branch=True)
    dims = Dims(heads=2, head_dim=64, time_dim=512, text_dim=64, patch_in_channels=33, out_channels=16, patch=2, max_text=16,
                num_layers=1, is_branch=True)
    pm = pack_state_dict(sd, dims, "cpu")
    b0 = pm.blocks[0]
    assert b0.qkv_w.shape == (384, 128) and pm.kpad == 192 and pm.patch_w.shape == (128, 192)
    assert torch.equal(b0.qkv_w[128:256], sd["transformer_blocks.0.attn1.to_k.weight"].bfloat16())
    assert torch.equal(pm.patch_w[:, :132], sd["patch_embed.proj.weight"].reshape(128, 132).bfloat16())
    assert (pm.patch_w[:, 132:] == 0).all() and len(pm.branch_w) == 1
    # PEFT-style names: base_layer + lora_A/lora_B -> merged W + B @ A (scale 1.0 at inference)
    p = "transformer_blocks.0.attn1.to_q"
    A, Bm = torch.randn(8, 128) * 0.1, torch.randn(128, 8) * 0.1
    sd2 = dict(sd)
    sd2[p + ".base_layer.weight"] = sd2.pop(p + ".weight")
    sd2[p + ".base_layer.bias"] = sd2.pop(p + ".bias")
    sd2[p + ".lora_A.default.weight"] = A
    sd2[p + ".lora_B.default.weight"] = Bm
    pm2 = pack_state_dict(sd2, dims, "cpu")
    want = O.lora_merge(sd, {p: (A, Bm)})[p + ".weight"].bfloat16()
    assert torch.equal(pm2.blocks[0].qkv_w[:128], want)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")
def test_forward_signatures_equal_reference():
    import sys
    sys.path.insert(0, REF)
    from diffusers import CogVideoXTransformer3DModel as RT, CogvideoXBranchModel as RB  # type: ignore
    import videopainter_b200 as vp

    def names(f):
        return [(n, p.default) for n, p in inspect.signature(f).parameters.items()]
    assert names(vp.CogVideoXTransformer3DModel.forward) == names(RT.forward)
    assert names(vp.CogvideoXBranchModel.forward) == names(RB.forward)
    ours = set(inspect.signature(vp.CogVideoXTransformer3DModel.__init__).parameters) - {"device", "dtype"}
    assert ours == set(inspect.signature(RT.__init__).parameters)
    # install() swaps the class-level forward and uninstall() restores it
    orig = RT.forward
    vp.install()
    assert RT.forward is not orig and RB.forward.__name__ == "branch_forward"
    vp.uninstall()
    assert RT.forward is orig


def test_torch_ops_registered_with_the_c_prototypes():
    """csrc/torch_ops.cpp: every compute entry point of the header is a torch.ops.vp_b200 op whose schema is the C prototype
    minus the trailing stream (pointer -> Tensor?, pointer array -> int[], integer -> int, float -> float)."""
    from videopainter_b200.build import build, TORCH_LIB
    from videopainter_b200._lib import SIGNATURES
    build()
    torch.ops.load_library(TORCH_LIB)
    assert torch.ops.vp_b200.version() >= 100
    hdr = open(os.path.join(ROOT, "include", "vp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    host_only = {"vp_peer_alloc", "vp_peer_open", "vp_peer_close", "vp_peer_free", "vp_peer_set_timeout_ms"}
    n_ops = 0
    for name in SIGNATURES:
        if name in host_only:
            continue
        proto = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", hdr, flags=re.S).group(1)
        params = [a.strip() for a in proto.split(",") if a.strip()]
        assert params[-1].endswith("stream"), name
        schema = getattr(torch.ops.vp_b200, name[3:]).default._schema
        assert len(schema.arguments) == len(params) - 1, name
        for a, c in zip(schema.arguments, params[:-1]):
            want = "List[int]" if "* const*" in c else ("Optional[Tensor]" if "*" in c else ("float" if c.startswith("float") else "int"))
            assert str(a.type) == want, (name, c, str(a.type))
        n_ops += 1
    assert n_ops == 18
