"""CUDA-graph replay of the two forwards (videopainter_b200/graphs.py) must be BIT-IDENTICAL to the eager
launches: the graph contains the same kernels with the same arguments; only the host side differs.  Checked on the tiny
configuration over several steps with changing latents / timesteps (first call eager, second captured, then replays), for both
attention processors, and on a second window that carries the first window's hidden states by address."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


def _models(resample):
    import videopainter_b200 as vp
    from oracle import cogvideox_oracle as O
    cfg = O.tiny_config(id_pool_resample_learnable=resample)
    cfg_b = O.tiny_config(num_layers=1)
    kw = cfg.to_kwargs(); kw.pop("norm_eps")
    tr = vp.CogVideoXTransformer3DModel(**kw, device="cuda", dtype=BF16)
    tr.load_state_dict({k: v.to(BF16) for k, v in O.init_state_dict(cfg, 41).items()}, strict=True)
    kwb = cfg_b.to_kwargs(); kwb.pop("norm_eps")
    br = vp.CogvideoXBranchModel(**kwb, device="cuda", dtype=BF16)
    br.load_state_dict({k: v.to(BF16) for k, v in O.init_state_dict(cfg_b, 42, branch=True).items()}, strict=True)
    return cfg, tr, br


def _step(tr, br, inp, rope, attention_kwargs=None, keep_hs=False):
    lat_in = torch.cat([inp["latents"], inp["image_latents"]], dim=2).to(BF16)
    cond = torch.cat([inp["masked_latents"], inp["mask"]], dim=2).to(BF16)
    text = inp["text"].to(BF16)
    samples = br(hidden_states=inp["latents"].to(BF16), encoder_hidden_states=text, branch_cond=cond, timestep=inp["timestep"],
                 image_rotary_emb=rope, return_dict=False)[0]
    out, hs, rmask = tr(hidden_states=lat_in, encoder_hidden_states=text, timestep=inp["timestep"], image_rotary_emb=rope,
                        branch_block_samples=samples, attention_kwargs=attention_kwargs,
                        branch_block_masks=inp["mask"][:, :, :1].to(BF16), return_hidden_states=True,
                        return_resample_mask=True, return_dict=False)
    res = [s.clone() for s in samples] + [out.clone(), hs[-1].clone(), rmask.clone()]
    return res, (hs if keep_hs else None), rmask


@pytest.mark.parametrize("resample", [False, True])
def test_graph_replay_is_bit_identical_to_eager(resample):
    import videopainter_b200 as vp
    from videopainter_b200 import graphs, ops
    from oracle import cogvideox_oracle as O
    cfg, tr, br = _models(resample)
    inputs = [O.make_inputs(cfg, 100 + i, device="cuda") for i in range(4)]
    for i, inp in enumerate(inputs):
        inp["timestep"] = torch.full_like(inp["timestep"], 999 - 250 * i)
    rope = inputs[0]["rope"]
    graphs.enable_graphs(False)
    eager = [_step(tr, br, inp, rope)[0] for inp in inputs]
    n0 = ops.launch_count
    _step(tr, br, inputs[0], rope)
    per_step = ops.launch_count - n0
    try:
        graphs.enable_graphs(True)
        n0 = ops.launch_count
        graphed = [_step(tr, br, inp, rope)[0] for inp in inputs]            # eager, capture + replay, replay, replay
        assert ops.launch_count - n0 == 4 * per_step                            # the launches inside a replay are counted
        for a, b in zip(eager, graphed):
            for x, y in zip(a, b):
                assert torch.equal(x, y)
        pm_t, pm_b = tr.__dict__["_vp_packed"]["pm"], br.__dict__["_vp_packed"]["pm"]
        assert graphs.stats(pm_t) == {"graphs": 1, "replays": 3, "captures": 1}
        assert graphs.stats(pm_b) == {"graphs": 1, "replays": 3, "captures": 1}

        # second window: the previous window's states are captured by address; their CONTENT may change between replays
        graphs.enable_graphs(False)
        _, hs_e, rm_e = _step(tr, br, inputs[0], rope, keep_hs=True)
        prev = {i: h.clone() for i, h in enumerate(hs_e)}
        kw = dict(prev_hidden_states=prev, prev_clip_weight=0.5, prev_resample_mask=rm_e)
        eager2 = [_step(tr, br, inp, rope, attention_kwargs=kw)[0] for inp in inputs[1:]]
        for h in prev.values():
            h.mul_(0.5)
        eager2b = _step(tr, br, inputs[3], rope, attention_kwargs=kw)[0]
        for h in prev.values():
            h.mul_(2.0)
        graphs.enable_graphs(True)
        graphed2 = [_step(tr, br, inp, rope, attention_kwargs=kw)[0] for inp in inputs[1:]]
        for h in prev.values():
            h.mul_(0.5)
        graphed2b = _step(tr, br, inputs[3], rope, attention_kwargs=kw)[0]
        for a, b in zip(eager2 + [eager2b], graphed2 + [graphed2b]):
            for x, y in zip(a, b):
                assert torch.equal(x, y)
        assert graphs.stats(pm_t)["graphs"] == 2

        # a new shape releases the workspace and with it every captured graph
        graphs.clear(pm_t)
        assert graphs.stats(pm_t)["graphs"] == 0
    finally:
        graphs.enable_graphs(False)


def test_graph_keeps_the_hidden_state_list_until_the_signature_runs_again():
    from videopainter_b200 import graphs
    from oracle import cogvideox_oracle as O
    cfg, tr, br = _models(False)
    a, b = O.make_inputs(cfg, 7, device="cuda"), O.make_inputs(cfg, 8, device="cuda")
    rope = a["rope"]
    try:
        graphs.enable_graphs(True)
        _step(tr, br, a, rope)
        _, hs1, _ = _step(tr, br, a, rope, keep_hs=True)                       # captured: hs1 are views of the graph's arena
        snap = [h.clone() for h in hs1]
        # a call with ANOTHER signature (previous-window states present) must not touch them
        kw = dict(prev_hidden_states={i: h for i, h in enumerate(hs1)}, prev_clip_weight=0.5)
        _step(tr, br, b, rope, attention_kwargs=kw)
        _step(tr, br, b, rope, attention_kwargs=kw)
        assert all(torch.equal(x, y) for x, y in zip(snap, hs1))
    finally:
        graphs.enable_graphs(False)
