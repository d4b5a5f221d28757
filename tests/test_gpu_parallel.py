"""Sequence-parallel (Ulysses) + CFG-split data path on ONE GPU: virtual ranks (tests/_fabric.py) run the sharded forwards
through the real kernels and must reproduce the single-GPU result.  Token-wise kernels compute every row independently and
attention computes every head independently, so the comparison is exact (bit-for-bit) for the noise prediction."""
import pytest
import torch

from _fabric import run_virtual_ranks

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


def _build(cfg, cfg_b, sd_t, sd_b):
    import videopainter_b200 as vp
    kw = cfg.to_kwargs(); kw.pop("norm_eps")
    tr = vp.CogVideoXTransformer3DModel(**kw, device="cuda", dtype=BF16)
    tr.load_state_dict({k: v.to(BF16) for k, v in sd_t.items()}, strict=True)
    kwb = cfg_b.to_kwargs(); kwb.pop("norm_eps")
    br = vp.CogvideoXBranchModel(**kwb, device="cuda", dtype=BF16)
    br.load_state_dict({k: v.to(BF16) for k, v in sd_b.items()}, strict=True)
    return tr, br


def _step(tr, br, inp, attention_kwargs=None):
    lat_in = torch.cat([inp["latents"], inp["image_latents"]], dim=2).to(BF16)
    cond = torch.cat([inp["masked_latents"], inp["mask"]], dim=2).to(BF16)
    text = inp["text"].to(BF16)
    samples = br(hidden_states=inp["latents"].to(BF16), encoder_hidden_states=text, branch_cond=cond, timestep=inp["timestep"],
                 image_rotary_emb=inp["rope"], return_dict=False)[0]
    out, hs, rmask = tr(hidden_states=lat_in, encoder_hidden_states=text, timestep=inp["timestep"],
                        image_rotary_emb=inp["rope"], branch_block_samples=samples, attention_kwargs=attention_kwargs,
                        branch_block_masks=inp["mask"][:, :, :1].to(BF16), return_hidden_states=True,
                        return_resample_mask=True, return_dict=False)
    torch.cuda.synchronize()
    return samples, out, hs, rmask


@pytest.mark.parametrize("p2p", [True, False], ids=["peer-memory", "nccl-layout"])
@pytest.mark.parametrize("world,heads,resample", [(2, 2, False), (4, 2, False), (4, 2, True), (8, 4, False)])
def test_virtual_ranks_match_single_gpu(world, heads, resample, p2p):
    from oracle import cogvideox_oracle as O
    cfg = O.tiny_config(num_attention_heads=heads, id_pool_resample_learnable=resample)
    cfg_b = O.tiny_config(num_attention_heads=heads, num_layers=1)
    sd_t, sd_b = O.init_state_dict(cfg, 51), O.init_state_dict(cfg_b, 52, branch=True)
    inp = O.make_inputs(cfg, 5, device="cuda")
    inp2 = O.make_inputs(cfg, 6, device="cuda")
    tr, br = _build(cfg, cfg_b, sd_t, sd_b)
    s1, out1, hs1, rm1 = _step(tr, br, inp)
    kw1 = dict(prev_hidden_states={i: h for i, h in enumerate(hs1)}, prev_clip_weight=0.5, prev_resample_mask=rm1)
    _, out1b, _, _ = _step(tr, br, inp2, attention_kwargs=kw1)

    def rank_fn(r, rt):
        trr, brr = _build(cfg, cfg_b, sd_t, sd_b)                 # one model (and workspace) per virtual rank
        s, out, hs, rm = _step(trr, brr, inp)
        kw = dict(prev_hidden_states={i: h for i, h in enumerate(hs)}, prev_clip_weight=0.5, prev_resample_mask=rm)
        _, outb, _, _ = _step(trr, brr, inp2, attention_kwargs=kw)   # second window: sharded prev states stay on the rank
        return s, out, hs, rm, outb

    res = run_virtual_ranks(world, rank_fn, p2p=p2p)
    sp = world // 2
    for r, (s, out, hs, rm, outb) in enumerate(res):
        assert torch.equal(rm, rm1)
        assert torch.equal(out, out1), f"rank {r}: noise prediction differs from the single-GPU result"
        assert torch.equal(outb, out1b), f"rank {r}: second-window noise prediction differs"
    # the sharded hidden states / branch samples of one CFG half, concatenated over its ranks, are the single-GPU tensors
    for g in range(2):
        last = torch.cat([res[g * sp + k][2][-1] for k in range(sp)], dim=1)
        assert torch.equal(last, hs1[-1][g:g + 1])
        bs = torch.cat([res[g * sp + k][0][0] for k in range(sp)], dim=1)
        assert torch.equal(bs, s1[0][g:g + 1])


def test_sharded_inputs_are_validated():
    from oracle import cogvideox_oracle as O
    cfg = O.tiny_config()
    sd_t = O.init_state_dict(cfg, 53)
    inp = O.make_inputs(cfg, 7, device="cuda")

    def rank_fn(r, rt):
        import videopainter_b200 as vp
        kw = cfg.to_kwargs(); kw.pop("norm_eps")
        tr = vp.CogVideoXTransformer3DModel(**kw, device="cuda", dtype=BF16)
        tr.load_state_dict({k: v.to(BF16) for k, v in sd_t.items()}, strict=True)
        lat_in = torch.cat([inp["latents"], inp["image_latents"]], dim=2).to(BF16)
        with pytest.raises(ValueError):
            tr(lat_in[:1], inp["text"][:1].to(BF16), inp["timestep"][:1], image_rotary_emb=inp["rope"], return_dict=False)
        return True

    assert all(run_virtual_ranks(2, rank_fn))


def test_peer_buffer_is_zeroed_device_memory_viewed_without_copy():
    from videopainter_b200 import ops
    buf = ops.PeerBuffer(1 << 20, torch.device("cuda", 0))
    t = buf.tensor
    assert t.is_cuda and t.dtype == torch.uint8 and t.numel() == 1 << 20 and t.data_ptr() == buf.ptr
    assert int(t.sum()) == 0 and len(buf.handle) == 64
    v = t.view(BF16).view(-1, 64)
    v.fill_(1.0)
    assert float(t.view(BF16).float().sum()) == (1 << 19)
    # a single-rank peer barrier returns immediately (its own flag)
    flags = ops.PeerBuffer(64, torch.device("cuda", 0))
    ops.peer_barrier([flags.ptr], 0, 1)
    ops.peer_barrier([flags.ptr], 0, 2)
    torch.cuda.synchronize()
    assert int(flags.tensor.view(torch.int32)[0]) == 2
    # epoch == 0: the kernel counts its own calls (word 9), which is what makes the launch replayable from a CUDA graph
    auto = ops.PeerBuffer(64, torch.device("cuda", 0))
    for _ in range(3):
        ops.peer_barrier([auto.ptr], 0, 0)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ops.peer_barrier([auto.ptr], 0, 0)
    g.replay(); g.replay()
    torch.cuda.synchronize()
    words = auto.tensor.view(torch.int32)
    assert int(words[0]) == 5 and int(words[9]) == 5 and int(words[8]) == 0      # 3 eager + 2 replays (capture does not run)
    buf.free(); flags.free(); auto.free()


def test_two_processes_on_one_gpu_exchange_through_ipc_and_the_device_barrier():
    """The pieces the virtual-rank tests replace — CUDA-IPC mapping of a peer's buffer and the device-side barrier — run for
    real between two processes that share this GPU (tests/ipc_pair_check.py), eagerly and from a CUDA graph."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
        procs.append(subprocess.Popen([sys.executable, os.path.join(root, "tests", "ipc_pair_check.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {rank}: ok" in out, out[-3000:]
