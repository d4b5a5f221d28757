"""Host-side logic of the multi-GPU layout on CPU: plan / shard arithmetic, and a world_size-2 gloo run of the two
collectives of the data path together with the send / receive layouts the kernels use (restated in torch here, as the
checker: include/vp_b200.h vp_gemm_qkv heads_per_dest, vp_a2a_unpack_heads, vp_gemm_gate_residual a_k_chunk)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from videopainter_b200 import parallel


def test_plan_layout():
    p = parallel.make_plan(1, 0)
    assert (p.cfg_groups, p.sp, p.describe()) == (1, 1, "single GPU")
    assert p.batch_slice(2) == slice(0, 2)
    p = parallel.make_plan(2, 1)
    assert (p.cfg_groups, p.sp, p.cfg_index, p.sp_rank) == (2, 1, 1, 0)
    assert p.batch_slice(2) == slice(1, 2)
    p = parallel.make_plan(8, 6)
    assert (p.cfg_groups, p.sp, p.cfg_index, p.sp_rank, p.sp_ranks()) == (2, 4, 1, 2, (4, 5, 6, 7))
    assert p.describe() == "cfg2 x ulysses4"
    with pytest.raises(ValueError):
        parallel.make_plan(3, 0)
    with pytest.raises(ValueError):
        parallel.make_plan(4, 4)
    with pytest.raises(ValueError):
        parallel.make_plan(2, 0).local_batch(1)


@pytest.mark.parametrize("sp", [1, 2, 4, 8, 16])
def test_shard_rows_of_the_production_sequence(sp):
    S, St, H = 17776, 226, 48
    shards = [parallel.Shard(sp, r, S, St, H) for r in range(sp)]
    assert sum(s.rows for s in shards) == S
    assert sum(s.text_rows for s in shards) == St and sum(s.video_rows for s in shards) == S - St
    pos = 0
    for s in shards:
        assert s.row0 == pos and s.heads_local * sp == H
        assert s.video0 == max(pos - St, 0)
        assert (s.text_rows > 0) == (pos < St)
        pos += s.rows
    assert parallel.send_block_shape(shards[0], 3) == (sp, 3, H // sp, S // sp, 64)
    assert parallel.qkv_dest_stride(shards[0], 3) == 3 * (H // sp) * (S // sp) * 64


def test_shard_rejects_indivisible_shapes():
    with pytest.raises(ValueError):
        parallel.Shard(3, 0, 17776, 226, 48)
    with pytest.raises(ValueError):
        parallel.Shard(4, 0, 224, 16, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, S, St, H):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # world 2 is one CFG group per rank (sp = 1); exercise the sequence-parallel collectives with a plan whose
        # sequence group is the whole world
        plan = parallel.Plan(world=world, rank=rank, cfg_groups=1, sp=world)
        rt = parallel.Runtime(plan, sp_group=None, world_group=None)
        sh = rt.shard(S, St, H)
        P, Hl, R = sh.sp, sh.heads_local, sh.rows
        g = torch.Generator().manual_seed(0)
        full = torch.randn(3, H, S, 64, generator=g)                     # q, k, v of the whole sequence, all heads
        # what vp_gemm_qkv writes on this rank: [dest][slot][head within dest][owned row][64]
        send = torch.empty(parallel.send_block_shape(sh, 3))
        for d in range(P):
            send[d] = full[:, d * Hl:(d + 1) * Hl, sh.row0:sh.row0 + R]
        recv = torch.empty_like(send)
        rt.all_to_all(recv, send)
        # what vp_a2a_unpack_heads produces: per slot [head_local][peer * R + row][64]
        got = recv.permute(1, 2, 0, 3, 4).reshape(3, Hl, P * R, 64)
        assert torch.equal(got, full[:, rank * Hl:(rank + 1) * Hl])
        # attention output of the local heads over the whole sequence [S, Hl * 64] -> token shards
        ao_full = torch.randn(S, H * 64, generator=g)
        ao_local = ao_full[:, rank * Hl * 64:(rank + 1) * Hl * 64].contiguous()
        ao_recv = torch.empty(P, R, Hl * 64)
        rt.all_to_all(ao_recv, ao_local)
        # K-chunked A operand of the out-projection: column k of row r lives in chunk k // (Hl * 64)
        a = ao_recv.permute(1, 0, 2).reshape(R, H * 64)
        assert torch.equal(a, ao_full[sh.row0:sh.row0 + R])
        # final gather: every rank's [R, n] slot -> the joint [S, n] on every rank
        slot = ao_full[sh.row0:sh.row0 + R, :64].contiguous()
        joint = torch.empty(S, 64)
        rt.all_gather(joint, slot)
        assert torch.equal(joint, ao_full[:, :64])
        # parallel.init builds the CFG x Ulysses groups from the live process group
        rt2 = parallel.init()
        assert rt2.plan == parallel.make_plan(world, rank) and parallel.current() is rt2
        parallel.shutdown()
        assert parallel.current() is None
    finally:
        dist.destroy_process_group()


def test_gloo_world2_collectives_and_layouts():
    mp.spawn(_worker, args=(2, _free_port(), 224, 16, 4), nprocs=2, join=True)
