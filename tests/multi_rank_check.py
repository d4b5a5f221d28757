"""Run under torchrun on N GPUs (N = 2, 4, 8): one denoise step of a small model through the CFG-split / Ulysses path over
NCCL, compared on every rank with the same step computed by that rank alone.  Exit code 0 = identical.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_rank_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
BF16 = torch.bfloat16


def main():
    import videopainter_b200 as vp
    from videopainter_b200 import parallel
    from oracle import cogvideox_oracle as O
    from test_gpu_parallel import _build, _step
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = dist.get_world_size(), dist.get_rank()
    ok = True
    for resample in (False, True):
        cfg = O.tiny_config(num_attention_heads=4, id_pool_resample_learnable=resample)
        cfg_b = O.tiny_config(num_attention_heads=4, num_layers=1)
        sd_t, sd_b = O.init_state_dict(cfg, 61), O.init_state_dict(cfg_b, 62, branch=True)
        inp, inp2 = O.make_inputs(cfg, 8, device="cuda"), O.make_inputs(cfg, 9, device="cuda")
        tr, br = _build(cfg, cfg_b, sd_t, sd_b)
        rt = parallel.init()
        s, out, hs, rm = _step(tr, br, inp)
        kw = dict(prev_hidden_states={i: h for i, h in enumerate(hs)}, prev_clip_weight=0.5, prev_resample_mask=rm)
        _, outb, _, _ = _step(tr, br, inp2, attention_kwargs=kw)
        parallel.shutdown()
        tr1, br1 = _build(cfg, cfg_b, sd_t, sd_b)
        s1, out1, hs1, rm1 = _step(tr1, br1, inp)
        kw1 = dict(prev_hidden_states={i: h for i, h in enumerate(hs1)}, prev_clip_weight=0.5, prev_resample_mask=rm1)
        _, out1b, _, _ = _step(tr1, br1, inp2, attention_kwargs=kw1)
        sh = rt.shard(224, 16, 4)
        g = rt.plan.cfg_index
        same = (torch.equal(out, out1) and torch.equal(outb, out1b) and torch.equal(rm, rm1)
                and torch.equal(hs[-1], hs1[-1][g:g + 1, sh.row0:sh.row0 + sh.rows]))
        print(f"rank {rank}/{world} ({rt.plan.describe()}) resample={resample}: "
              f"{'identical to single GPU' if same else 'MISMATCH'}; max |d| = {(out.float() - out1.float()).abs().max().item():.3e}",
              flush=True)
        ok = ok and same
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
