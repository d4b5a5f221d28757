#!/usr/bin/env python
"""The denoise loop of the any-length inpainting pipeline (PIPE:845-1034, without VAE / T5 / FluxFill: SURVEY §8f N3-N4) on
synthetic latents, driven entirely by this repo's B200 path: branch + backbone forwards (engine), fused step end, CFG halves
x Ulysses on several GPUs.  BASELINE.json configs 3 and 4:

    python examples/inpaint_loop.py --steps 50                                         # one 49-frame clip, 50-step CFG
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        examples/inpaint_loop.py --steps 50
    python examples/inpaint_loop.py --steps 8 --windows 3 --resample --lora-rank 256   # chained clips, ID resampling + LoRA

Prints one JSON line (rank 0): steps/s over the whole loop, per window, and a checksum of the final latents that must be
the same on every rank and for every GPU count (the sharded path is bit-identical to one GPU)."""
import argparse
import hashlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BF16 = torch.bfloat16
FULL = dict(num_attention_heads=48, attention_head_dim=64, in_channels=32, out_channels=16, time_embed_dim=512,
            text_embed_dim=4096, num_layers=42, sample_width=90, sample_height=60, sample_frames=49, patch_size=2,
            max_text_seq_length=226, use_rotary_positional_embeddings=True, use_learned_positional_embeddings=True)


def add_lora(model, rank, gen):
    """PEFT-style adapters on attn1.to_q / to_k / to_v / to_out.0 (TRAINID:1520-1526) as plain tensors in the state dict:
    the packer merges W + B A (inference scale 1.0)."""
    sd = model.state_dict()
    out = {}
    for k, v in sd.items():
        hit = [t for t in ("attn1.to_q.", "attn1.to_k.", "attn1.to_v.", "attn1.to_out.0.") if t in k]
        if hit:
            base = k.rsplit(".", 1)
            out[f"{base[0]}.base_layer.{base[1]}"] = v
            if base[1] == "weight":
                out[f"{base[0]}.lora_A.default.weight"] = (torch.randn(rank, v.shape[1], generator=gen, device="cpu") / rank).to(v)
                out[f"{base[0]}.lora_B.default.weight"] = (torch.randn(v.shape[0], rank, generator=gen, device="cpu") * 0.02).to(v)
        else:
            out[k] = v
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--windows", type=int, default=1)
    ap.add_argument("--layers", type=int, default=42)
    ap.add_argument("--resample", action="store_true", help="ID-resample attention processor (doubled K/V)")
    ap.add_argument("--lora-rank", type=int, default=0)
    ap.add_argument("--prev-clip-weight", type=float, default=0.5)
    ap.add_argument("--guidance-scale", type=float, default=6.0)
    ap.add_argument("--graphs", type=int, default=1, help="1: replay the forwards from CUDA graphs (videopainter_b200/graphs.py)")
    args = ap.parse_args()

    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch.distributed as dist
    import videopainter_b200 as vp
    from videopainter_b200 import engine, parallel
    from videopainter_b200.models import dims_from_module
    from videopainter_b200.rope import pipeline_rope
    from videopainter_b200.step_end import StepEnd

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    plan = parallel.init(world, rank).plan
    vp.enable_graphs(bool(args.graphs))

    torch.manual_seed(1234)
    tr = vp.CogVideoXTransformer3DModel(**dict(FULL, num_layers=args.layers), id_pool_resample_learnable=args.resample, device=dev, dtype=BF16)
    br = vp.CogvideoXBranchModel(**dict(FULL, num_layers=2), device=dev, dtype=BF16)
    with torch.no_grad():
        for m in (tr, br):
            m.patch_embed.pos_embedding.normal_(0, 0.5)
        for blk in br.branch_blocks:
            blk.weight.normal_(0, 0.02)
    pm_lora = None
    if args.lora_rank:        # a PEFT-wrapped state dict (base_layer + lora_A / lora_B), merged by the packer
        pm_lora = engine.pack_state_dict(add_lora(tr, args.lora_rank, torch.Generator().manual_seed(5)), dims_from_module(tr, False), dev)

    def run_transformer(**kw):
        if pm_lora is None:
            return tr(**kw, return_dict=False)
        return engine.transformer_forward(pm_lora, kw["hidden_states"], kw["encoder_hidden_states"], kw["timestep"], kw["image_rotary_emb"],
                                          kw["attention_kwargs"], kw["branch_block_samples"], kw["branch_block_masks"], False,
                                          True, True, kw["id_pool_resample_learnable"])

    # scheduler tables (CogVideoX-5B configuration: v-prediction, zero terminal SNR, trailing spacing; DPM:199-221, 293-298)
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    ac = torch.cumprod(1.0 - betas, dim=0)
    s = ac.sqrt()
    ac = ((s - s[-1]) * (s[0] / (s[0] - s[-1]))) ** 2
    timesteps = (np.round(np.arange(1000, 0, -1000 / args.steps)).astype(np.int64) - 1).tolist()
    se = StepEnd(ac, timesteps, guidance_scale=args.guidance_scale, use_dynamic_cfg=True)

    g = torch.Generator().manual_seed(42)       # same host-side stream on every rank: replicated pipeline state
    shape = (1, 13, 16, 60, 90)
    text = torch.randn(2, 226, 4096, generator=g).to(BF16).to(dev)
    rope = tuple(t.to(dev) for t in pipeline_rope(64, 480, 720, 13))
    prev_states, prev_mask = None, None
    window_ms = []
    for w in range(args.windows):
        gt = torch.randn(shape, generator=g).to(BF16).to(dev)                       # VAE-encoded video of this window
        image = torch.randn(shape, generator=g).to(BF16).to(dev)
        mask = (torch.rand((1, 13, 1, 60, 90), generator=g) > 0.75).to(BF16).to(dev)
        mask[:, 0] = 0                                                               # first frame is ground truth
        masked = gt * (1 - mask)
        noise0 = torch.randn(shape, generator=g).to(BF16).to(dev)
        latents = noise0.clone()                                                     # strength 1.0: start from pure noise
        old = None
        # the scheduler's draws (DPM:423, 431), in the order the reference makes them, drawn before the loop so that the
        # host generator does not sit between two steps (the second-order draw exists from the second step on)
        draws = []
        for i in range(len(timesteps)):
            n1 = torch.randn(shape, generator=g).to(BF16)
            n2 = torch.randn(shape, generator=g).to(BF16) if se.coefficients(i, i > 0)[-1] else None
            draws.append((n1.to(dev), None if n2 is None else n2.to(dev)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for i, t in enumerate(timesteps):
            lat2 = torch.cat([latents, latents])                                     # CFG batch (PIPE:937-942)
            lat_in = torch.cat([lat2, torch.cat([image, image])], dim=2)
            cond = torch.cat([torch.cat([masked, masked]), torch.cat([mask, mask])], dim=2)
            ts = torch.full((2,), t, dtype=torch.int64, device=dev)
            samples = br(hidden_states=lat2, encoder_hidden_states=text, branch_cond=cond, timestep=ts, image_rotary_emb=rope,
                         return_dict=False)[0]
            kw = None
            if prev_states is not None:
                kw = dict(prev_hidden_states=prev_states, prev_clip_weight=args.prev_clip_weight, prev_resample_mask=prev_mask)
            noise_pred, hs, rmask = run_transformer(hidden_states=lat_in, encoder_hidden_states=text, timestep=ts,
                                                    image_rotary_emb=rope, attention_kwargs=kw, branch_block_samples=samples,
                                                    branch_block_masks=torch.cat([mask, mask]),
                                                    id_pool_resample_learnable=args.resample, return_hidden_states=True,
                                                    return_resample_mask=True)
            if w < args.windows - 1 and i == len(timesteps) - 1:                     # PIPE:982-988
                next_states, next_mask = {k: h for k, h in enumerate(hs)}, rmask
            n1, n2 = draws[i]
            latents, old = se(i, noise_pred, latents, old, n1, n2, gt=gt, noise0=noise0, mask=mask)
        e1.record()
        torch.cuda.synchronize()
        window_ms.append(e0.elapsed_time(e1))
        if w < args.windows - 1:
            prev_states, prev_mask = next_states, next_mask
    digest = hashlib.sha256(latents.float().cpu().numpy().tobytes()).hexdigest()[:16]
    ok = True
    if world > 1:
        all_d = [None] * world
        dist.all_gather_object(all_d, digest)
        ok = len(set(all_d)) == 1
        tms = torch.tensor(window_ms, device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        window_ms = tms.tolist()
    if rank == 0:
        total = sum(window_ms)
        rec = {"example": "inpaint_loop", "n_gpus": world, "parallelism": plan.describe(), "layers": args.layers, "steps": args.steps,
               "windows": args.windows, "resample": args.resample, "lora_rank": args.lora_rank, "cuda_graphs": bool(args.graphs),
               "steps_per_s": args.steps * args.windows / (total / 1000.0), "window_ms": window_ms, "latents_sha256_16": digest,
               "ranks_agree": ok, "finite": bool(torch.isfinite(latents.float()).all())}
        os.write(real_stdout, (json.dumps(rec) + "\n").encode())
    parallel.shutdown()          # drops the captured graphs: NCCL cannot finalise a communicator that a live graph still uses
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
