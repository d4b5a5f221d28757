#!/usr/bin/env python
"""Benchmark of the denoising hot path (BASELINE.json metric: denoise steps/s at 49x480x720 with CFG).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's B200 path
    python bench.py --impl reference [--steps K] [--warmup W]      # reference algorithm on the host CPU (oracle port)

One "step" = one `branch(...)` + one `transformer(...)` call at CFG batch 2 on the CogVideoX-5B-I2V shape with the 2-layer
VideoPainter branch (PIPE:947-980), random-init bf16 weights, synthetic inputs.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FULL = dict(num_attention_heads=48, attention_head_dim=64, in_channels=32, out_channels=16, time_embed_dim=512,
            text_embed_dim=4096, num_layers=42, sample_width=90, sample_height=60, sample_frames=49, patch_size=2,
            max_text_seq_length=226, use_rotary_positional_embeddings=True, use_learned_positional_embeddings=True)
BRANCH_LAYERS = 2
S_TEXT, S_VIDEO, D_MODEL = 226, 13 * 30 * 45, 3072
S_TOTAL = S_TEXT + S_VIDEO


def step_flops(batch: int, layers: int = 42 + BRANCH_LAYERS) -> float:
    """Algorithmic FLOPs of one step: per block per sample 24 S D^2 + 4 S^2 D, plus embeds / heads."""
    S, D = S_TOTAL, D_MODEL
    block = 24.0 * S * D * D + 4.0 * S * S * D
    extra = 2.0 * S_VIDEO * 128 * D + 2.0 * S_VIDEO * 132 * D + 2 * 2.0 * S_TEXT * 4096 * D \
        + BRANCH_LAYERS * 2.0 * S_VIDEO * D * D + 2.0 * S_VIDEO * D * 64
    return batch * (layers * block + extra)


# ------------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_block_seconds(repeats: int, seq_video_frames: int = 13) -> float:
    """Median time of ONE full-size CogVideoXBlock forward (D=3072, 48x64 heads, S=17 776, batch 1, fp32) with the
    oracle restatement on all host threads."""
    import torch
    from oracle import cogvideox_oracle as O
    torch.set_num_threads(os.cpu_count())
    cfg = O.full_config(num_layers=1)
    g = torch.Generator().manual_seed(0)
    sd = {k: v for k, v in O.init_state_dict(cfg, 7).items() if k.startswith("transformer_blocks.0.")}
    h = torch.randn(1, S_VIDEO, D_MODEL, generator=g)
    e = torch.randn(1, S_TEXT, D_MODEL, generator=g)
    temb = torch.randn(1, 512, generator=g)
    rope = O.pipeline_rope(cfg, 480, 720, 13)
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.block(sd, "transformer_blocks.0.", cfg, h, e, temb, rope, head_chunk=4)
            times.append(time.perf_counter() - t0)
    times.sort()
    return times[len(times) // 2]


def cpu_tiny_step_ms(repeats: int = 20, warmup: int = 3) -> float:
    """Median time of one branch + transformer step of BASELINE.json configs[0] (tiny: 2 layers, 2x64 heads, 1-layer branch,
    13x8x8 latent, CFG batch 2, fp32) with the oracle restatement on all host threads (SURVEY.md §8d CPU baseline (i))."""
    import torch
    from oracle import cogvideox_oracle as O
    torch.set_num_threads(os.cpu_count())
    cfg, cfg_b = O.tiny_config(), O.tiny_config(num_layers=1)
    sd_t, sd_b = O.init_state_dict(cfg, 11), O.init_state_dict(cfg_b, 12, branch=True)
    inp = O.make_inputs(cfg, 1)
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            O.denoise_step(sd_t, sd_b, cfg, cfg_b, inp)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    times.sort()
    return 1e3 * times[len(times) // 2]


def cpu_baseline_record(block_s: float, tiny_ms: float = None) -> dict:
    blocks_per_step = (42 + BRANCH_LAYERS) * 2
    rec = {"value": 1.0 / (block_s * blocks_per_step), "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
           "sample": f"one full-size CogVideoXBlock forward (fp32, batch 1, S=17776) = {block_s:.2f} s on the host; "
                     f"steps/s = 1 / ({blocks_per_step} block-samples x that), embeds/head ignored"}
    if tiny_ms is not None:
        rec["tiny_config_step_ms"] = tiny_ms
        rec["tiny_config"] = "BASELINE.json configs[0]: 2 layers, 2x64 heads, 1-layer branch, 13x8x8 latent, CFG batch 2, fp32"
    return rec


def source_sha256(names) -> str:
    import hashlib
    h = hashlib.sha256()
    for n in names:
        with open(os.path.join(ROOT, "videopainter_b200", "csrc", n), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


ATTENTION_SOURCES = ("attention.cu", "attention.cuh", "common.cuh")


def gpu_eager_block_ms(dev, batch: int = 2, repeats: int = 3) -> float:
    """BASELINE ONLY, not on any product path: one full-size CogVideoXBlock (SURVEY.md §3.6 recipe) as the reference executes it —
    eager bf16 PyTorch, one ATen / cuBLAS call per op, `F.scaled_dot_product_attention` (AP:2192) — on this GPU, random weights.
    Gives the step a GPU-vs-GPU anchor: the reference's step is 44 of these plus embeds / head."""
    import torch
    import torch.nn.functional as F
    bf16 = torch.bfloat16
    S, St, D, H = S_TOTAL, S_TEXT, D_MODEL, 48
    g = torch.Generator(device=dev).manual_seed(5)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev, dtype=torch.float32) * sc).to(bf16)   # noqa: E731
    w = {n: rnd(o, i, sc=i ** -0.5) for n, (o, i) in dict(q=(D, D), k=(D, D), v=(D, D), o=(D, D), f1=(4 * D, D), f2=(D, 4 * D),
                                                          n1=(6 * D, 512), n2=(6 * D, 512)).items()}
    b = {n: rnd(t.shape[0], sc=0.02) for n, t in w.items()}
    ln = {n: (1 + rnd(d, sc=0.1), rnd(d, sc=0.1)) for n, d in dict(n1=D, n2=D, nq=64, nk=64).items()}
    from videopainter_b200.rope import pipeline_rope
    cos, sin = (t.to(dev) for t in pipeline_rope(64, 480, 720, 13))
    h, e, temb = rnd(batch, S - St, D), rnd(batch, St, D), rnd(batch, 512)

    def norm_zero(n, h, e):                                           # NRM:373-379
        sh, sc, gt, esh, esc, egt = F.linear(F.silu(temb), w[n], b[n]).chunk(6, dim=1)
        h = F.layer_norm(h, (D,), *ln[n], 1e-5) * (1 + sc)[:, None] + sh[:, None]
        e = F.layer_norm(e, (D,), *ln[n], 1e-5) * (1 + esc)[:, None] + esh[:, None]
        return h, e, gt[:, None], egt[:, None]

    def rope(x):                                                      # EMB:675-692
        xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
        rot = torch.stack([-xi, xr], dim=-1).flatten(3)
        return (x.float() * cos + rot.float() * sin).to(x.dtype)

    def block(h, e):                                                  # T3D:125-184, AP:2107-2209
        nh, ne, gt, egt = norm_zero("n1", h, e)
        x = torch.cat([ne, nh], dim=1)
        q, k, v = (F.linear(x, w[n], b[n]).view(batch, S, H, 64).transpose(1, 2) for n in "qkv")
        q, k = F.layer_norm(q, (64,), *ln["nq"], 1e-6), F.layer_norm(k, (64,), *ln["nk"], 1e-6)
        q[:, :, St:] = rope(q[:, :, St:])
        k[:, :, St:] = rope(k[:, :, St:])
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(batch, S, D)
        o = F.linear(o, w["o"], b["o"])
        h, e = h + gt * o[:, St:], e + egt * o[:, :St]
        nh, ne, gt, egt = norm_zero("n2", h, e)
        f = F.linear(F.gelu(F.linear(torch.cat([ne, nh], dim=1), w["f1"], b["f1"]), approximate="tanh"), w["f2"], b["f2"])
        return h + gt * f[:, St:], e + egt * f[:, :St]

    times = []
    with torch.no_grad():
        for i in range(repeats + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            block(h, e)
            e1.record()
            torch.cuda.synchronize()
            if i > 0:
                times.append(e0.elapsed_time(e1))
    times.sort()
    return times[len(times) // 2]


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = []
    for i in range(args.warmup + args.steps):
        t = cpu_block_seconds(1)
        if i >= args.warmup:
            times.append(t)
    block_s = sum(times) / len(times)
    rec = cpu_baseline_record(block_s, cpu_tiny_step_ms())
    out = {"impl": "reference", "metric": "denoise steps/s (49x480x720, CFG)", "value": rec["value"], "unit": "steps/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / rec["value"],
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
           "config": {"workload": "CogVideoX-5B-I2V + 2-layer VideoPainter branch, 49x480x720, CFG batch 2; each timed step is "
                                  "a bounded sample: one full-size block (see cpu_baseline.sample)"},
           "cpu_baseline": rec,
           "e2e": {"value": rec["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--layers", type=int, default=42, help="(debug only) fewer layers -> line is marked invalid")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--breakdown", default=None, help="write the per-kernel breakdown JSON to this path")
    ap.add_argument("--graphs", type=int, default=int(os.environ.get("VP_B200_GRAPH", "1")),
                    help="1 (default): the two forwards of a step are replayed from CUDA graphs (videopainter_b200/graphs.py); "
                         "0: every kernel is launched from Python")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    # rank 0 prints exactly ONE line on stdout: libraries that write to fd 1 (NCCL's version banner) go to stderr instead
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: str) -> None:
        os.write(real_stdout, (line + "\n").encode())

    import torch
    import torch.distributed as dist
    import videopainter_b200 as vp
    from videopainter_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from videopainter_b200 import parallel
    plan = parallel.init(world, rank).plan        # CFG halves on disjoint GPU groups x Ulysses inside each (parallel.py)

    bf16 = torch.bfloat16
    torch.manual_seed(1234)
    cfg = dict(FULL, num_layers=args.layers)
    tr = vp.CogVideoXTransformer3DModel(**cfg, device=dev, dtype=bf16)
    br = vp.CogvideoXBranchModel(**dict(FULL, num_layers=BRANCH_LAYERS), device=dev, dtype=bf16)
    with torch.no_grad():
        for m in (tr, br):
            m.patch_embed.pos_embedding.normal_(0, 0.5)
    B_global = 2
    B = B_global          # every rank is handed the whole CFG batch, as the pipeline would; the model takes its share

    # synthetic inputs of the pipeline's shapes (PIPE:937-945), pinned on the host for the end-to-end leg
    g = torch.Generator().manual_seed(99)
    host = {
        "latents": torch.randn(B, 13, 16, 60, 90, generator=g).to(bf16),
        "image": torch.randn(B, 13, 16, 60, 90, generator=g).to(bf16),
        "masked": torch.randn(B, 13, 16, 60, 90, generator=g).to(bf16),
        "mask": (torch.rand(B, 13, 1, 60, 90, generator=g) > 0.75).to(bf16),
        "text": torch.randn(B, 226, 4096, generator=g).to(bf16),
        "timestep": torch.full((B,), 999, dtype=torch.int64),
    }
    host = {k: v.pin_memory() for k, v in host.items()}
    from videopainter_b200.rope import pipeline_rope
    rope = tuple(t.to(dev) for t in pipeline_rope(64, 480, 720, 13))
    out_host = torch.empty(B, 13, 16, 60, 90, dtype=torch.float32).pin_memory()

    def step(d):
        lat_in = torch.cat([d["latents"], d["image"]], dim=2)
        cond = torch.cat([d["masked"], d["mask"]], dim=2)
        samples = br(hidden_states=d["latents"], encoder_hidden_states=d["text"], branch_cond=cond, timestep=d["timestep"],
                     image_rotary_emb=rope, return_dict=False)[0]
        noise, hs, rmask = tr(hidden_states=lat_in, encoder_hidden_states=d["text"], timestep=d["timestep"],
                              image_rotary_emb=rope, branch_block_samples=samples, branch_block_masks=d["mask"],
                              return_hidden_states=True, return_resample_mask=True, return_dict=False)
        return noise

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    resident = {k: v.to(dev) for k, v in host.items()}
    import hashlib

    def timed(fn):
        """K calls of fn between a barrier + synchronize on both sides; CUDA events on the launching stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / args.steps

    last = {}

    def resident_step():
        last["noise"] = step(resident)

    def e2e_step():
        d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        noise = step(d)
        out_host.copy_(noise.float(), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    vp.enable_graphs(bool(args.graphs))
    with torch.no_grad():
        # with graphs the first call of a forward is eager and the second captures it: both must stay out of the timed region
        for _ in range(max(args.warmup, 2) if args.graphs else args.warmup):
            step(resident)
        # ---------------- device-resident timing (value): no per-op events, nothing but the step's own launches ----------------
        sampler = ClockSampler(local)
        sampler.start()
        launches0 = ops.launch_count
        torch.cuda.cudart().cudaProfilerStart()      # `ncu --profile-from-start off` then lists exactly the timed steps
        ms_per_step = timed(resident_step)
        torch.cuda.cudart().cudaProfilerStop()
        clocks = sampler.stop()
        launches = ops.launch_count - launches0
        # every rank holds the complete noise prediction; its SHA-256 must not depend on N (sharded == single GPU, bit for bit)
        noise_sha = hashlib.sha256(last["noise"].contiguous().view(torch.int16).cpu().numpy().tobytes()).hexdigest()
        noise_absmax = float(last["noise"].float().abs().max())
        # ---------------- end-to-end timing: host buffers in, host result out ----------------
        e2e_ms = timed(e2e_step)
        # ---------------- per-kernel pass: the same K steps with CUDA events around every launch (roofline / breakdown) --------
        vp.enable_graphs(False)                  # events around every launch need the launches on the stream
        ops.start_profile()
        prof_ms_per_step = timed(resident_step)
        prof = ops.stop_profile()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out_host.numel() * out_host.element_size()

    last.clear()
    uses_p2p = bool(parallel.current() and parallel.current().p2p)
    parallel.shutdown()                        # drops captured graphs: NCCL cannot finalise while graphs hold its collectives
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s"
    breakdown = {}
    tot_ms = sum(v[1] for v in prof.values())
    for name, (n, t, w) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        breakdown[name] = {"launches": n, "ms_per_step": t / args.steps, "share": t / tot_ms if tot_ms else 0.0,
                           "avg_launch_ms": t / n, "tflops": (w / (t * 1e-3) / 1e12) if (w and name != "ln_modulate") else None,
                           "gbs": (w / (t * 1e-3) / 1e9) if name == "ln_modulate" else None}
    # DRAM traffic of the dominant kernel per launch: from the committed ncu --set full capture of one launch at this shape —
    # only if that capture was taken from the kernel source that is built now (source hash recorded with the capture)
    traffic, traffic_note = None, "no ncu --set full capture of the current attention kernel source is committed"
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_attention_latest.json")))
        to_bytes = lambda v: float(v.split()[0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[v.split()[1]]   # noqa: E731
        if ncu.get("source_sha256") == source_sha256(ATTENTION_SOURCES):
            traffic = (to_bytes(ncu["dram__bytes_read.sum"]) + to_bytes(ncu["dram__bytes_write.sum"]))
            traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum of one attention launch (B=2, 48 heads, S=17776) from "
                            "profiles/ncu_attention_latest.json (same kernel source, sha256 checked); algorithmic q+k+v+o bytes = 873.6 MB")
        else:
            traffic_note = "profiles/ncu_attention_latest.json was captured from a different attention kernel source: dropped"
    except Exception:
        pass
    dom = max(prof.items(), key=lambda kv: kv[1][1])
    dn, (n_l, t_l, w_l) = dom
    achieved = w_l / (t_l * 1e-3) / 1e12
    roofline = {"kernel": dn, "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "traffic": traffic if (dn == "attention" and world == 1) else None,
                "traffic_note": traffic_note, "peak_source": peak_src,
                "flops_per_launch": w_l / n_l, "avg_launch_ms": t_l / n_l, "share_of_step": t_l / tot_ms,
                "timed_in": f"a separate pass of the same {args.steps} steps with CUDA events around every launch "
                            f"({prof_ms_per_step:.1f} ms/step with events vs {ms_per_step:.1f} without)"}
    flops = step_flops(B_global, args.layers + BRANCH_LAYERS)
    out = {"metric": "denoise steps/s (49x480x720, CFG)", "value": 1000.0 / ms_per_step, "unit": "steps/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": "CogVideoX-5B-I2V (42 layers, 48x64 heads) + 2-layer VideoPainter branch, one denoise step at "
                                  "49x480x720 (17776 tokens), CFG batch 2, return_hidden_states=True as PIPE:967-980",
                      "parallelism": plan.describe() + ((", all-to-all fused into the GEMM / attention epilogues over NVLink peer memory"
                                                         if uses_p2p else ", NCCL all-to-all") if plan.sp > 1 else ""),
                      "l2": "per-step working set (11.7 GB weights, 218 MB activations per "
                      "layer) exceeds the 126 MB L2; no explicit flush", "random_init": True},
           "step_tflops": flops / (ms_per_step * 1e-3) / 1e12,
           "frac_of_bf16_peak": {"sustained": flops / (ms_per_step * 1e-3) / 1e12 / (peak_tf * world),
                                 "burst": flops / (ms_per_step * 1e-3) / 1e12 / (peaks.get("bf16_tflops", 1650.0) * world),
                                 "note": "whole-job algorithmic FLOP/s over n_gpus x the measured cuBLAS peak"},
           "roofline": roofline, "clocks": clocks, "gpu_launches": launches, "cuda_graphs": bool(args.graphs),
           "noise_sha256": noise_sha, "noise_absmax": noise_absmax,
           "e2e": {"value": 1000.0 / e2e_ms, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": e2e_ms}}
    if args.layers != 42:
        out["invalid"] = "debug run with fewer layers than the named config"
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline_record(cpu_block_seconds(1), cpu_tiny_step_ms())
    if not args.no_gpu_eager and world == 1:
        try:
            blk = gpu_eager_block_ms(dev)
            out["gpu_eager_reference"] = {
                "block_ms": blk, "steps_per_s_extrapolated": 1000.0 / (blk * (42 + BRANCH_LAYERS)),
                "what": "BASELINE ONLY: one full-size CogVideoXBlock (CFG batch 2, S=17776) in eager bf16 PyTorch with "
                        "F.scaled_dot_product_attention, as the reference executes it, on this GPU; x 44 blocks, embeds / head "
                        "ignored; random weights"}
        except Exception as e:   # noqa: BLE001 - a baseline must never break the measurement
            out["gpu_eager_reference"] = {"unavailable": str(e)[:200]}
    if args.breakdown:
        with open(args.breakdown, "w") as f:
            json.dump({"ms_per_step": ms_per_step, "ms_per_step_with_events": prof_ms_per_step, "kernels": breakdown,
                       "clocks": clocks}, f, indent=1)
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
