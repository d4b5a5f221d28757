"""Stand-alone timing of the hot kernels at the production shapes (CUDA events, 3 warm-ups, L2 flushed between
iterations).  Also the command profiled by ncu (profiles/): `python tests/prof_kernels.py [--iters N]`."""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BF16 = torch.bfloat16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from videopainter_b200 import ops
    dev = "cuda"
    B, H, S, St, D = 2, 48, 17776, 226, 3072
    M = B * S
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).to(BF16)   # noqa: E731
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}

    def timeit(name, fn, work, unit):
        if args.only and args.only not in name:
            return
        for _ in range(3):
            fn()
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        res[name] = {"ms": ms, unit: work / (ms * 1e-3) / (1e12 if unit == "tflops" else 1e9)}
        print(name, res[name], flush=True)

    q, k, v = rn(B, H, S, 64), rn(B, H, S, 64), rn(B, H, S, 64)
    ao = torch.empty(B, S, D, dtype=BF16, device=dev)
    timeit("attention_S17776", lambda: ops.attention(q, k, v, ao, B, H, S, S, 0.125), 4.0 * B * H * S * S * 64, "tflops")
    timeit("attention_2seg", lambda: ops.attention(q, k, v, ao, B, H, S, S, 0.125, k1=k, v1=v, kv_len1=S),
           8.0 * B * H * S * S * 64, "tflops")
    # the peer-memory instantiation (Ulysses output stores), here with a single "peer": same work, batch folded into heads
    ao1 = torch.empty(1, S, B * H * 64, dtype=BF16, device=dev)
    q1, k1, v1 = q.view(1, B * H, S, 64), k.view(1, B * H, S, 64), v.view(1, B * H, S, 64)
    timeit("attention_peer_S17776", lambda: ops.attention_peer(q1, k1, v1, [ao1.data_ptr()], 0, B * H * 64, B * H, S, S, 0.125),
           4.0 * B * H * S * S * 64, "tflops")
    x = rn(M, D)
    w_qkv, b_qkv = rn(3 * D, D, sc=0.02), rn(3 * D)
    nq = (rn(64), rn(64))
    ang = torch.rand(S - St, 32, device=dev) * 6.28          # pair-repeated tables as the pipeline builds them (EMB:641-642)
    cos, sin = torch.cos(ang).repeat_interleave(2, 1).contiguous(), torch.sin(ang).repeat_interleave(2, 1).contiguous()
    pairs = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).reshape(S - St, 64).contiguous()
    timeit("gemm_qkv", lambda: ops.gemm_qkv(x, w_qkv, b_qkv, M, D, S, H, 0, q, k, v, nq, nq, 1e-6, (cos, sin, pairs), St),
           2.0 * M * 3 * D * D, "tflops")
    w_o, b_o = rn(D, D, sc=0.02), rn(D)
    gate = torch.randn(B, 6 * D, device=dev)
    xo = torch.empty(B, S, D, dtype=BF16, device=dev)
    timeit("gemm_out_gate_res", lambda: ops.gemm_gate_residual(x, w_o, b_o, xo, M, D, D, S, S, 0, x, S, 0, gate=gate,
                                                               gate_video_off=2 * D, gate_text_off=5 * D, text_len=St),
           2.0 * M * D * D, "tflops")
    w1, b1 = rn(4 * D, D, sc=0.02), rn(4 * D)
    ffm = torch.empty(M, 4 * D, dtype=BF16, device=dev)
    timeit("gemm_ff1_gelu", lambda: ops.gemm_gelu(x, w1, b1, ffm, M, 4 * D, D), 2.0 * M * 4 * D * D, "tflops")
    w2, b2 = rn(D, 4 * D, sc=0.01), rn(D)
    timeit("gemm_ff2_gate_res", lambda: ops.gemm_gate_residual(ffm, w2, b2, xo, M, D, 4 * D, S, S, 0, x, S, 0, gate=gate,
                                                               gate_video_off=2 * D, gate_text_off=5 * D, text_len=St),
           2.0 * M * 4 * D * D, "tflops")
    timeit("gemm_plain_bias_3072", lambda: ops.gemm_bias(x, w_o, b_o, xo, M, D, D, M, 0, 0), 2.0 * M * D * D, "tflops")
    gam = rn(D)
    timeit("ln_modulate", lambda: ops.ln_modulate(x, S, 0, xo, B, S, D, gam, gam, 1e-5, gate, (0, D, 3 * D, 4 * D), St),
           4.0 * M * D, "gbs")
    x3 = x.view(B, S, D)
    timeit("ln_modulate_copy_for_scale", lambda: xo.copy_(x3), 4.0 * M * D, "gbs")   # same bytes moved by a plain device copy
    # all adaLN-zero tables of one backbone forward: 42 blocks x 2 LayerNormZero x [18432, 512] bf16 = 1.59 GB of weights
    n_ada = 42 * 2 * 6 * D
    w_ada, b_ada = torch.empty(n_ada, 512, dtype=BF16, device=dev).normal_(0, 0.02), rn(n_ada)
    temb = torch.randn(B, 512, device=dev)
    tab = torch.empty(B, n_ada, dtype=torch.float32, device=dev)
    timeit("gemv_adaln_tables", lambda: ops.gemv(temb, w_ada, b_ada, True, out=tab), 2.0 * n_ada * 512, "gbs")
    # the library GEMM / SDPA of this box, for scale only (not part of the product path)
    wt = w1.t().contiguous()
    timeit("cublas_ff1_for_scale", lambda: torch.matmul(x, wt), 2.0 * M * 4 * D * D, "tflops")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "prof_kernels.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
