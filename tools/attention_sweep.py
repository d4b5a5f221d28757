"""BASELINE.json config 5: attention kernel sweep, B*H = 96 heads of 64, bf16, non-causal, scale 1/8, against the reference's
path on the same box (F.scaled_dot_product_attention, AP:2192) for time and against fp32 math for error (small slice).
Development / reporting tool: `python tools/attention_sweep.py > gpurun_out/attention_sweep.json`."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BF16 = torch.bfloat16


def timeit(fn, flush, iters=5):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    from videopainter_b200 import ops
    dev = "cuda"
    B, H = 2, 48
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    rows = []
    for S, Skv in [(4096, 4096), (8192, 8192), (12288, 12288), (17776, 17776), (24576, 24576), (17776, 35552), (36864, 36864)]:
        q = torch.randn(B, H, S, 64, device=dev, generator=g).to(BF16)
        k = torch.randn(B, H, Skv, 64, device=dev, generator=g).to(BF16)
        v = torch.randn(B, H, Skv, 64, device=dev, generator=g).to(BF16)
        out = torch.empty(B, S, H * 64, dtype=BF16, device=dev)
        if Skv == 2 * S:      # the resample processor's two segments
            half = Skv // 2
            k0, k1, v0, v1 = (t.contiguous() for t in (k[:, :, :half], k[:, :, half:], v[:, :, :half], v[:, :, half:]))
            fn = lambda: ops.attention(q, k0, v0, out, B, H, S, half, 0.125, k1=k1, v1=v1, kv_len1=half)   # noqa: E731
        else:
            fn = lambda: ops.attention(q, k, v, out, B, H, S, Skv, 0.125)   # noqa: E731
        ms = timeit(fn, flush)
        ms_ref = timeit(lambda: F.scaled_dot_product_attention(q, k, v), flush)
        # error on one (batch, head) against fp32 math
        ref = torch.softmax(q[0, 0].float() @ k[0, 0].float().t() / 8.0, dim=-1) @ v[0, 0].float()
        got = out[0, :, :64].float()
        sd = F.scaled_dot_product_attention(q[:1, :1], k[:1, :1], v[:1, :1])[0, 0].float()
        fl = 4.0 * B * H * S * Skv * 64
        rows.append({"seq_q": S, "seq_kv": Skv, "ms": ms, "tflops": fl / ms / 1e9, "sdpa_ms": ms_ref, "sdpa_tflops": fl / ms_ref / 1e9,
                     "speedup_vs_sdpa": ms_ref / ms, "max_abs_err_vs_fp32": (got - ref).abs().max().item(),
                     "sdpa_max_abs_err_vs_fp32": (sd - ref).abs().max().item()})
        print(json.dumps(rows[-1]), flush=True)
        del q, k, v, out, ref


if __name__ == "__main__":
    main()
