"""Development aid for the attention kernel: one process = one build of the library (VP_B200_LIB selects a variant built with
videopainter_b200.build.build_variant).  Checks the result against explicit fp32 softmax attention on two heads of the
production sequence, times the production shape (B=2, 48 heads, S=17776) with CUDA events, optionally times
F.scaled_dot_product_attention on the same box for scale, and dumps the in-kernel phase trace of a VP_ATTN_TRACE=1 build.

    VP_B200_LIB=$PWD/videopainter_b200/csrc/libvp_b200_<name>.so python tools/attn_lab.py [--sdpa] [--trace out.json] [--seq S]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BF16 = torch.bfloat16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sdpa", action="store_true")
    ap.add_argument("--trace", default=None)
    ap.add_argument("--seq", type=int, default=17776)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    from videopainter_b200 import ops
    from videopainter_b200._lib import LIB, lib
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g).to(BF16)   # noqa: E731
    rec = {"lib": os.path.basename(LIB), "seq": args.seq}
    S = args.seq
    if not args.no_check:
        q, k, v = rn(1, 2, S, 64), rn(1, 2, S, 64), rn(1, 2, S, 64)
        out = torch.zeros(1, S, 128, dtype=BF16, device=dev)
        ops.attention(q, k, v, out, 1, 2, S, S, 0.125)
        torch.cuda.synchronize()
        ref = torch.softmax(q.float() @ k.float().transpose(-1, -2) / 8.0, dim=-1) @ v.float()
        ref = ref.transpose(1, 2).reshape(1, S, 128)
        err = (out.float() - ref).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item()
        rec.update(max_abs_err=err, cos=cos, ref_absmax=ref.abs().max().item(), nan=int(torch.isnan(out.float()).sum()))
        del ref
    B, H = 2, 48
    q, k, v = rn(B, H, S, 64), rn(B, H, S, 64), rn(B, H, S, 64)
    ao = torch.empty(B, S, H * 64, dtype=BF16, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]

    fl = 4.0 * B * H * S * S * 64
    ms = timeit(lambda: ops.attention(q, k, v, ao, B, H, S, S, 0.125))
    rec.update(ms=ms, tflops=fl / ms / 1e9)
    if args.sdpa:
        ms2 = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
        rec.update(sdpa_ms=ms2, sdpa_tflops=fl / ms2 / 1e9)
    print(json.dumps(rec), flush=True)
    if args.trace:
        L = lib()
        if not hasattr(L, "vp_debug_attn_trace"):
            raise SystemExit("this build has no trace support (VP_ATTN_TRACE=1)")
        n = 18 * 48 * 8
        buf = (ctypes.c_ulonglong * n)()
        L.vp_debug_attn_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
        got = L.vp_debug_attn_trace(buf, n)
        vals = list(buf)[:max(got, 0)]
        t0 = min(x for x in vals if x)
        ev = [[[(vals[(w * 48 + t) * 8 + e] - t0) if vals[(w * 48 + t) * 8 + e] else None for e in range(8)] for t in range(48)]
              for w in range(18)]
        json.dump({"events": "softmax warp: 0 wait S, 1 S ready, 2 S loaded, 3 max / rescale done, 4 exps issued, 5 arrived on P; "
                             "warp 16 (MMA issuer), per key tile: 4t+0 P(t, half 0) seen, 4t+1 its P V issued, 4t+2 P(t, half 1) seen, 4t+3 its P V and the next Q K^T issued", "clk": ev}, open(args.trace, "w"))


if __name__ == "__main__":
    main()
