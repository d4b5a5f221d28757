"""Development aid: sustained (seconds-long, power-capped) throughput, SM clock and board power of the hot kernels run back
to back, next to cuBLAS on the FFN-1 shape.  `python tools/sustained_probe.py [seconds]`"""
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BF16 = torch.bfloat16


class Sampler:
    def __init__(self):
        self.rows = []
        self.proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                      "-lms", "100"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.proc.stdout:
            try:
                c, p = line.split(",")
                self.rows.append((time.time(), float(c), float(p)))
            except ValueError:
                pass

    def window(self, t0, t1):
        r = [(c, p) for t, c, p in self.rows if t0 + 0.5 <= t <= t1]
        if not r:
            return None, None
        r.sort()
        return r[len(r) // 2][0], sorted(p for _, p in r)[len(r) // 2]


def main():
    from videopainter_b200 import ops
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
    dev = "cuda"
    B, H, S, St, D = 2, 48, 17776, 226, 3072
    M = B * S
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).to(BF16)   # noqa: E731
    q, k, v = rn(B, H, S, 64), rn(B, H, S, 64), rn(B, H, S, 64)
    ao = torch.empty(B, S, D, dtype=BF16, device=dev)
    x = rn(M, D)
    w1, b1 = rn(4 * D, D, sc=0.02), rn(4 * D)
    ffm = torch.empty(M, 4 * D, dtype=BF16, device=dev)
    wt = w1.t().contiguous()
    sm = Sampler()
    cases = [("attention", lambda: ops.attention(q, k, v, ao, B, H, S, S, 0.125), 4.0 * B * H * S * S * 64),
             ("gemm_ff1_gelu", lambda: ops.gemm_gelu(x, w1, b1, ffm, M, 4 * D, D), 2.0 * M * 4 * D * D),
             ("cublas_ff1", lambda: torch.matmul(x, wt), 2.0 * M * 4 * D * D)]
    for name, fn, work in cases:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.time()
        n = 0
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        while time.time() - t0 < secs:
            for _ in range(20):
                fn()
            n += 20
            torch.cuda.synchronize()
        b.record()
        torch.cuda.synchronize()
        t1 = time.time()
        ms = a.elapsed_time(b) / n
        clk, pw = sm.window(t0, t1)
        print(f"{name}: {ms:.3f} ms/launch sustained over {t1 - t0:.1f} s = {work / ms / 1e9:.0f} TFLOP/s, SM clock {clk} MHz, {pw} W", flush=True)
        time.sleep(1.0)
    sm.proc.terminate()


if __name__ == "__main__":
    main()
