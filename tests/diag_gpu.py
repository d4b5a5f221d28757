"""Diagnostic dump for blind kernel bring-up (not a test): runs tiny structured problems through the tcgen05 kernels and
saves raw outputs + expectations to gpurun_out/diag_*.pt so that descriptor / swizzle mistakes can be decoded offline.
Each probe runs in its own process (a trapped kernel poisons the CUDA context)."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
BF16 = torch.bfloat16


def probe_gemm():
    from videopainter_b200 import ops
    M, N, K = 128, 256, 64
    res = {}
    # A = one-hot rows: out[m, n] = W[n, m % K]  -> reveals how (row, k) pairs are matched
    A = torch.zeros(M, K, device="cuda")
    A[torch.arange(M), torch.arange(M) % K] = 1
    W = (torch.arange(N, device="cuda")[:, None] + torch.arange(K, device="cuda")[None, :] / 64.0)
    out = torch.zeros(M, N, dtype=BF16, device="cuda")
    ops.gemm_bias(A.to(BF16), W.to(BF16), None, out, M, N, K, rows_per_batch=M, out_batch_rows=0, out_row_offset=0)
    torch.cuda.synchronize()
    res["onehot_out"] = out.float().cpu()
    res["onehot_ref"] = (A.to(BF16).float() @ W.to(BF16).float().t()).cpu()
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(M, K, device="cuda", generator=g).to(BF16)
    W = torch.randn(N, K, device="cuda", generator=g).to(BF16)
    ops.gemm_bias(A, W, None, out, M, N, K, rows_per_batch=M, out_batch_rows=0, out_row_offset=0)
    torch.cuda.synchronize()
    res["rand_out"] = out.float().cpu()
    res["rand_ref"] = (A.float() @ W.float().t()).cpu()
    e = (res["rand_out"] - res["rand_ref"]).abs().max().item()
    print("gemm probe: max err", e, "onehot err", (res["onehot_out"] - res["onehot_ref"]).abs().max().item())
    torch.save(res, os.path.join(OUT, "diag_gemm.pt"))


def probe_attn():
    from videopainter_b200 import ops
    res = {}
    B, H, S = 1, 1, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    # (1) uniform attention (q = 0): out = mean of V rows -> isolates the P·V MMA (TMEM A operand, MN-major V)
    q = torch.zeros(B, H, S, 64, device="cuda", dtype=BF16)
    k = torch.randn(B, H, S, 64, device="cuda", generator=g).to(BF16)
    v = torch.randn(B, H, S, 64, device="cuda", generator=g).to(BF16)
    out = torch.zeros(B, S, 64, dtype=BF16, device="cuda")
    ops.attention(q, k, v, out, B, H, S, S, 0.125)
    torch.cuda.synchronize()
    res["uniform_out"] = out.float().cpu()
    res["uniform_ref"] = v.float().mean(dim=2).cpu()
    # (2) one-hot attention: huge q·k on the diagonal -> out[i] = V[i]: isolates the QKᵀ MMA + P column order
    e = torch.zeros(S, 64, device="cuda")
    e[torch.arange(S), torch.arange(S) % 64] = 30.0
    e[torch.arange(S), (torch.arange(S) // 64 + 7) % 64] += 20.0 * (torch.arange(S, device="cuda") // 64)
    q2 = e.to(BF16)[None, None]
    ops.attention(q2, q2, v, out, B, H, S, S, 1.0)
    torch.cuda.synchronize()
    res["onehot_out"] = out.float().cpu()
    s = (q2.float() @ q2.float().transpose(-1, -2))
    res["onehot_ref"] = (torch.softmax(s, -1) @ v.float())[0, 0].cpu()
    res["v"] = v.float().cpu()
    print("attn probe: uniform err", (res["uniform_out"] - res["uniform_ref"]).abs().max().item(),
          "onehot err", (res["onehot_out"] - res["onehot_ref"]).abs().max().item())
    torch.save(res, os.path.join(OUT, "diag_attn.pt"))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:
        {"gemm": probe_gemm, "attn": probe_attn}[sys.argv[1]]()
    else:
        for name in ("gemm", "attn"):
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=300)
            print(f"== probe {name}: exit {r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}")
