"""Shared helpers for the GPU parity tests (test infrastructure)."""
import torch


def cos_sim(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def report(name: str, got: torch.Tensor, ref: torch.Tensor) -> str:
    g, r = got.float(), ref.float()
    err = (g - r).abs()
    denom = r.abs().max().item() + 1e-30
    msg = (f"[{name}] cos={cos_sim(g, r):.7f} max_abs_err={err.max().item():.4e} ref_absmax={denom:.4e} "
           f"rel_max={err.max().item() / denom:.4e} mean_abs_err={err.mean().item():.4e} "
           f"nan={int(torch.isnan(g).sum())} shape={tuple(g.shape)}")
    return msg


def assert_close_bf16(name: str, got: torch.Tensor, ref: torch.Tensor, cos_min=0.9999, rel_max=2e-2):
    """bf16 result vs fp32 reference: cosine and max-abs error relative to the reference's max magnitude."""
    msg = report(name, got, ref)
    print(msg)
    g, r = got.float(), ref.float()
    assert not torch.isnan(g).any(), msg
    assert cos_sim(g, r) >= cos_min, msg
    assert (g - r).abs().max().item() <= rel_max * (r.abs().max().item() + 1e-30), msg


def block_error_map(got: torch.Tensor, ref: torch.Tensor, rb=8, cb=32, thr=5e-2) -> str:
    """Coarse map of which (rb x cb) blocks of a 2-D result are wrong — helps to tell swizzle / descriptor bugs apart."""
    g, r = got.float(), ref.float()
    R, Cc = g.shape
    R2, C2 = R // rb * rb, Cc // cb * cb
    e = (g[:R2, :C2] - r[:R2, :C2]).abs().reshape(R2 // rb, rb, C2 // cb, cb).amax(dim=(1, 3))
    bad = e > thr * (r.abs().max().item() + 1e-30)
    lines = []
    for i in range(min(bad.shape[0], 32)):
        lines.append("".join("X" if bad[i, j] else "." for j in range(min(bad.shape[1], 64))))
    return "\n".join(lines)
