"""GPU parity of the token-wise kernels against plain PyTorch fp32 (floating-point kernels keep a torch reference)."""
import math

import pytest
import torch
import torch.nn.functional as F

from _util import assert_close_bf16, report

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from videopainter_b200 import ops as _ops
    return _ops


def _randn(*shape, seed=0, dtype=BF16, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda", dtype=torch.float32) * scale).to(dtype)


@pytest.mark.parametrize("D,B,S,St", [(128, 2, 224, 16), (3072, 2, 500, 226), (1920, 1, 77, 10)])
def test_ln_modulate(ops, D, B, S, St):
    x = _randn(B, S, D, seed=1, scale=2.0)
    gamma, beta = _randn(D, seed=2) * 0.1 + 1, _randn(D, seed=3) * 0.1
    mod = _randn(B, 6 * D, seed=4, dtype=torch.float32) * 0.3
    y = torch.empty(B * S, D, dtype=BF16, device="cuda")
    ops.ln_modulate(x, S, 0, y, B, S, D, gamma, beta, 1e-5, mod, (0, D, 3 * D, 4 * D), St)
    xn = F.layer_norm(x.float(), (D,), gamma.float(), beta.float(), 1e-5)
    sh, sc, _, esh, esc, _ = mod.chunk(6, dim=1)
    ref = torch.cat([xn[:, :St] * (1 + esc[:, None]) + esh[:, None], xn[:, St:] * (1 + sc[:, None]) + sh[:, None]], dim=1)
    assert_close_bf16(f"ln_modulate D={D}", y.view(B, S, D), ref, cos_min=0.99999, rel_max=1e-2)


def test_ln_final(ops):
    D, B, S, St = 3072, 2, 300, 226
    Sv = S - St
    x = _randn(B, S, D, seed=1, scale=3.0)
    g1, b1, g2, b2 = (_randn(D, seed=s) * 0.1 + (1 if s % 2 == 0 else 0) for s in (2, 3, 4, 5))
    mod = _randn(B, 2 * D, seed=6, dtype=torch.float32) * 0.3
    y = torch.empty(B * Sv, D, dtype=BF16, device="cuda")
    ops.ln_final(x, S, St, y, B, Sv, D, g1, b1, g2, b2, 1e-5, mod, 0, D)
    h = F.layer_norm(x.float(), (D,), g1.float(), b1.float(), 1e-5)[:, St:]
    h = F.layer_norm(h, (D,), g2.float(), b2.float(), 1e-5)
    ref = h * (1 + mod[:, None, D:]) + mod[:, None, :D]
    assert_close_bf16("ln_final", y.view(B, Sv, D), ref, cos_min=0.99999, rel_max=1e-2)


@pytest.mark.parametrize("B,N,K,act", [(2, 512, 3072, False), (2, 18432, 512, True), (1, 128, 64, True)])
def test_gemv(ops, B, N, K, act):
    x = _randn(B, K, seed=1, dtype=torch.float32)
    W = _randn(N, K, seed=2) * (1 / math.sqrt(K))
    b = _randn(N, seed=3) * 0.1
    out = ops.gemv(x, W, b, act)
    ref = F.linear(F.silu(x) if act else x, W.float(), b.float())
    print(report("gemv", out, ref))
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)


def test_time_sinusoid(ops):
    t = torch.tensor([999, 19], device="cuda", dtype=torch.int64)
    out = ops.time_sinusoid(t, 3072, True, 0.0)
    half = 1536
    ex = torch.exp(-math.log(10000) * torch.arange(half, device="cuda", dtype=torch.float32) / half)
    ang = t[:, None].float() * ex[None]
    ref = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)
    print(report("time_sinusoid", out, ref))
    torch.testing.assert_close(out, ref, rtol=0, atol=2e-3)   # fp32 sin/cos of arguments up to ~1e3


@pytest.mark.parametrize("C0,C1,H,W", [(32, 0, 8, 8), (16, 17, 60, 90)])
def test_patchify_matches_conv(ops, C0, C1, H, W):
    BF, D = 4, 64
    s0 = _randn(BF, C0, H, W, seed=1)
    s1 = _randn(BF, C1, H, W, seed=2) if C1 else None
    C = C0 + C1
    kpad = (4 * C + 63) // 64 * 64
    A = torch.full((BF * (H // 2) * (W // 2), kpad), 7.0, dtype=BF16, device="cuda")
    ops.patchify(s0, s1, A, BF, H, W, kpad)
    src = s0 if s1 is None else torch.cat([s0, s1], dim=1)
    Wc = _randn(D, C, 2, 2, seed=3)
    ref = F.conv2d(src.float(), Wc.float(), stride=2).flatten(2).transpose(1, 2).reshape(-1, D)   # (bf, y, x) rows
    got = A[:, : 4 * C].float() @ Wc.float().reshape(D, -1).t()
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)
    assert (A[:, 4 * C:] == 0).all()


def test_mask_pool_and_unpatchify(ops):
    BF, H, W = 6, 8, 12
    m = (torch.rand(BF, 1, H, W, device="cuda") > 0.7).to(BF16)
    out = torch.empty(BF * (H // 2) * (W // 2), dtype=torch.uint8, device="cuda")
    ops.mask_pool(m, out, BF, H, W)
    ref = (F.avg_pool2d(m.float(), 2) > 0).flatten()
    assert torch.equal(out.bool(), ref)
    C = 16
    proj = _randn(BF * (H // 2) * (W // 2), C * 4, seed=5)
    o = torch.empty(BF, C, H, W, dtype=BF16, device="cuda")
    ops.unpatchify(proj, o, BF, C, H, W)
    ref = proj.reshape(1, BF, H // 2, W // 2, C, 2, 2).permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)[0]
    assert torch.equal(o, ref)
