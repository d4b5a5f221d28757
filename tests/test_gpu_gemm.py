"""GPU parity of the tcgen05 GEMM and its fused epilogues against PyTorch fp32 on the same bf16 inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

from _util import assert_close_bf16, block_error_map, report

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from videopainter_b200 import ops as _ops
    return _ops


def _randn(*shape, seed=0, dtype=BF16, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda", dtype=torch.float32) * scale).to(dtype)


def _mk(M, N, K, seed=0):
    a = _randn(M, K, seed=seed)
    w = _randn(N, K, seed=seed + 1, scale=1 / math.sqrt(K))
    b = _randn(N, seed=seed + 2, scale=0.5)
    return a, w, b


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (256, 512, 128), (300, 320, 192), (448, 128, 128),
                                   (1000, 64, 3072), (4096, 3072, 3072), (35552, 3072, 3072)])
def test_gemm_bias(ops, M, N, K):
    a, w, b = _mk(M, N, K)
    out = torch.zeros(M, N, dtype=BF16, device="cuda")
    ops.gemm_bias(a, w, b, out, M, N, K, rows_per_batch=M, out_batch_rows=0, out_row_offset=0, alpha=0.5)
    torch.cuda.synchronize()
    ref = (a.float() @ w.float().t() + b.float()) * 0.5
    try:
        assert_close_bf16(f"gemm_bias {M}x{N}x{K}", out, ref)
    except AssertionError:
        print(block_error_map(out, ref))
        raise


def test_gemm_gelu(ops):
    M, N, K = 777, 1024, 512
    a, w, b = _mk(M, N, K, seed=3)
    out = torch.zeros(M, N, dtype=BF16, device="cuda")
    ops.gemm_gelu(a, w, b, out, M, N, K)
    ref = F.gelu(a.float() @ w.float().t() + b.float(), approximate="tanh")
    assert_close_bf16("gemm_gelu", out, ref)


def test_gemm_gate_residual_inject(ops):
    B, S, St, D, K = 2, 300, 26, 256, 512
    Sv = S - St
    a, w, b = _mk(B * S, D, K, seed=5)
    res = _randn(B, S, D, seed=8)
    gate = _randn(B, 6 * D, seed=9, dtype=torch.float32)
    inj_full = _randn(B, S, D, seed=10)
    inj = inj_full[:, St:]                                  # non-contiguous view, like a branch sample slice
    mask = (torch.rand(B, Sv, device="cuda") > 0.5).to(torch.uint8)
    out = torch.zeros(B, S, D, dtype=BF16, device="cuda")
    ops.gemm_gate_residual(a, w, b, out, B * S, D, K, rows_per_batch=S, out_batch_rows=S, out_row_offset=0, res=res,
                           res_batch_rows=S, res_row_offset=0, gate=gate, gate_video_off=2 * D, gate_text_off=5 * D,
                           text_len=St, inject=inj, inject_batch_stride=inj.stride(0), ldi=inj.stride(1), inject_mask=mask,
                           video_len=Sv)
    y = (a.float() @ w.float().t() + b.float()).view(B, S, D)
    g = torch.cat([gate[:, None, 5 * D:6 * D].expand(B, St, D), gate[:, None, 2 * D:3 * D].expand(B, Sv, D)], dim=1)
    ref = res.float() + g * y
    ref[:, St:] = torch.where(mask[..., None] == 0, ref[:, St:] + inj.float(), ref[:, St:])
    assert_close_bf16("gemm_gate_residual+inject", out, ref)


def test_gemm_row_mapping_patch_embed_style(ops):
    # rows of a compact [B*Sv, K] operand land at x[b, St + s]; the residual is a batch-broadcast table
    B, Sv, St, D, K = 2, 208, 16, 128, 128
    S = St + Sv
    a, w, b = _mk(B * Sv, D, K, seed=11)
    pos = _randn(S, D, seed=12)
    x = torch.zeros(B, S, D, dtype=BF16, device="cuda")
    ops.gemm_gate_residual(a, w, b, x, B * Sv, D, K, rows_per_batch=Sv, out_batch_rows=S, out_row_offset=St, res=pos,
                           res_batch_rows=0, res_row_offset=St)
    ref = (a.float() @ w.float().t() + b.float()).view(B, Sv, D) + pos[St:].float()
    assert_close_bf16("patch-embed mapping", x[:, St:], ref)
    assert (x[:, :St] == 0).all()
    # negative offset drops the text rows (branch_blocks on video rows only)
    a2, w2, b2 = _mk(B * S, D, K, seed=13)
    o = torch.zeros(B, Sv, D, dtype=BF16, device="cuda")
    ops.gemm_bias(a2, w2, b2, o, B * S, D, K, rows_per_batch=S, out_batch_rows=Sv, out_row_offset=-St, alpha=1.0)
    ref2 = (a2.float() @ w2.float().t() + b2.float()).view(B, S, D)[:, St:]
    assert_close_bf16("negative row offset", o, ref2)


def _ref_qk(x, w, b, cos, sin, St):
    # x [B, H, S, 64] fp32 -> LayerNorm(64, eps 1e-6) -> RoPE on rows >= St (EMB:683-692)
    y = F.layer_norm(x, (64,), w.float(), b.float(), 1e-6)
    v = y[:, :, St:]
    xr, xi = v.reshape(*v.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    y[:, :, St:] = v * cos[None, None] + rot * sin[None, None]
    return y


@pytest.mark.parametrize("H,B,S,St", [(2, 2, 224, 16), (48, 1, 700, 226)])
def test_gemm_qkv(ops, H, B, S, St):
    D = H * 64
    Sv = S - St
    M = B * S
    a = _randn(M, D, seed=1)
    w = _randn(3 * D, D, seed=2, scale=1 / math.sqrt(D))
    bias = _randn(3 * D, seed=3, scale=0.5)
    nq = (_randn(64, seed=4) * 0.1 + 1, _randn(64, seed=5) * 0.1)
    nk = (_randn(64, seed=6) * 0.1 + 1, _randn(64, seed=7) * 0.1)
    ang = torch.rand(Sv, 32, device="cuda") * 6.28
    cos = torch.cos(ang).repeat_interleave(2, dim=1).contiguous()
    sin = torch.sin(ang).repeat_interleave(2, dim=1).contiguous()
    mask2 = (torch.rand(M, device="cuda") > 0.5).to(torch.uint8)
    q, k, v, k2, v2 = (torch.zeros(B, H, S, 64, dtype=BF16, device="cuda") for _ in range(5))
    ops.gemm_qkv(a, w, bias, M, D, S, H, 0, q, k, v, nq, nk, 1e-6, (cos, sin), St, k2_out=k2, v2_out=v2, mask2=mask2)
    y = (a.float() @ w.float().t() + bias.float()).view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4).contiguous()
    mk = mask2.view(B, 1, S, 1).float()
    assert_close_bf16("qkv.q", q, _ref_qk(y[0].clone(), *nq, cos, sin, St))
    assert_close_bf16("qkv.k", k, _ref_qk(y[1].clone(), *nk, cos, sin, St))
    assert_close_bf16("qkv.v", v, y[2])
    assert_close_bf16("qkv.k2", k2, _ref_qk(y[1] * mk, *nk, cos, sin, St))
    assert_close_bf16("qkv.v2", v2, y[2] * mk)
    # the compact (cos, sin)-pair table gives the same numbers as the general per-element tables
    pairs = torch.stack([cos[:, 0::2], sin[:, 0::2]], dim=-1).reshape(Sv, 64).contiguous()
    q3, k3, v3, k23, v23 = (torch.zeros(B, H, S, 64, dtype=BF16, device="cuda") for _ in range(5))
    ops.gemm_qkv(a, w, bias, M, D, S, H, 0, q3, k3, v3, nq, nk, 1e-6, (cos, sin, pairs), St, k2_out=k23, v2_out=v23, mask2=mask2)
    assert_close_bf16("qkv.q compact rope", q3, _ref_qk(y[0].clone(), *nq, cos, sin, St))
    assert_close_bf16("qkv.k2 compact rope", k23, _ref_qk(y[1] * mk, *nk, cos, sin, St))
    assert (q3.float() - q.float()).abs().max().item() <= 2 ** -6 and torch.equal(v3, v)
    # a table that does NOT repeat its values pairwise must take the general path and still be exact
    cos_g, sin_g = torch.rand(Sv, 64, device="cuda"), torch.rand(Sv, 64, device="cuda")
    qg, kg, vg = (torch.zeros(B, H, S, 64, dtype=BF16, device="cuda") for _ in range(3))
    ops.gemm_qkv(a, w, bias, M, D, S, H, 0, qg, kg, vg, nq, nk, 1e-6, (cos_g, sin_g), St)
    assert_close_bf16("qkv.q general rope", qg, _ref_qk(y[0].clone(), *nq, cos_g, sin_g, St))
    # K/V-only projection of previous-window states with a per-row scale (AP:2247-2252)
    rs = torch.rand(M, device="cuda") * (torch.rand(M, device="cuda") > 0.3)
    pk, pv = (torch.zeros(B, H, S, 64, dtype=BF16, device="cuda") for _ in range(2))
    ops.gemm_qkv(a, w[D:], bias[D:], M, D, S, H, 1, None, pk, pv, None, nk, 1e-6, (cos, sin), St, row_scale=rs.contiguous())
    sc = rs.view(B, 1, S, 1)
    assert_close_bf16("kv.k", pk, _ref_qk(y[1] * sc, *nk, cos, sin, St))
    assert_close_bf16("kv.v", pv, y[2] * sc)
