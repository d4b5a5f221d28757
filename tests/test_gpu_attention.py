"""GPU parity of the tcgen05 flash-attention kernel against explicit fp32 softmax attention."""
import math

import pytest
import torch

from _util import assert_close_bf16

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from videopainter_b200 import ops as _ops
    return _ops


def _randn(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda", dtype=torch.float32) * scale).to(BF16)


def _ref(q, k, v):
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) / 8.0
    return torch.matmul(torch.softmax(s, dim=-1), v.float())            # [B, H, S, 64]


def _to_out_layout(o):   # [B, H, S, 64] -> [B, S, H*64]
    B, H, S, d = o.shape
    return o.transpose(1, 2).reshape(B, S, H * d)


@pytest.mark.parametrize("B,H,Sq,Skv,scale", [(1, 1, 128, 128, 1.0), (1, 1, 256, 256, 1.0), (1, 2, 256, 384, 1.0),
                                               (2, 2, 224, 224, 1.0), (1, 3, 300, 1000, 3.0), (2, 4, 2048, 2048, 2.0),
                                               (1, 48, 1111, 1111, 1.0)])
def test_attention_single_segment(ops, B, H, Sq, Skv, scale):
    q = _randn(B, H, Sq, 64, seed=1, scale=scale)
    k = _randn(B, H, Skv, 64, seed=2, scale=scale)
    v = _randn(B, H, Skv, 64, seed=3)
    out = torch.zeros(B, Sq, H * 64, dtype=BF16, device="cuda")
    ops.attention(q, k, v, out, B, H, Sq, Skv, 0.125)
    torch.cuda.synchronize()
    assert_close_bf16(f"attention {B}x{H}x{Sq}x{Skv}", out, _to_out_layout(_ref(q, k, v)), cos_min=0.9998, rel_max=3e-2)


def test_attention_two_segments_and_blend(ops):
    B, H, S = 2, 2, 300
    q, k, v = _randn(B, H, S, 64, seed=1, scale=2.0), _randn(B, H, S, 64, seed=2, scale=2.0), _randn(B, H, S, 64, seed=3)
    k2, v2 = _randn(B, H, S, 64, seed=4, scale=2.0), _randn(B, H, S, 64, seed=5)
    out = torch.zeros(B, S, H * 64, dtype=BF16, device="cuda")
    ops.attention(q, k, v, out, B, H, S, S, 0.125, k1=k2, v1=v2, kv_len1=S)
    ref = _ref(q, torch.cat([k, k2], dim=2), torch.cat([v, v2], dim=2))
    assert_close_bf16("attention two segments", out, _to_out_layout(ref), cos_min=0.9998, rel_max=3e-2)
    # previous-window blend: (1 - w) * attn(q, k, v) + w * attn(q, k2, v2)   (AP:2176-2189)
    w = 0.3
    ops.attention(q, k, v, out, B, H, S, S, 0.125, out_scale=1 - w)
    ops.attention(q, k2, v2, out, B, H, S, S, 0.125, out_scale=w, accumulate=True)
    ref = (1 - w) * _ref(q, k, v) + w * _ref(q, k2, v2)
    assert_close_bf16("attention blend", out, _to_out_layout(ref), cos_min=0.9998, rel_max=3e-2)


def test_attention_large_scores_rescale_path(ops):
    # growing keys force the running maximum to move by more than 2^8 several times (lazy rescale path)
    B, H, S = 1, 2, 640
    q = _randn(B, H, S, 64, seed=1, scale=4.0)
    k = _randn(B, H, S, 64, seed=2, scale=1.0)
    k = (k.float() * torch.linspace(0.2, 6.0, S, device="cuda")[None, None, :, None]).to(BF16)
    v = _randn(B, H, S, 64, seed=3)
    out = torch.zeros(B, S, H * 64, dtype=BF16, device="cuda")
    ops.attention(q, k, v, out, B, H, S, S, 0.125)
    assert_close_bf16("attention rescale", out, _to_out_layout(_ref(q, k, v)), cos_min=0.9995, rel_max=5e-2)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json full size (S = 17 776 tokens, 48 heads x 64, CFG batch 2)
# ---------------------------------------------------------------------------------------------------------------------
S_FULL = 17776


def test_attention_full_sequence_against_fp32(ops):
    """Two heads of the production sequence length against explicit fp32 softmax attention (1.3 GB of scores per head)."""
    B, H = 1, 2
    q, k, v = _randn(B, H, S_FULL, 64, seed=11), _randn(B, H, S_FULL, 64, seed=12), _randn(B, H, S_FULL, 64, seed=13)
    out = torch.zeros(B, S_FULL, H * 64, dtype=BF16, device="cuda")
    ops.attention(q, k, v, out, B, H, S_FULL, S_FULL, 0.125)
    assert_close_bf16("attention S=17776", out, _to_out_layout(_ref(q, k, v)), cos_min=0.9998, rel_max=3e-2)
    # doubled K/V of the ID-resample processor (L_kv = 35 552), one head
    q1, k1, v1 = q[:, :1].contiguous(), k[:, :1].contiguous(), v[:, :1].contiguous()
    k2, v2 = _randn(1, 1, S_FULL, 64, seed=14), _randn(1, 1, S_FULL, 64, seed=15)
    out1 = torch.zeros(1, S_FULL, 64, dtype=BF16, device="cuda")
    ops.attention(q1, k1, v1, out1, 1, 1, S_FULL, S_FULL, 0.125, k1=k2, v1=v2, kv_len1=S_FULL)
    ref = _ref(q1, torch.cat([k1, k2], dim=2), torch.cat([v1, v2], dim=2))
    assert_close_bf16("attention S=17776, L_kv=35552", out1, _to_out_layout(ref), cos_min=0.9998, rel_max=3e-2)


def test_attention_full_size_properties(ops):
    """Size-independent properties on the whole production problem (B = 2, 48 heads, S = 17 776):
    softmax rows sum to one (V = 1 gives exactly 1), and the output is linear in V."""
    B, H = 2, 48
    q, k = _randn(B, H, S_FULL, 64, seed=21), _randn(B, H, S_FULL, 64, seed=22)
    ones = torch.ones(B, H, S_FULL, 64, dtype=BF16, device="cuda")
    out = torch.zeros(B, S_FULL, H * 64, dtype=BF16, device="cuda")
    ops.attention(q, k, ones, out, B, H, S_FULL, S_FULL, 0.125)
    assert (out.float() - 1.0).abs().max().item() <= 2 ** -7, "softmax rows do not sum to one"
    v1, v2 = _randn(B, H, S_FULL, 64, seed=23), _randn(B, H, S_FULL, 64, seed=24)
    o1, o2, o12 = (torch.zeros_like(out) for _ in range(3))
    ops.attention(q, k, v1, o1, B, H, S_FULL, S_FULL, 0.125)
    ops.attention(q, k, v2, o2, B, H, S_FULL, S_FULL, 0.125)
    ops.attention(q, k, (v1.float() + 2.0 * v2.float()).to(BF16), o12, B, H, S_FULL, S_FULL, 0.125)
    lin = o1.float() + 2.0 * o2.float()
    assert_close_bf16("attention linearity in V (full size)", o12, lin, cos_min=0.9995, rel_max=5e-2)
