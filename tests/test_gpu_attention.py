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
