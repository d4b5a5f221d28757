"""Torch-tensor front ends of the C ABI (include/vp_b200.h).  PyTorch supplies device memory and the current stream;
every computation happens in libvp_b200.so.  All functions raise if a tensor is not a contiguous CUDA tensor of the
expected dtype — there is no fallback path."""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import check, lib

BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor], dtype=None, name: str = "tensor") -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the B200 path has no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    return t.data_ptr()


def time_sinusoid(timestep: torch.Tensor, dim: int, flip_sin_to_cos: bool, freq_shift: float) -> torch.Tensor:
    B = timestep.shape[0]
    out = torch.empty(B, dim, dtype=torch.float32, device=timestep.device)
    if timestep.dtype == torch.int64:
        ti, tf = _p(timestep, torch.int64, "timestep"), None
    else:
        timestep = timestep.to(torch.float32).contiguous()
        ti, tf = None, _p(timestep, torch.float32, "timestep")
    check(lib().vp_time_sinusoid(ti, tf, _p(out), B, dim, int(flip_sin_to_cos), float(freq_shift), _stream()), "vp_time_sinusoid")
    return out


def gemv(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], act_silu: bool,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, K = x.shape
    N = weight.shape[0]
    if out is None:
        out = torch.empty(B, N, dtype=torch.float32, device=x.device)
    check(lib().vp_gemv(_p(x, torch.float32, "gemv.in"), _p(weight, BF16, "gemv.weight"), _p(bias, BF16, "gemv.bias"),
                        _p(out, torch.float32, "gemv.out"), B, N, K, int(act_silu), _stream()), "vp_gemv")
    return out


def ln_modulate(x: torch.Tensor, x_batch_rows: int, x_row_offset: int, y: torch.Tensor, batch: int, rows_per_batch: int,
                dim: int, gamma, beta, eps: float, mod: Optional[torch.Tensor], offs=(0, 0, 0, 0), text_len: int = 0):
    """offs = (shift_video, scale_video, shift_text, scale_text) element offsets into one batch row of `mod`."""
    check(lib().vp_ln_modulate(_p(x, BF16, "ln.x"), x_batch_rows, x_row_offset, _p(y, BF16, "ln.y"), batch, rows_per_batch, dim,
                               _p(gamma, BF16, "ln.gamma"), _p(beta, BF16, "ln.beta"), float(eps),
                               _p(mod, torch.float32, "ln.mod"), 0 if mod is None else mod.shape[1],
                               offs[0], offs[1], offs[2], offs[3], text_len, _stream()), "vp_ln_modulate")
    return y


def ln_final(x, x_batch_rows, x_row_offset, y, batch, rows_per_batch, dim, g1, b1, g2, b2, eps, mod, shift_off, scale_off):
    check(lib().vp_ln_final(_p(x, BF16, "lnf.x"), x_batch_rows, x_row_offset, _p(y, BF16, "lnf.y"), batch, rows_per_batch, dim,
                            _p(g1, BF16), _p(b1, BF16), _p(g2, BF16), _p(b2, BF16), float(eps),
                            _p(mod, torch.float32, "lnf.mod"), mod.shape[1], shift_off, scale_off, _stream()), "vp_ln_final")
    return y


def gemm_bias(a, w, bias, out, m, n, k, rows_per_batch, out_batch_rows, out_row_offset, alpha=1.0, lda=None, ldw=None, ldo=None):
    check(lib().vp_gemm_bias(_p(a, BF16, "gemm.a"), lda or k, _p(w, BF16, "gemm.w"), ldw or k, _p(bias, BF16, "gemm.bias"),
                             _p(out, BF16, "gemm.out"), ldo or n, m, n, k, rows_per_batch, out_batch_rows, out_row_offset,
                             float(alpha), _stream()), "vp_gemm_bias")
    return out


def gemm_gelu(a, w, bias, out, m, n, k):
    check(lib().vp_gemm_gelu(_p(a, BF16, "gemm.a"), k, _p(w, BF16, "gemm.w"), k, _p(bias, BF16, "gemm.bias"),
                             _p(out, BF16, "gemm.out"), n, m, n, k, _stream()), "vp_gemm_gelu")
    return out


def gemm_gate_residual(a, w, bias, out, m, n, k, rows_per_batch, out_batch_rows, out_row_offset, res, res_batch_rows,
                       res_row_offset, gate=None, gate_video_off=0, gate_text_off=0, text_len=0, inject=None,
                       inject_batch_stride=0, ldi=0, inject_mask=None, video_len=0, lda=None, ldw=None):
    check(lib().vp_gemm_gate_residual(
        _p(a, BF16, "gemm.a"), lda or k, _p(w, BF16, "gemm.w"), ldw or k, _p(bias, BF16, "gemm.bias"), _p(out, BF16, "gemm.out"),
        n, m, n, k, rows_per_batch, out_batch_rows, out_row_offset, _p(res, BF16, "gemm.res"), n, res_batch_rows, res_row_offset,
        _p(gate, torch.float32, "gemm.gate"), 0 if gate is None else gate.shape[1], gate_video_off, gate_text_off, text_len,
        None if inject is None else inject.data_ptr(), inject_batch_stride, ldi,
        _p(inject_mask, torch.uint8, "gemm.inject_mask"), video_len, _stream()), "vp_gemm_gate_residual")
    return out


def gemm_qkv(a, w, bias, m, k, batch_rows, heads, qkv_first, q_out, k_out, v_out, norm_q, norm_k, qk_eps, rope, text_len,
             k2_out=None, v2_out=None, mask2=None, row_scale=None, ldw=None):
    cos, sin = (None, None) if rope is None else rope
    check(lib().vp_gemm_qkv(
        _p(a, BF16, "qkv.a"), k, _p(w, BF16, "qkv.w") if w.is_contiguous() else w.data_ptr(), ldw or k, _p(bias, BF16, "qkv.bias"),
        m, k, batch_rows, heads, qkv_first, _p(q_out, BF16), _p(k_out, BF16), _p(v_out, BF16), _p(k2_out, BF16), _p(v2_out, BF16),
        _p(mask2, torch.uint8, "qkv.mask2"), _p(row_scale, torch.float32, "qkv.row_scale"),
        _p(norm_q[0], BF16) if norm_q else None, _p(norm_q[1], BF16) if norm_q else None, _p(norm_k[0], BF16), _p(norm_k[1], BF16),
        float(qk_eps), _p(cos, torch.float32, "rope.cos"), _p(sin, torch.float32, "rope.sin"), text_len, _stream()), "vp_gemm_qkv")


def attention(q, k0, v0, out, batch, heads, seq_q, kv_len0, softmax_scale, k1=None, v1=None, kv_len1=0, out_scale=1.0,
              accumulate=False, ldo=None):
    check(lib().vp_attention(_p(q, BF16, "attn.q"), _p(k0, BF16, "attn.k"), _p(v0, BF16, "attn.v"), kv_len0, _p(k1, BF16), _p(v1, BF16),
                             kv_len1, _p(out, BF16, "attn.out"), ldo or heads * 64, batch, heads, seq_q, float(softmax_scale),
                             float(out_scale), int(accumulate), _stream()), "vp_attention")
    return out


def patchify(src0, src1, out, bf, h, w, kpad):
    c0 = src0.shape[-3]
    c1 = 0 if src1 is None else src1.shape[-3]
    check(lib().vp_patchify(_p(src0, BF16, "patchify.src0"), c0, _p(src1, BF16, "patchify.src1"), c1, bf, h, w, _p(out, BF16), kpad,
                            _stream()), "vp_patchify")
    return out


def mask_pool(mask, out, bf, h, w):
    check(lib().vp_mask_pool(_p(mask, BF16, "mask"), bf, h, w, _p(out, torch.uint8), _stream()), "vp_mask_pool")
    return out


def unpatchify(proj, out, bf, c, h, w):
    check(lib().vp_unpatchify(_p(proj, BF16, "unpatchify.proj"), bf, c, h, w, _p(out, BF16), _stream()), "vp_unpatchify")
    return out
