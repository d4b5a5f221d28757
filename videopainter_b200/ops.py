"""Torch-tensor front ends of the C ABI (include/vp_b200.h).  PyTorch supplies device memory and the current stream;
every computation happens in libvp_b200.so.  All functions raise if a tensor is not a contiguous CUDA tensor of the
expected dtype — there is no fallback path.

The launches go through `torch.ops.vp_b200.*` (csrc/torch_ops.cpp: TORCH_LIBRARY thin wrappers around the C ABI, SURVEY.md
§8b): same arguments as the C prototypes in the same order, tensors instead of pointers, no `stream` (the wrapper passes the
current CUDA stream).  Host-side utilities without a stream (peer-memory allocation) are called through ctypes (_lib.py), and
so is everything when VP_B200_LIB selects a differently-tuned development build of the library."""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _lib
from ._lib import check, lib

BF16 = torch.bfloat16

_torch_ns = None     # torch.ops.vp_b200 once loaded


def _ns():
    global _torch_ns
    if _torch_ns is None:
        from .build import TORCH_LIB
        if not os.path.exists(TORCH_LIB):
            raise RuntimeError(f"{TORCH_LIB} is missing: build it with `python -m videopainter_b200.build` (there is no CPU or "
                               "PyTorch fallback for the denoising path)")
        lib()                                        # fail loudly if libvp_b200.so itself is missing
        torch.ops.load_library(TORCH_LIB)
        _torch_ns = torch.ops.vp_b200
    return _torch_ns


_VIA_CTYPES = "VP_B200_LIB" in os.environ            # development builds are only reachable through ctypes


def _launch(name: str, *args) -> None:
    """One kernel launch: `args` are the C arguments of vp_<name> without the trailing stream — tensors (or None) where
    the prototype has a device pointer, lists of device addresses where it has a pointer array, ints and floats otherwise."""
    if not _VIA_CTYPES:
        getattr(_ns(), name)(*args)
        return
    conv = []
    for a in args:
        if torch.is_tensor(a):
            conv.append(a.data_ptr())
        elif isinstance(a, (list, tuple)):
            conv.append(_ptr_array(a))
        else:
            conv.append(a)
    check(getattr(lib(), "vp_" + name)(*conv, _stream()), "vp_" + name)

# ---- optional per-op device timing (bench.py): CUDA events on the launching stream around every C-ABI call ----------
_prof = None          # None, or dict name -> list[(start_event, end_event, work)]
launch_count = 0      # kernels launched through this module since import (bench.py reports the delta)


def start_profile() -> None:
    global _prof
    _prof = {}


def stop_profile():
    """Returns {name: (launches, total_ms, total_work)}; call after a device synchronise."""
    global _prof
    out = {}
    for name, recs in (_prof or {}).items():
        out[name] = (len(recs), sum(a.elapsed_time(b) for a, b, _ in recs), sum(w for _, _, w in recs))
    _prof = None
    return out


def _timed(name, work_fn=None):
    def deco(fn):
        def wrapper(*args, **kwargs):
            global launch_count
            launch_count += 1
            if _prof is None:
                return fn(*args, **kwargs)
            a = torch.cuda.Event(enable_timing=True)
            b = torch.cuda.Event(enable_timing=True)
            a.record()
            r = fn(*args, **kwargs)
            b.record()
            _prof.setdefault(name, []).append((a, b, work_fn(*args, **kwargs) if work_fn else 0.0))
            return r
        wrapper.__name__ = fn.__name__
        wrapper.__doc__ = fn.__doc__
        return wrapper
    return deco


def _gemm_flops(a, w, bias, out, m, n, k, *args, **kwargs):
    return 2.0 * m * n * k


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _base(t: Optional[torch.Tensor], dtype=BF16) -> Optional[torch.Tensor]:
    """A CUDA tensor that the kernel indexes with its own strides (views into larger buffers): only its base address counts."""
    if t is None:
        return None
    if not t.is_cuda or t.dtype != dtype:
        raise RuntimeError(f"expected a CUDA {dtype} tensor (the B200 path has no CPU fallback)")
    return t


def _p(t: Optional[torch.Tensor], dtype=None, name: str = "tensor") -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the B200 path has no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    return t


@_timed('time_sinusoid')
def time_sinusoid(timestep: torch.Tensor, dim: int, flip_sin_to_cos: bool, freq_shift: float) -> torch.Tensor:
    B = timestep.shape[0]
    out = torch.empty(B, dim, dtype=torch.float32, device=timestep.device)
    if timestep.dtype == torch.int64:
        ti, tf = _p(timestep, torch.int64, "timestep"), None
    else:
        timestep = timestep.to(torch.float32).contiguous()
        ti, tf = None, _p(timestep, torch.float32, "timestep")
    _launch("time_sinusoid", ti, tf, _p(out), B, dim, int(flip_sin_to_cos), float(freq_shift))
    return out


@_timed('gemv')
def gemv(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], act_silu: bool,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, K = x.shape
    N = weight.shape[0]
    if out is None:
        out = torch.empty(B, N, dtype=torch.float32, device=x.device)
    _launch("gemv", _p(x, torch.float32, "gemv.in"), _p(weight, BF16, "gemv.weight"), _p(bias, BF16, "gemv.bias"),
                        _p(out, torch.float32, "gemv.out"), B, N, K, int(act_silu))
    return out


@_timed('ln_modulate', lambda x, xbr, xro, y, batch, rpb, dim, *a, **k: 4.0 * batch * rpb * dim)
def ln_modulate(x: torch.Tensor, x_batch_rows: int, x_row_offset: int, y: torch.Tensor, batch: int, rows_per_batch: int,
                dim: int, gamma, beta, eps: float, mod: Optional[torch.Tensor], offs=(0, 0, 0, 0), text_len: int = 0):
    """offs = (shift_video, scale_video, shift_text, scale_text) element offsets into one batch row of `mod`."""
    _launch("ln_modulate", _p(x, BF16, "ln.x"), x_batch_rows, x_row_offset, _p(y, BF16, "ln.y"), batch, rows_per_batch, dim,
                               _p(gamma, BF16, "ln.gamma"), _p(beta, BF16, "ln.beta"), float(eps),
                               _base(mod, torch.float32), 0 if mod is None else mod.stride(0),
                               offs[0], offs[1], offs[2], offs[3], text_len)
    return y


@_timed('ln_final')
def ln_final(x, x_batch_rows, x_row_offset, y, batch, rows_per_batch, dim, g1, b1, g2, b2, eps, mod, shift_off, scale_off):
    _launch("ln_final", _p(x, BF16, "lnf.x"), x_batch_rows, x_row_offset, _p(y, BF16, "lnf.y"), batch, rows_per_batch, dim,
                            _p(g1, BF16), _p(b1, BF16), _p(g2, BF16), _p(b2, BF16), float(eps),
                            _p(mod, torch.float32, "lnf.mod"), mod.shape[1], shift_off, scale_off)
    return y


@_timed('gemm_bias', _gemm_flops)
def gemm_bias(a, w, bias, out, m, n, k, rows_per_batch, out_batch_rows, out_row_offset, alpha=1.0, lda=None, ldw=None, ldo=None):
    _launch("gemm_bias", _p(a, BF16, "gemm.a"), lda or k, _p(w, BF16, "gemm.w"), ldw or k, _p(bias, BF16, "gemm.bias"),
                             _p(out, BF16, "gemm.out"), ldo or n, m, n, k, rows_per_batch, out_batch_rows, out_row_offset,
                             float(alpha))
    return out


@_timed('gemm_gelu', _gemm_flops)
def gemm_gelu(a, w, bias, out, m, n, k):
    _launch("gemm_gelu", _p(a, BF16, "gemm.a"), k, _p(w, BF16, "gemm.w"), k, _p(bias, BF16, "gemm.bias"),
                             _p(out, BF16, "gemm.out"), n, m, n, k)
    return out


@_timed('gemm_gate_residual', _gemm_flops)
def gemm_gate_residual(a, w, bias, out, m, n, k, rows_per_batch, out_batch_rows, out_row_offset, res, res_batch_rows,
                       res_row_offset, gate=None, gate_video_off=0, gate_text_off=0, text_len=0, inject=None,
                       inject_batch_stride=0, ldi=0, inject_mask=None, video_len=0, lda=None, ldw=None, a_k_chunk=0,
                       a_chunk_stride=0):
    _launch("gemm_gate_residual", 
        _p(a, BF16, "gemm.a"), lda or k, _p(w, BF16, "gemm.w"), ldw or k, _p(bias, BF16, "gemm.bias"), _p(out, BF16, "gemm.out"),
        n, m, n, k, rows_per_batch, out_batch_rows, out_row_offset, _p(res, BF16, "gemm.res"), n, res_batch_rows, res_row_offset,
        _base(gate, torch.float32), 0 if gate is None else gate.stride(0), gate_video_off, gate_text_off, text_len,
        _base(inject), inject_batch_stride, ldi,
        _p(inject_mask, torch.uint8, "gemm.inject_mask"), video_len, a_k_chunk, a_chunk_stride)
    return out


@_timed('gemm_qkv', lambda a, w, bias, m, k, batch_rows, heads, qkv_first, *r, **kw: 2.0 * m * k * (3 - qkv_first) * heads * 64)
def gemm_qkv(a, w, bias, m, k, batch_rows, heads, qkv_first, q_out, k_out, v_out, norm_q, norm_k, qk_eps, rope, text_len,
             k2_out=None, v2_out=None, mask2=None, row_scale=None, ldw=None, heads_per_dest=None, dest_stride=0):
    """heads_per_dest / dest_stride: Ulysses send layout (include/vp_b200.h); q_out .. v2_out may then be views into one send
    buffer, so only their base addresses are taken."""
    cos, sin, cs = _rope3(rope)
    _launch("gemm_qkv", 
        _p(a, BF16, "qkv.a"), k, _base(w), ldw or k, _p(bias, BF16, "qkv.bias"),
        m, k, batch_rows, heads, qkv_first, _base(q_out), _base(k_out), _base(v_out), _base(k2_out), _base(v2_out),
        _p(mask2, torch.uint8, "qkv.mask2"), _p(row_scale, torch.float32, "qkv.row_scale"),
        _p(norm_q[0], BF16) if norm_q else None, _p(norm_q[1], BF16) if norm_q else None, _p(norm_k[0], BF16), _p(norm_k[1], BF16),
        float(qk_eps), _base(cos, torch.float32), _base(sin, torch.float32), _base(cs, torch.float32), text_len,
        heads_per_dest or heads, dest_stride)


def _rope3(rope):
    """rope = None | (cos, sin) | (cos, sin, compact [Sv, 32, 2] pairs or None)"""
    if rope is None:
        return None, None, None
    if len(rope) == 2:
        return rope[0], rope[1], None
    return rope


def _ptr_array(ptrs):
    import ctypes as C
    return (C.c_void_p * len(ptrs))(*ptrs)


@_timed('gemm_qkv', lambda a, w, bias, m, k, heads, qkv_first, *r, **kw: 2.0 * m * k * (3 - qkv_first) * heads * 64)
def gemm_qkv_peer(a, w, bias, m, k, heads, qkv_first, q_out, k_out, v_out, norm_q, norm_k, qk_eps, rope, text_len, peer_ptrs,
                  local_base, seq_total, row_offset, k2_out=None, v2_out=None, mask2=None, row_scale=None, ldw=None):
    """QKV GEMM whose epilogue stores each head into its destination rank's attention buffer over peer memory
    (include/vp_b200.h vp_gemm_qkv_peer).  peer_ptrs: data pointers of every rank's buffer, by rank."""
    cos, sin, cs = _rope3(rope)
    _launch("gemm_qkv_peer", 
        _p(a, BF16, "qkv.a"), k, _base(w), ldw or k, _p(bias, BF16, "qkv.bias"),
        m, k, heads, qkv_first, _base(q_out), _base(k_out), _base(v_out), _base(k2_out), _base(v2_out),
        _p(mask2, torch.uint8, "qkv.mask2"), _p(row_scale, torch.float32, "qkv.row_scale"),
        _p(norm_q[0], BF16) if norm_q else None, _p(norm_q[1], BF16) if norm_q else None, _p(norm_k[0], BF16), _p(norm_k[1], BF16),
        float(qk_eps), _base(cos, torch.float32), _base(sin, torch.float32), _base(cs, torch.float32), text_len,
        list(peer_ptrs), len(peer_ptrs),
        _base(local_base), seq_total, row_offset)


@_timed('attention', lambda q, k0, v0, peer_ptrs, my_rank, ldo, heads, seq_q, kv_len0, scale, k1=None, v1=None, kv_len1=0, **kw: 4.0 * heads * seq_q * (kv_len0 + kv_len1) * 64)
def attention_peer(q, k0, v0, peer_ptrs, my_rank, ldo, heads, seq_q, kv_len0, softmax_scale, k1=None, v1=None, kv_len1=0,
                   out_scale=1.0):
    _launch("attention_peer", _p(q, BF16, "attn.q"), _p(k0, BF16, "attn.k"), _p(v0, BF16, "attn.v"), kv_len0, _p(k1, BF16),
                                  _p(v1, BF16), kv_len1, list(peer_ptrs), len(peer_ptrs), my_rank, ldo, heads, seq_q,
                                  float(softmax_scale), float(out_scale))


@_timed('peer_scatter')
def peer_scatter(src, peer_ptrs, my_rank, bytes_per_peer):
    _launch("peer_scatter", _p(src, name="scatter.src"), list(peer_ptrs), len(peer_ptrs), my_rank, bytes_per_peer)


@_timed('peer_barrier')
def peer_barrier(flag_ptrs, my_rank, epoch):
    _launch("peer_barrier", list(flag_ptrs), len(flag_ptrs), my_rank, epoch & 0xffffffff)


@_timed('a2a_unpack', lambda src, dsts, peers, heads_local, rows_per_peer: 4.0 * len(dsts) * peers * heads_local * rows_per_peer * 64)
def a2a_unpack_heads(src, dsts, peers, heads_local, rows_per_peer):
    """src [peers][len(dsts)][heads_local][rows_per_peer][64] -> dsts[i] [heads_local][peers * rows_per_peer][64]."""
    ptrs = [_p(d, BF16, "unpack.dst") for d in dsts] + [None] * (5 - len(dsts))
    _launch("a2a_unpack_heads", _p(src, BF16, "unpack.src"), *ptrs, len(dsts), peers, heads_local, rows_per_peer)


@_timed('attention', lambda q, k0, v0, out, batch, heads, seq_q, kv_len0, scale, k1=None, v1=None, kv_len1=0, **kw: 4.0 * batch * heads * seq_q * (kv_len0 + kv_len1) * 64)
def attention(q, k0, v0, out, batch, heads, seq_q, kv_len0, softmax_scale, k1=None, v1=None, kv_len1=0, out_scale=1.0,
              accumulate=False, ldo=None):
    _launch("attention", _p(q, BF16, "attn.q"), _p(k0, BF16, "attn.k"), _p(v0, BF16, "attn.v"), kv_len0, _p(k1, BF16), _p(v1, BF16),
                             kv_len1, _p(out, BF16, "attn.out"), ldo or heads * 64, batch, heads, seq_q, float(softmax_scale),
                             float(out_scale), int(accumulate))
    return out


@_timed('patchify')
def patchify(src0, src1, out, bf, h, w, kpad):
    c0 = src0.shape[-3]
    c1 = 0 if src1 is None else src1.shape[-3]
    _launch("patchify", _p(src0, BF16, "patchify.src0"), c0, _p(src1, BF16, "patchify.src1"), c1, bf, h, w, _p(out, BF16), kpad)
    return out


@_timed('mask_pool')
def mask_pool(mask, out, bf, h, w):
    _launch("mask_pool", _p(mask, BF16, "mask"), bf, h, w, _p(out, torch.uint8))
    return out


@_timed('unpatchify')
def unpatchify(proj, out, bf, c, h, w):
    _launch("unpatchify", _p(proj, BF16, "unpatchify.proj"), bf, c, h, w, _p(out, BF16))
    return out


# ---- peer-visible device memory (CUDA IPC) ----------------------------------------------------------------------------
class PeerBuffer:
    """`nbytes` of zero-filled device memory allocated by the library (cudaMalloc, so that it can be exported through CUDA
    IPC) and viewed as a torch uint8 tensor without copying."""

    def __init__(self, nbytes: int, device):
        import ctypes as C
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        with torch.cuda.device(device):
            check(lib().vp_peer_alloc(nbytes, C.byref(ptr), handle), "vp_peer_alloc")
        self.ptr, self.nbytes, self.handle = ptr.value, nbytes, bytes(handle)
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2,
                                         "strides": None}
        self.tensor = torch.as_tensor(self, device=device)
        assert self.tensor.data_ptr() == self.ptr

    def free(self):
        if self.ptr:
            self.tensor = None
            check(lib().vp_peer_free(self.ptr), "vp_peer_free")
            self.ptr = None


def peer_close(ptr: int) -> None:
    """Unmap a peer buffer opened with peer_open."""
    check(lib().vp_peer_close(ptr), "vp_peer_close")


def peer_set_timeout_ms(ms: int) -> None:
    """Bound of the device-side peer barrier's wait (0 = wait for ever).  A barrier that runs out of time adds one to word 8
    of the rank's flag buffer and lets the stream continue; it never traps."""
    check(lib().vp_peer_set_timeout_ms(int(ms)), "vp_peer_set_timeout_ms")


def peer_open(handle: bytes, device) -> int:
    """Map a peer rank's PeerBuffer (its 64-byte IPC handle) into this process; returns the device pointer."""
    import ctypes as C
    ptr = C.c_void_p()
    buf = (C.c_ubyte * 64).from_buffer_copy(handle)
    with torch.cuda.device(device):
        check(lib().vp_peer_open(buf, C.byref(ptr)), "vp_peer_open")
    return ptr.value
