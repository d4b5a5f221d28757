"""ctypes binding of the C ABI in include/vp_b200.h.  There is no fallback: if the shared library is missing the
import of any op fails loudly."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB as _DEFAULT_LIB

# VP_B200_LIB selects a differently-tuned build of the SAME kernels (development aid); there is still no fallback.
LIB = os.environ.get("VP_B200_LIB", _DEFAULT_LIB)

_c_void_p, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> argtypes, exactly the prototypes of include/vp_b200.h
SIGNATURES = {
    "vp_time_sinusoid": [_c_void_p, _c_void_p, _c_void_p, _i, _i, _i, _f, _c_void_p],
    "vp_gemv": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _i, _i, _i, _i, _c_void_p],
    "vp_ln_modulate": [_c_void_p, _ll, _i, _c_void_p, _i, _i, _i, _c_void_p, _c_void_p, _f, _c_void_p, _ll, _i, _i, _i, _i,
                       _i, _c_void_p],
    "vp_ln_final": [_c_void_p, _ll, _i, _c_void_p, _i, _i, _i, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _f, _c_void_p,
                    _ll, _i, _i, _c_void_p],
    "vp_gemm_bias": [_c_void_p, _ll, _c_void_p, _ll, _c_void_p, _c_void_p, _i, _i, _i, _i, _i, _ll, _i, _f, _c_void_p],
    "vp_gemm_gelu": [_c_void_p, _ll, _c_void_p, _ll, _c_void_p, _c_void_p, _i, _i, _i, _i, _c_void_p],
    "vp_gemm_gate_residual": [_c_void_p, _ll, _c_void_p, _ll, _c_void_p, _c_void_p, _i, _i, _i, _i, _i, _ll, _i,
                              _c_void_p, _i, _ll, _i, _c_void_p, _ll, _i, _i, _i, _c_void_p, _ll, _i, _c_void_p, _i,
                              _i, _ll, _c_void_p],
    "vp_gemm_qkv": [_c_void_p, _ll, _c_void_p, _ll, _c_void_p, _i, _i, _i, _i, _i, _c_void_p, _c_void_p, _c_void_p,
                    _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _f,
                    _c_void_p, _c_void_p, _c_void_p, _i, _i, _ll, _c_void_p],
    "vp_gemm_qkv_peer": [_c_void_p, _ll, _c_void_p, _ll, _c_void_p, _i, _i, _i, _i, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                         _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _f, _c_void_p, _c_void_p,
                         _c_void_p, _i, _c_void_p, _i, _c_void_p, _i, _i, _c_void_p],
    "vp_attention_peer": [_c_void_p, _c_void_p, _c_void_p, _i, _c_void_p, _c_void_p, _i, _c_void_p, _i, _i, _i, _i, _i, _f, _f,
                          _c_void_p],
    "vp_peer_barrier": [_c_void_p, _i, _i, C.c_uint, _c_void_p],
    "vp_step_end": [_c_void_p, _f, _c_void_p, _c_void_p, _c_void_p, _f, _f, _f, _f, _f, _f, _f, _i, _c_void_p, _c_void_p, _c_void_p,
                    _c_void_p, _c_void_p, _c_void_p, _i, _ll, _f, _f, _i, _i, _ll, _c_void_p],
    "vp_peer_scatter": [_c_void_p, _c_void_p, _i, _i, _ll, _c_void_p],
    "vp_peer_alloc": [_ll, _c_void_p, _c_void_p],
    "vp_peer_open": [_c_void_p, _c_void_p],
    "vp_peer_close": [_c_void_p],
    "vp_peer_free": [_c_void_p],
    "vp_peer_set_timeout_ms": [_ll],
    "vp_a2a_unpack_heads": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _i, _i, _i, _i, _c_void_p],
    "vp_attention": [_c_void_p, _c_void_p, _c_void_p, _i, _c_void_p, _c_void_p, _i, _c_void_p, _i, _i, _i, _i, _f, _f, _i,
                     _c_void_p],
    "vp_patchify": [_c_void_p, _i, _c_void_p, _i, _i, _i, _i, _c_void_p, _i, _c_void_p],
    "vp_mask_pool": [_c_void_p, _i, _i, _i, _c_void_p, _c_void_p],
    "vp_unpatchify": [_c_void_p, _i, _i, _i, _i, _c_void_p, _c_void_p],
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(
                f"{LIB} is missing: build it with `python -m videopainter_b200.build` (there is no CPU or PyTorch "
                "fallback for the denoising path)")
        L = C.CDLL(LIB)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _i
        L.vp_version.restype = _i
        L.vp_last_error.restype = C.c_char_p
        L.vp_last_cuda_error.restype = _i
        _lib = L
    return _lib


class VpError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        L = lib()
        raise VpError(f"{what} failed: rc={rc} ({L.vp_last_error().decode()}; cudaError={L.vp_last_cuda_error()})")
