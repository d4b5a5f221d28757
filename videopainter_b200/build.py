"""Builds csrc/*.cu into csrc/libvp_b200.so for sm_100a with nvcc (in-tree; the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libvp_b200.so")
TORCH_LIB = os.path.join(CSRC, "libvp_b200_torch.so")      # TORCH_LIBRARY(vp_b200) wrappers around the C ABI (torch_ops.cpp)
SOURCES = ["capi.cu", "gemm.cu", "attention.cu", "attention_v1.cu", "attention_v4.cu", "elementwise.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--threads", "4",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(os.path.dirname(CSRC)), "include", "vp_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name: str, defines) -> str:
    """Development aid: a differently-tuned copy of the library (csrc/libvp_b200_<name>.so), selected at run time with
    the VP_B200_LIB environment variable (see _lib.py)."""
    out = os.path.join(CSRC, f"libvp_b200_{name}.so")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + SOURCES + ["-o", out]
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return out


def build_torch_ops(force: bool = False) -> str:
    """csrc/torch_ops.cpp -> csrc/libvp_b200_torch.so: host-only C++ (g++), links libtorch and, by $ORIGIN rpath, libvp_b200.so."""
    src = os.path.join(CSRC, "torch_ops.cpp")
    hdr = os.path.join(os.path.dirname(os.path.dirname(CSRC)), "include", "vp_b200.h")
    if not force and os.path.exists(TORCH_LIB) and os.path.getmtime(TORCH_LIB) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return TORCH_LIB
    import torch
    from torch.utils import cpp_extension as ce
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    cmd += [f"-I{p}" for p in ce.include_paths()] + [f"-I{cuda_inc}", "torch_ops.cpp", "-o", TORCH_LIB, f"-L{CSRC}",
                                                     "-l:libvp_b200.so", f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda",
                                                     "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"]
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed on torch_ops.cpp:\n" + r.stdout + r.stderr)
    return TORCH_LIB


def build(force: bool = False, verbose: bool = False) -> str:
    lib = _build_cuda(force, verbose)
    build_torch_ops(force)
    return lib


def _build_cuda(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libvp_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + SOURCES + ["-o", LIB]
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
