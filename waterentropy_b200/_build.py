"""Build the CUDA extension in-tree with nvcc for sm_100a (no JIT cache, no torch arch list)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(CSRC, "libwaterentropy_b200.so")
SOURCES = ["we_api.cu"]
HEADERS = ["we_kernels.cuh", "we_numerics.cuh", "we_eig3.cuh", "we_pow64_tables.h", os.path.join("..", "..", "include", "waterentropy_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",  # bit-exact float32/float64 replay of the reference: no FMA contraction
    "-Xcompiler", "-fPIC", "-shared",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        cmd = [find_nvcc(), *NVCC_FLAGS, "-o", SO, *[os.path.join(CSRC, s) for s in SOURCES]]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd, cwd=CSRC)
    return SO
