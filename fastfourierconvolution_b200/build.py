"""Builds libffc_b200.so (hand-written sm_100a CUDA, plain C ABI) in-tree with nvcc.

    python -m fastfourierconvolution_b200.build            # build if stale
    python -m fastfourierconvolution_b200.build --force

nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container; the
resulting .so is git-ignored but travels to the GPU box with the working tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libffc_b200.so")
UNIT = os.path.join(CSRC, "ffc_lib.cu")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas=-v",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))


def is_stale(target: str = LIB) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources() + [os.path.abspath(__file__)])


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libffc_b200.so cannot be built on this machine")
    return nvcc


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + ["-o", LIB + ".tmp", UNIT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


def build_emulation(out_path: str) -> str:
    """Host emulation build of the same sources (g++ -DFFC_EMU).  TEST INFRASTRUCTURE ONLY:
    called from tests/, never loaded by the package."""
    srcs_t = max(os.path.getmtime(s) for s in sources())
    if os.path.exists(out_path) and os.path.getmtime(out_path) > srcs_t:
        return out_path
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-DFFC_EMU", "-x", "c++", "-shared", "-fPIC",
           "-ffp-contract=off", "-o", out_path + ".tmp", UNIT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ (emulation build) failed:\n" + r.stdout + r.stderr)
    os.replace(out_path + ".tmp", out_path)
    return out_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
