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
# one translation unit per source file, compiled in parallel
UNITS = [os.path.join(CSRC, f) for f in ("ffc_api.cu", "ffc_fft2.cu", "ffc_dft2.cu", "ffc_conv.cu", "ffc_conv_v4.cu", "ffc_conv_v5.cu", "ffc_wgrad_v5.cu", "ffc_conv_small.cu", "ffc_bnact.cu",
                                         "ffc_fu_fused.cu", "ffc_fu2.cu", "ffc_fu4.cu", "ffc_fu2_bwd.cu", "ffc_specnorm.cu", "ffc_fu3.cu", "ffc_fu3_mix.cu", "ffc_glue.cu")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    "-Xptxas=-v",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))


def source_hash() -> str:
    """Content hash of every source the library is built from (independent of file times, which a copy of the working
    tree to another machine does not preserve)."""
    import hashlib
    h = hashlib.sha256()
    for s in sources():
        h.update(os.path.basename(s).encode())
        with open(s, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale(target: str = LIB) -> bool:
    """True when the library is missing or was built from other sources than the ones present (hash stamp beside it)."""
    stamp = target + ".stamp"
    if not os.path.exists(target) or not os.path.exists(stamp):
        return True
    with open(stamp) as f:
        return f.read().strip() != source_hash()


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libffc_b200.so cannot be built on this machine")
    return nvcc


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = find_nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [s for s in sources() if not s.endswith(".cu")] + [os.path.abspath(__file__)]
    hdr_t = max(os.path.getmtime(h) for h in headers)
    procs, objs = [], []
    for unit in UNITS:                                   # translation units compile in parallel
        obj = os.path.join(objdir, os.path.basename(unit)[:-3] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(hdr_t, os.path.getmtime(unit)):
            continue                                     # object is newer than its source and every header
        procs.append((obj, subprocess.Popen([nvcc] + NVCC_FLAGS + ["-c", "-o", obj, unit],
                                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    log = ""
    for obj, pr in procs:
        out, err = pr.communicate()
        log += err
        if pr.returncode != 0:
            if os.path.exists(obj):
                os.remove(obj)
            raise RuntimeError("nvcc failed:\n" + out + err)
    r = subprocess.run([nvcc, "-shared", "-o", LIB + ".tmp"] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    os.replace(LIB + ".tmp", LIB)
    with open(LIB + ".stamp", "w") as f:
        f.write(source_hash() + "\n")
    if verbose:
        sys.stderr.write(log)
    return LIB


def build_emulation(out_path: str) -> str:
    """Host emulation build of the same sources (g++ -DFFC_EMU).  TEST INFRASTRUCTURE ONLY:
    called from tests/, never loaded by the package."""
    srcs_t = max(os.path.getmtime(s) for s in sources())
    if os.path.exists(out_path) and os.path.getmtime(out_path) > srcs_t:
        return out_path
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-DFFC_EMU", "-x", "c++", "-shared", "-fPIC",
           "-ffp-contract=off", "-o", out_path + ".tmp"] + UNITS
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ (emulation build) failed:\n" + r.stdout + r.stderr)
    os.replace(out_path + ".tmp", out_path)
    return out_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
