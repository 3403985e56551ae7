"""Build ``lib/libr2ik.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m reachy2_symbolic_ik_b200.build [--force] [-v]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib", "libr2ik.so")
SOURCES = ["r2ik_kernels.cu", "r2ik_pipeline.cu"]
HEADERS = ["r2ik_math.cuh", "r2ik_device.cuh", "r2ik_device_f32.cuh", "r2ik_control.cuh", "r2ik_cont_codes.cuh", "r2ik_discrete_compact.cuh", "r2ik_host.h", os.path.join("..", "..", "include", "r2ik.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    # FP64 parity: no fast-math; FMA contraction stays on (rounding-level effect only)
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libr2ik.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    extra = os.environ.get("R2IK_NVCC_EXTRA", "").split()   # tuning experiments only (e.g. -DR2IK_K1_MINBLOCKS=5)
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
