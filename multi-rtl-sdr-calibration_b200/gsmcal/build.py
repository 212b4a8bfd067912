"""Builds libgsmcal.so (the C-ABI library of sm_100a kernels) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc")
LIB = os.path.join(CSRC, "libgsmcal.so")
SOURCES = ["gsmcal_api.cu"]
DEPS = ["gsmcal_api.cu", "gsmcal_kernels.cuh", "gsmcal_burst8.cuh", "gsmcal_hostcopy.inc", "gsmcal_demod.cuh", "gsmcal_demod_api.inc", "chn_filter_taps.inc",
        "../../include/gsmcal.h"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = ["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, cwd=CSRC, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
