"""Compile librcw_b200.so for sm_100a with nvcc (csrc/Makefile).  No GPU is needed to build."""
from __future__ import annotations

import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB = os.path.join(_PKG, "lib", "librcw_b200.so")
SOURCES = ("rcw_kernels.cu", "rcw_capi.cu", "rcw_internal.h", "rcw_topview.cuh", "Makefile")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES]
    deps.append(os.path.join(os.path.dirname(_PKG), "include", "rcw_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or is_stale():
        r = subprocess.run(["make", "-C", CSRC, "-B"], capture_output=True, text=True)
        if verbose or r.returncode:
            print(r.stdout)
            print(r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc build of librcw_b200.so failed")
    return LIB
