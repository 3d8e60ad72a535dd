"""Build libmppi_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m quadrotor_manipulator_mppi_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.environ.get("MPPI_B200_LIB") or os.path.join(PKG, "libmppi_b200.so")     # override: a prebuilt library
SOURCES = [os.path.join(CSRC, "mppi_b200.cu")]
HEADERS = [os.path.join(CSRC, "mppi_device.cuh"), os.path.join(CSRC, "mppi_kernels.cuh"),
           os.path.join(CSRC, "mppi_vec.cuh"), os.path.join(CSRC, "fk_tables_gen.cuh"),
           os.path.join(ROOT, "include", "mppi_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libmppi_b200.so cannot be built (there is no CPU fallback)")


def nvcc_command(out: str = LIB, extra=()) -> list:
    return [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "-shared", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
            "-I", os.path.join(ROOT, "include"), "-I", CSRC, *extra, "-o", out, *SOURCES]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        cmd = nvcc_command(extra=("-Xptxas", "-v") if verbose else ())
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
