"""Build libmppi_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m quadrotor_manipulator_mppi_b200.build [--force] [--verbose]

The kernels are ~110 template instantiations; they are compiled as independent translation units (csrc/mppi_model_unit.cu
once per (model, part)) in parallel and linked with the C-ABI unit (csrc/mppi_b200.cu).  A content hash of every source
is stored next to the library: a library whose hash does not match the sources on disk is rebuilt, never loaded.
"""
from __future__ import annotations

import concurrent.futures
import glob
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.environ.get("MPPI_B200_LIB") or os.path.join(PKG, "libmppi_b200.so")     # override: a prebuilt library
HASH = LIB + ".srchash"
MAIN = os.path.join(CSRC, "mppi_b200.cu")
UNIT = os.path.join(CSRC, "mppi_model_unit.cu")
UNITS = [(m, p) for m in range(4) for p in range(2)]       # (model, part)


def sources() -> list:
    """Everything the library is built from (any edit to any of them invalidates a built .so)."""
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  glob.glob(os.path.join(ROOT, "include", "*.h")))


def source_hash() -> str:
    h = hashlib.sha256()
    for p in sources():
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libmppi_b200.so cannot be built (there is no CPU fallback)")


def _flags(extra=()) -> list:
    return ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
            "-ccbin", "/usr/bin/g++", "-I", os.path.join(ROOT, "include"), "-I", CSRC, *extra]


def compile_commands(extra=()) -> list:
    """[(object path, nvcc command)] for every translation unit."""
    cmds = [(os.path.join(OBJ, "mppi_b200.o"), [_nvcc(), *_flags(extra), "-c", MAIN, "-o", os.path.join(OBJ, "mppi_b200.o")])]
    for m, p in UNITS:
        obj = os.path.join(OBJ, f"unit_m{m}_p{p}.o")
        cmds.append((obj, [_nvcc(), *_flags(extra), f"-DMPPI_UNIT_MODEL={m}", f"-DMPPI_UNIT_PART={p}", "-c", UNIT, "-o", obj]))
    return cmds


def link_command(objs, out: str = LIB) -> list:
    return [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-ccbin", "/usr/bin/g++", "-o", out, *objs, "-ldl"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    if os.environ.get("MPPI_B200_LIB"):
        return False                       # a prebuilt library was named explicitly
    try:
        with open(HASH) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    if not (force or needs_build()):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    cmds = compile_commands(extra=("-Xptxas", "-v") if verbose else ())

    def run(item):
        obj, cmd = item
        r = subprocess.run(cmd, capture_output=True, text=True)
        return obj, cmd, r
    logs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 4)) as ex:
        for obj, cmd, r in ex.map(run, cmds):
            if r.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
            logs.append(r.stdout + r.stderr)
    cmd = link_command([o for o, _ in cmds])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    with open(HASH, "w") as f:
        f.write(source_hash() + "\n")
    if verbose:
        print("".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
