#!/usr/bin/env python
"""Build an experimental variant of libmppi_b200.so next to the shipped one (for A/B timing on the GPU box).

    python tools/build_variant.py NAME -DMPPI_WB_NB=2 [-DMPPI_WB_MINB=5 ...] [--units 3:0,3:1]

Recompiles the named (model:part) translation units with the extra flags, links them with the shipped objects in
csrc/_obj/ (run the normal build first) into variants/libmppi_b200_NAME.so and prints the register / stack report of
the whole-body rollout kernels.  Select it at run time with MPPI_B200_LIB=variants/libmppi_b200_NAME.so.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quadrotor_manipulator_mppi_b200 import build as B  # noqa: E402


def main():
    name, flags, units = sys.argv[1], [], [(3, 0)]
    for a in sys.argv[2:]:
        if a.startswith("--units"):
            units = [tuple(int(x) for x in u.split(":")) for u in a.split("=", 1)[1].split(",")]
        else:
            flags.append(a)
    B.build()
    vdir = os.path.join(ROOT, "variants")
    os.makedirs(vdir, exist_ok=True)
    objs = {os.path.basename(o): o for o, _ in B.compile_commands()}
    procs = []
    for m, p in units:
        obj = os.path.join(vdir, f"{name}_unit_m{m}_p{p}.o")
        cmd = [B._nvcc(), *B._flags(("-Xptxas", "-v", *flags)), f"-DMPPI_UNIT_MODEL={m}", f"-DMPPI_UNIT_PART={p}", "-c", B.UNIT, "-o", obj]
        procs.append((obj, m, p, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for obj, m, p, pr in procs:
        out = pr.communicate()[0]
        if pr.returncode:
            raise SystemExit(out)
        objs[f"unit_m{m}_p{p}.o"] = obj
        # ptxas -v: "Compiling entry function '<mangled>'", then "Function properties", stack line, "Used N registers"
        for blk in re.split(r"(?=Compiling entry function)", out):
            mm = re.search(r"Compiling entry function '(\S+)'", blk)
            if not mm:
                continue
            fn = subprocess.run(["c++filt", mm.group(1)], capture_output=True, text=True).stdout.strip()
            if re.search(r"rollout_cost_kernel<\d, 0, \w+, false", fn):
                regs = re.search(r"Used (\d+) registers", blk)
                stack = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", blk)
                print(fn[6:66], "regs", regs.group(1) if regs else "?", "stack/spill", stack.groups() if stack else "?")
    lib = os.path.join(vdir, f"libmppi_b200_{name}.so")
    r = subprocess.run(B.link_command(list(objs.values()), lib), capture_output=True, text=True)
    if r.returncode:
        raise SystemExit(r.stdout + r.stderr)
    print(lib)


if __name__ == "__main__":
    main()
