#!/usr/bin/env python
"""Executed FP32 FLOPs of a kernel from an ncu report's SASS page (run here, no GPU needed).

    python tools/ncu_flops.py gpurun_out/r02a_r10_full.ncu-rep [kernel-regex]

ncu's `smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on` counters only see the SCALAR FP32 instructions on
sm_100: the packed FFMA2 / FMUL2 / FADD2 this build leans on are not in them (the scalar counters read 202 FLOP per
whole-body rollout-step where the instruction mix says ~800).  So the executed FLOPs are taken from the per-instruction
"Thread Instructions Executed" column of the source page: FFMA 2, FMUL / FADD 1, FFMA2 4, FMUL2 / FADD2 2 per thread
(predicated-on threads only).  Also reports the FMA-pipe cycle share of the Philox multiplies (IMAD.WIDE / IMAD.HI at
4 cycles, IMAD at 2, FP32 scalar 1, packed 2 -- tools/probe_pipes.cu).
"""
import csv
import io
import json
import re
import subprocess
import sys

FLOP = {"FFMA": 2, "FMUL": 1, "FADD": 1, "FFMA2": 4, "FMUL2": 2, "FADD2": 2}
PIPE_CYC = {"FFMA": 1, "FMUL": 1, "FADD": 1, "FFMA2": 2, "FMUL2": 2, "FADD2": 2, "IMAD.WIDE": 4, "IMAD.HI": 4, "IMAD": 2,
            "FMNMX": 1, "FMNMX3": 1, "FSEL": 1, "FSETP": 1}      # FMA-pipe residents (approximate for the last four)


def kernels(rep, regex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = {"name": next(csv.reader([line]))[1], "lines": []}
            blocks.append(cur)
        elif cur is not None:
            cur["lines"].append(line)
    return blocks


def analyse(block):
    rows = list(csv.reader(io.StringIO("\n".join(block["lines"]))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    flops = 0
    thread_inst = 0
    warp_inst = 0
    pipe = {}
    mix = {}
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        src = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
        op = src.split()[0] if src else ""
        base = op.split(".")[0]
        key = ".".join(op.split(".")[:2]) if base == "IMAD" and len(op.split(".")) > 1 and op.split(".")[1] in ("WIDE", "HI") else base
        t = int(r[ix["Predicated-On Thread Instructions Executed"]] or 0)
        w = int(r[ix["Instructions Executed"]] or 0)
        thread_inst += t
        warp_inst += w
        mix[key] = mix.get(key, 0) + w
        if base in FLOP:
            flops += FLOP[base] * t
        if key in PIPE_CYC:
            pipe[key] = pipe.get(key, 0) + PIPE_CYC[key] * w
    tot_pipe = sum(pipe.values())
    philox = sum(v for k, v in pipe.items() if k.startswith("IMAD"))
    return {"kernel": block["name"][:110], "executed_fp32_flop": flops, "thread_instructions": thread_inst, "warp_instructions": warp_inst,
            "fma_pipe_cycles_model": tot_pipe, "philox_imad_pipe_share": philox / tot_pipe if tot_pipe else None,
            "warp_inst_mix_top": dict(sorted(mix.items(), key=lambda kv: -kv[1])[:14])}


if __name__ == "__main__":
    rep = sys.argv[1]
    rx = sys.argv[2] if len(sys.argv) > 2 else "rollout_cost_kernel"
    res = [analyse(b) for b in kernels(rep, rx)]
    print(json.dumps(res, indent=1))
