#!/usr/bin/env python
"""Register-operand-bandwidth model of a kernel's hot loop, from its SASS (no GPU needed).

    python tools/sass_operand_model.py OBJECT_OR_SO 'rollout_cost_kernel<3, 0, true, false, 7>' [--loop N]

tools/probe_pipes2.cu measured on B200 that an SM sub-partition delivers TWO 32-bit register source operands per lane
per cycle: FFMA with a constant / uniform operand issues every cycle, a three-register FFMA every 1.5, FFMA2 with three
distinct register pairs every 3 (not 2), and MUFU (8 cycles) overlaps the FMA pipe fully.  So the cost of a warp
instruction is max(pipe cycles, register source words / 2), and the bound of a loop is the larger of its issue slots, its
per-pipe cycle sums and its total register reads / 2.  Operands served by the reuse cache (`.reuse` set by the previous
reader of the same register in the same operand slot) are free.

The hot loop is the N-th largest backward branch region (default: the largest).
"""
import re
import subprocess
import sys
from collections import Counter

FMA_PIPE = {"FFMA": 1, "FMUL": 1, "FADD": 1, "FFMA2": 2, "FMUL2": 2, "FADD2": 2, "IMAD": 2, "IMAD.WIDE": 4, "IMAD.HI": 4,
            "HFMA2": 1, "IMAD.MOV": 1, "IMAD.SHL": 1, "IMAD.IADD": 1, "FSWZADD": 1}
ALU_PIPE = {"LOP3", "SHF", "IADD3", "MOV", "SEL", "FSEL", "FMNMX", "ISETP", "FSETP", "PRMT", "LEA", "IABS", "I2FP", "F2FP", "PLOP3",
            "FCHK", "IADD", "CS2R", "VIADD", "UMOV", "FMNMX3", "LOP", "POPC", "FLO", "BREV", "VIMNMX", "IMNMX", "SGXT", "BMSK", "P2R", "R2P"}
FLOP = {"FFMA": 2, "FMUL": 1, "FADD": 1, "FFMA2": 4, "FMUL2": 2, "FADD2": 2}
XU = {"MUFU": 8, "F2I": 4, "I2F": 4, "F2F": 4, "FRND": 4}


_SASS = {}


def function_sass(path, pattern):
    if path not in _SASS:
        txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        blks = txt.split("Function : ")[1:]
        names = subprocess.run(["c++filt"], input="\n".join(b.split("\n", 1)[0].strip() for b in blks), capture_output=True, text=True).stdout.split("\n")
        _SASS[path] = list(zip(names, blks))
    for name, blk in _SASS[path]:
        if pattern in name:
            return name, blk
    raise LookupError(f"no function matching {pattern!r}")


def parse(blk):
    ins = []
    for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", blk):
        text = m.group(2).strip()
        pred = None
        if text.startswith("@"):
            pred, text = text.split(None, 1)
        ins.append((int(m.group(1), 16), pred, text))
    return ins


def opcode_key(op):
    parts = op.split(".")
    if parts[0] == "IMAD" and len(parts) > 1 and parts[1] in ("WIDE", "HI", "MOV", "SHL", "IADD"):
        return parts[0] + "." + parts[1]
    return parts[0]


def source_words(text):
    """[(slot, register number, words, sets_reuse)] of the register SOURCE operands."""
    op, _, rest = text.partition(" ")
    ops = [o.strip() for o in rest.split(",")]
    if not ops or ops == [""]:
        return []
    key = opcode_key(op)
    n_dst = 1
    if key in ("ISETP", "FSETP", "PLOP3"):
        n_dst = 2
    if key in ("ST", "STS", "STG", "STL", "RED", "BAR", "BRA", "EXIT", "BSSY", "BSYNC", "ATOMS", "ATOMG", "MEMBAR", "NOP", "WARPSYNC", "SYNCS", "UTMALDG", "UBLKCP"):
        n_dst = 0
    srcs = ops[n_dst:]
    out = []
    for slot, o in enumerate(srcs):
        m = re.search(r"(?<![UP\w])R(\d+)((?:\.\w+)*)", o)
        if not m or re.match(r"^-?\|?RZ", o.lstrip("-~|")):
            continue
        mods = m.group(2)
        words = 2 if ("F32x2" in mods or ".64" in mods or (key in ("FFMA2", "FMUL2", "FADD2") and ".F32" not in mods.replace(".F32x2", ""))) else 1
        if key == "IMAD.WIDE" and slot == 2:
            words = 2                      # 64-bit addend
        if "[" in o:                      # address register of a memory operand
            words = 2 if ".64" in o else 1
        out.append((slot, int(m.group(1)), words, ".reuse" in mods))
    return out


def model(path, pattern, nth=0):
    """Counts and cost sums of the nth-largest loop of the first function whose demangled name contains `pattern`."""
    name, blk = function_sass(path, pattern)
    ins = parse(blk)
    loops = []
    for addr, pred, text in ins:
        m = re.match(r"BRA(?:\.\w+)*\s+(?:P\d,\s*)?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            loops.append((addr - int(m.group(1), 16), int(m.group(1), 16), addr))
    loops.sort(reverse=True)
    _, lo, hi = loops[nth]
    body = [(a, p, t) for a, p, t in ins if lo <= a <= hi]
    pipe, reads_by, counts, unknown = Counter(), Counter(), Counter(), Counter()
    reads, cost_max, flop = 0, 0.0, 0
    reuse = {}                             # slot -> register kept in the reuse cache by the previous instruction
    for a, p, t in body:
        key = opcode_key(t.split(" ", 1)[0])
        counts[key] += 1
        w, new_reuse = 0, {}
        for slot, reg, words, sets in source_words(t):
            if reuse.get(slot) != reg:
                w += words
            if sets:
                new_reuse[slot] = reg
        reuse = new_reuse
        reads += w
        if key in FMA_PIPE:
            pipe["fma"] += FMA_PIPE[key]
            reads_by["fma"] += w
        elif key in XU:
            pipe["xu"] += XU[key]
            reads_by["xu"] += w
        elif key in ALU_PIPE:
            pipe["alu"] += 2
            reads_by["alu"] += w
        else:
            pipe["other"] += 1
            reads_by["other"] += w
            unknown[key] += 1
        cost_max += max(FMA_PIPE.get(key, 1), w / 2.0)
        flop += FLOP.get(key, 0)
    return {"fp32_flop_per_thread": flop, "name": name, "loop": (lo, hi), "instructions": len(body), "counts": counts, "other": unknown, "pipe": pipe,
            "register_source_words": reads, "words_by_pipe": reads_by, "serial_cost_cycles": cost_max}


def main():
    path, pattern = sys.argv[1], sys.argv[2]
    nth = int(sys.argv[sys.argv.index("--loop") + 1]) if "--loop" in sys.argv else 0
    m = model(path, pattern, nth)
    pipe, reads = m["pipe"], m["register_source_words"]
    print(m["name"][:100])
    print(f"loop 0x{m['loop'][0]:x}..0x{m['loop'][1]:x}: {m['instructions']} instructions")
    print("opcode counts:", dict(m["counts"].most_common(16)))
    print("other-pipe opcodes:", dict(m["other"]))
    print(f"pipe cycles per iteration: fma {pipe['fma']}, alu {pipe['alu']} (half-rate pipe), xu {pipe['xu']}")
    print(f"register source words per iteration: {reads}  -> {reads / 2:.0f} cycles at 2 words/lane/cycle   by pipe: {dict(m['words_by_pipe'])}")
    print(f"sum over instructions of max(FMA-pipe cycles or 1 issue slot, words/2): {m['serial_cost_cycles']:.0f} cycles")
    print(f"FP32 FLOP per thread per iteration (FFMA 2, FMUL / FADD 1, packed x2): {m['fp32_flop_per_thread']}")
    print(f"lower bounds per warp-iteration per scheduler: issue {m['instructions']}, fma pipe {pipe['fma']}, xu {pipe['xu']}, operand bandwidth {reads / 2:.0f}")


if __name__ == "__main__":
    main()
