#!/usr/bin/env python
"""Static per-source-line instruction count of a kernel's main loop (nvdisasm -gi output).

    cuobjdump -xelf all libmppi_b200.so; nvdisasm -gi *.cubin > all.sass
    python tools/sass_lines.py all.sass <mangled-kernel-substring>
"""
import collections
import re
import sys

txt = open(sys.argv[1]).read()
want = sys.argv[2]
secs = re.split(r'\n\s*//-+ \.text\.', txt)
sec = [s for s in secs if want in s.split('\n', 1)[0]][0]
cur, ins = None, []
for l in sec.split('\n'):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2), cur))
best = None
labels = {}
# resolve branch targets written as `(.L_x_N) by locating label lines
pos = {}
addr = 0
for l in sec.split('\n'):
    m = re.match(r'\s*(\.L_x_\d+):', l)
    if m:
        pos[m.group(1)] = None
        labels[m.group(1)] = len(pos)
order = []
pending = []
for l in sec.split('\n'):
    m = re.match(r'\s*(\.L_x_\d+):', l)
    if m:
        pending.append(m.group(1))
        continue
    m = re.search(r'/\*([0-9a-f]{4,5})\*/', l)
    if m and pending:
        for p in pending:
            pos[p] = int(m.group(1), 16)
        pending = []
for a, t, _ in ins:
    if 'BRA' in t:
        m = re.search(r'`\((\.L_x_\d+)\)', t)
        tgt = pos.get(m.group(1)) if m else None
        if tgt is not None and tgt < a and (best is None or a - tgt > best[1] - best[0]):
            best = (tgt, a)
agg = collections.Counter()
ops = collections.Counter()
for a, t, c in ins:
    if best and best[0] <= a <= best[1]:
        agg[c] += 1
        ops[re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0]] += 1
print("loop", best, "static instrs", sum(agg.values()))
print(dict(ops.most_common(14)))
for (c, n) in agg.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 25):
    print(f"{n:5d}  {c}")
