#!/usr/bin/env python
"""loop.py <obj> <func-substr> [marker=BREV|UFLO...]: print the innermost loop that contains the marker
(from the loop head to the backward branch) and per-class instruction counts."""
import re, subprocess, sys, collections
obj, fn = sys.argv[1], sys.argv[2]
marker = sys.argv[3] if len(sys.argv) > 3 else "BREV"
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
body = [f for f in funcs if fn in f.split("\n")[0]]
assert body, "function not found"
ins = []
for line in body[0].split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
mi = next(i for i, (a, t) in enumerate(ins) if marker in t)
# backward branches after the marker whose target is at or before the marker
best = None
for i in range(mi, len(ins)):
    a, t = ins[i]
    m = re.search(r"BRA(?:\.U)?(?:\.ANY)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) <= ins[mi][0]:
        best = (i, int(m.group(1), 16)); break
assert best, "no backward branch"
hi = best[0]; lo = next(i for i, (a, t) in enumerate(ins) if a >= best[1])
loop = ins[lo:hi + 1]
cls = collections.Counter()
for a, t in loop:
    op = t.split()[1] if t.startswith("@") else t.split()[0]
    cls[op.split(".")[0]] += 1
if "-q" not in sys.argv:
    for a, t in loop: print(f"{a:05x}  {t}")
print(f"{fn}: loop {len(loop)} instrs, total {len(ins)};", dict(cls.most_common()))
