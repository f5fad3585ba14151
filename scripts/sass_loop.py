#!/usr/bin/env python
"""Static opcode histogram of the largest backward-branch loop of each kernel in a cuobjdump -sass listing:
   cuobjdump -sass x.cubin | python scripts/sass_loop.py [name-filter]"""
import collections
import re
import sys

flt = sys.argv[1] if len(sys.argv) > 1 else ""
funcs, cur = [], None
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = (m.group(1), [])
        funcs.append(cur)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur is not None:
        cur[1].append((int(m.group(1), 16), m.group(2).strip()))
def report(name, ins, best):
    body = [t for a, t in ins if best[0] <= a <= best[1]]
    c = collections.Counter()
    for t in body:
        toks = t.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        c[op.split(".")[0]] += 1
    print("%s\n  loop 0x%x..0x%x: %d instructions (of %d)" % (name, best[0], best[1], len(body), len(ins)))
    print("  " + ", ".join("%s %d" % kv for kv in c.most_common(24)))


for name, ins in funcs:
    if flt not in name:
        continue
    loops = {}
    for a, t in ins:
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a:
                loops[tgt] = max(loops.get(tgt, 0), a)
    for best in sorted(loops.items(), key=lambda kv: kv[0] - kv[1])[:2]:
        report(name, ins, best)

