#!/usr/bin/env python
"""Opcode histogram (executed warp instructions, stall samples, shared wavefronts) of an exported ncu source page:
   python scripts/sass_hist.py gpurun_out/<x>_sass.csv.gz [top]"""
import collections
import csv
import gzip
import sys


def load(p):
    op = gzip.open if p.endswith(".gz") else open
    rows = list(csv.reader(l for l in op(p, "rt") if not l.startswith("==")))
    hdr = rows[1]
    return hdr, rows[2:]


def main():
    hdr, rows = load(sys.argv[1])
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    isrc, ie, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    iwf = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
    c, s, w = collections.Counter(), collections.Counter(), collections.Counter()
    tot = 0
    for r in rows:
        try:
            n = float(r[ie])
        except (ValueError, IndexError):
            continue
        toks = r[isrc].strip().split()
        if not toks:
            continue
        o = toks[1] if toks[0].startswith("@") else toks[0]
        o = o.rstrip(";")
        key = o if o.startswith(("LDS", "STS", "LDL", "STL", "LDG", "STG")) else o.split(".")[0]
        c[key] += n; s[key] += float(r[ismp]); tot += n
        if iwf is not None:
            w[key] += float(r[iwf])
    ssum = sum(s.values())
    print("executed warp instructions %d, samples %d, shared wavefronts %d" % (tot, ssum, sum(w.values())))
    for k, n in c.most_common(top):
        print("%-14s %6.2f %% of instructions  %6.2f %% of stall samples  wavefronts %d" % (k, 100 * n / tot, 100 * s[k] / max(ssum, 1), w[k]))


if __name__ == "__main__":
    main()
