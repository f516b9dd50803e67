#!/usr/bin/env python
"""Hot region of an ncu source page (--page source --csv): the SASS lines ranked by executed instructions, with opcode classes.
usage: python scripts/source_hot.py gpurun_out/r02_c3a_source.csv [min_fraction_of_max]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ia, isrc, iex, ith, ismp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
body = [r for r in rows[2:] if len(r) > ith]
mx = max(int(r[iex]) for r in body)
tot = sum(int(r[iex]) for r in body)
tots = sum(int(r[ismp]) for r in body)
cls = collections.Counter(); cnt = 0; smp = 0
def kind(op):
    o = op.split()[0] if not op.strip().startswith("@") else op.split()[1]
    o = o.split(".")[0]
    if o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"): return "fp64"
    return o
per = collections.Counter()
for r in body:
    e = int(r[iex])
    per[kind(r[isrc])] += e
    if e >= frac * mx:
        cls[kind(r[isrc])] += 1; cnt += 1; smp += int(r[ismp])
print("total warp instructions %d, samples %d; hot lines (>= %.2f of max %d): %d, holding %.1f %% of samples" % (tot, tots, frac, mx, cnt, 100.0 * smp / max(tots, 1)))
print("hot-line opcode mix:", dict(cls.most_common()))
print("whole-kernel executed mix (%):", {k: round(100.0 * v / tot, 1) for k, v in per.most_common(25)})
