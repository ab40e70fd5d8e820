#!/usr/bin/env python
"""Per CUDA source line: warp instructions executed, stall samples and the top stall reasons, from
`ncu -i rep --page source --csv --print-source cuda,sass` (source rows carry the totals of their SASS rows).
usage: ncu_lines.py file.csv [n] [inst|samp]"""
import csv, sys, collections, os
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(r for r in rows if r and r[0] == "Line No" and len(r) > 10)
isamp, iex = h.index('# Samples'), h.index('Instructions Executed')
stall = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
cur, out = "", []
for r in rows:
    if r and r[0] == "File Name":
        cur = os.path.basename(r[1])
    if len(r) < len(h) - 1 or not r[0].isdigit():
        continue
    ex, sa = (int(r[iex]) if r[iex].isdigit() else 0), (int(r[isamp]) if r[isamp].isdigit() else 0)
    st = collections.Counter({h[i][6:]: int(r[i]) for i in stall if i < len(r) and r[i].isdigit() and int(r[i])})
    out.append((cur, r[0], ex, sa, st, r[1]))
tex, ts = sum(o[2] for o in out), sum(o[3] for o in out)
print("total warp inst", tex, "samples", ts)
k = 2 if (len(sys.argv) > 3 and sys.argv[3] == "inst") else 3
for f, line, ex, sa, st, txt in sorted(out, key=lambda o: -o[k])[:n]:
    print(f"{f[:14]:14s}{line:>4} inst {ex:>9} ({100*ex/max(tex,1):4.1f}%) samp {sa:>6} ({100*sa/max(ts,1):4.1f}%) {dict(st.most_common(2))} | {txt.strip()[:80]}")
