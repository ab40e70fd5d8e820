#!/usr/bin/env python
"""Top stall sites of one kernel from `ncu -i rep --page source --csv` output (SASS view)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Address")
h = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(h) and r[0].startswith("0x")]
ia, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
tot = sum(int(r[isamp]) for r in body)
totex = sum(int(r[iex]) for r in body)
print('total samples', tot, 'total warp inst', totex, 'sass lines', len(body))
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
agg = {h[i]: sum(int(r[i]) for r in body) for i in stall_cols}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for idx, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][isamp]))[:n]:
    st = {h[i][6:]: int(r[i]) for i in stall_cols if int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{idx:5d} {int(r[isamp]):6d} {int(r[iex]):9d}  {r[ia].strip()[:64]:64s} {st}")
