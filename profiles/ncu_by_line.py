#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python ncu_by_line.py src.csv [topN]
Only the first kernel section of the dump is read."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
h = rows[hdr_idx[0]]
end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
ci = {n: h.index(n) for n in ['# Samples', 'Instructions Executed', 'Thread Instructions Executed']}
agg = collections.OrderedDict()
cur = None
for r in rows[hdr_idx[0] + 1:end]:
    if len(r) < len(h) - 5:
        continue
    if r[0].strip():
        cur = (r[0], r[1].strip()[:110])
    if cur is None:
        continue
    def num(x):
        try: return float(x)
        except: return 0.0
    a = agg.setdefault(cur, [0, 0, 0])
    if r[2].strip():  # a SASS row
        a[0] += num(r[ci['# Samples']]); a[1] += num(r[ci['Instructions Executed']]); a[2] += num(r[ci['Thread Instructions Executed']])
tot = [sum(v[k] for v in agg.values()) for k in range(3)]
print(f"total samples={tot[0]:.0f} warp-instr={tot[1]:.0f} thread-instr={tot[2]:.0f}")
print("--- by samples"); 
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0]:8.0f} {100*v[0]/max(tot[0],1):5.1f}% instr={v[1]:10.0f} {100*v[1]/max(tot[1],1):5.1f}%  L{k[0]}: {k[1]}")
print("--- by warp instructions")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"instr={v[1]:10.0f} {100*v[1]/max(tot[1],1):5.1f}% samples={v[0]:8.0f}  L{k[0]}: {k[1]}")
