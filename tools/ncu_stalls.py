"""Aggregate warp-stall samples by reason for one kernel from `ncu --page source --csv`."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
seen = set()
agg = Counter()
ninst = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] in seen:
        continue
    seen.add(r[0])
    ninst += 1
    for s in stalls:
        try:
            agg[s] += int(r[col[s]] or 0)
        except ValueError:
            pass
tot = sum(agg.values())
print("kernel:", rows[0][1][:110], "| SASS instructions:", ninst, "| samples:", tot)
for s, v in agg.most_common():
    if v:
        print(f"  {s[6:]:20s} {v:7d} {100*v/tot:5.1f}%")
