"""Warp-stall samples per CUDA source line from `ncu -i rep --page source --csv --print-source sass,cuda`."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg, text = Counter(), {}
fpath, col = "", None
seen = set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        col = {h: i for i, h in enumerate(r)}
        continue
    if col is None or len(r) < len(col):
        continue
    try:
        n = int(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    key = (fpath, r[0])
    addr = r[2]
    if (key, addr) in seen:
        continue
    seen.add((key, addr))
    agg[key] += n
    text[key] = r[1].strip()[:110]
tot = sum(agg.values())
print("total samples", tot)
for k, v in agg.most_common(top):
    print(f"{v:7d} {100 * v / tot:5.1f}%  {k[0]}:{k[1]:>4s}  {text[k]}")
