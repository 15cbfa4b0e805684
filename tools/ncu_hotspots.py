"""Top stall locations of one kernel from `ncu -i rep --page source --csv` (SASS view)."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
# find the header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
samp = col["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) <= samp:
        continue
    try:
        n = int(r[samp])
    except ValueError:
        continue
    data.append((n, r))
total = sum(n for n, _ in data)
print("kernel:", rows[0][1][:120] if rows[0] else "")
print("total samples", total)
for n, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[col[s]] or 0), s) for s in stalls), reverse=True)[:3]
    sts = " ".join(f"{s[6:]}={v}" for v, s in st if v)
    print(f"{n:7d} {100*n/total:5.1f}%  {r[col['Address']][-6:]}  {r[col['Source']][:90]:90s} {sts}")
