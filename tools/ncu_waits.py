"""List the sampled SASS hot spots (>= pct of samples) plus every barrier wait of one kernel."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
pct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
seen, data = set(), []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] in seen or r[0] == "Address":
        continue
    seen.add(r[0])
    data.append(r)
def n_of(r):
    try:
        return int(r[col["# Samples"]] or 0)
    except ValueError:
        return 0
tot = sum(n_of(r) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for i, r in enumerate(data):
    n = n_of(r)
    src = r[col["Source"]]
    if "TRYWAIT" in src or n > tot * pct / 100:
        st = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
        print(f"{i:5d} {n:6d} {100 * n / tot:5.1f}%  {src[:90]:90s} {' '.join(f'{s}={v}' for v, s in st if v)}")
