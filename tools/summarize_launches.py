"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name."""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)
    m = re.search(r"(tc_gemm_kernel|ff_gemm_kernel|tc_wgrad_kernel)<(.*)$", name)
    if m:
        tail = m.group(2)
        epi = re.search(r"cmwg::(\w+Epi)", tail)
        bn = re.search(r"^(\d+)", tail)
        return f"{m.group(1)}<{bn.group(1) if bn else ''},{epi.group(1) if epi else ''}>"
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<.*$", "", name)
    return name[-70:]


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        rows.append((short(r["Kernel Name"]), ns))
    agg = defaultdict(lambda: [0, 0.0])
    for k, ns in rows:
        agg[k][0] += 1
        agg[k][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {total / 1e6:.3f} ms total device time (serialised, cold cache)")
    print(f"{'kernel':72s} {'n':>6s} {'ms':>9s} {'share':>7s} {'us/launch':>10s}")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} {n:6d} {ns / 1e6:9.3f} {100 * ns / total:6.1f}% {ns / n / 1e3:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
