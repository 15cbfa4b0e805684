"""Key roofline metrics per profiled launch from `ncu -i rep --page raw --csv` (units taken from the csv's unit row)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "second": 1e6, "s": 1e6}


def val(r, name):
    if name not in col:
        return None
    try:
        return float(r[col[name]].replace(",", ""))
    except ValueError:
        return None


def scaled(r, name, table):
    v = val(r, name)
    if v is None:
        return None
    return v * table.get(units[col[name]], 1.0)


print(f"{'kernel':46s} {'us':>8s} {'tensor%':>8s} {'L2%':>6s} {'dram%':>6s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>8s} {'regs':>5s}")
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("void ", "").replace("cmwg::", "").replace("(int)", "").replace("(bool)", "")[:46]
    us = scaled(r, "gpu__time_duration.sum", TIME)
    rd = scaled(r, "dram__bytes_read.sum", BYTES)
    wr = scaled(r, "dram__bytes_write.sum", BYTES)
    gbs = (rd + wr) / (us * 1e-6) / 1e9 if us and rd is not None and wr is not None else None
    f = lambda v, w, p=1: (f"{v:{w}.{p}f}" if v is not None else " " * w)
    print(f"{name:46s} {f(us, 8)} {f(val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'), 8)} "
          f"{f(val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'), 6)} "
          f"{f(val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), 6)} "
          f"{f(rd / 1e6 if rd is not None else None, 8)} {f(wr / 1e6 if wr is not None else None, 8)} {f(gbs, 8, 0)} "
          f"{f(val(r, 'launch__registers_per_thread'), 5, 0)}")
