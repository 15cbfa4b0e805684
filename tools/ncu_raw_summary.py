"""Key roofline metrics per profiled launch from `ncu -i rep --page raw --csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "us"), ("sm__cycles_elapsed.max", "cyc"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "inst")]
print(f"{'kernel':44s} " + " ".join(f"{n:>9s}" for _, n in want))
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    name = name.replace("void ", "").replace("cmwg::", "").replace("(int)", "").replace("(bool)", "")[:44]
    vals = []
    for m, _ in want:
        v = r[col[m]] if m in col else ""
        try:
            vals.append(f"{float(v.replace(',', '')):9.1f}")
        except ValueError:
            vals.append(f"{v:>9s}")
    print(f"{name:44s} " + " ".join(vals))
