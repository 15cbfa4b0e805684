"""Count the SASS mnemonics that prove tcgen05 / TMEM / TMA use, per kernel of libcmwg_b200.so (cuobjdump -sass)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "constant_memory_waveglow_b200", "lib", "libcmwg_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU.EX2", "MUFU.TANH", "MUFU.RCP",
        "F2FP.SATFINITE", "HMMA", "FFMA")
cur, counts = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*$", "", cur).replace("void ", "").replace("cmwg::", "")[:70]
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for w in want:
        if re.search(r"\b" + re.escape(w) + r"\b", ln) or (w.endswith(".") and w in ln) or (("." in w) and w in ln):
            counts[cur][w] += 1
print(f"{'kernel':72s} " + " ".join(f"{w:>9s}" for w in want))
tot = collections.Counter()
for k, c in counts.items():
    if not any(c[w] for w in want[:6]) and "mega" not in k and "tc_" not in k:
        continue
    print(f"{k:72s} " + " ".join(f"{c[w]:9d}" for w in want))
    tot.update(c)
print(f"{'TOTAL (all kernels)':72s} " + " ".join(f"{sum(c[w] for c in counts.values()):9d}" for w in want))
