"""Summarises an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list into kernel,launches,total_ms,share.
python tools/launch_summary.py gpurun_out/launches.csv "comment line" > profiles/rNN_launch_summary.csv"""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    name = r[ki]
    m = re.search(r"run_kernel<(\w+)>", name)
    name = m.group(1) if m else name.split("(")[0]
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u in ("nsecond", "ns") else v / 1e3 if u in ("usecond", "us") else v * 1e3 if u in ("second", "s") else v
    tot[name] += ms; cnt[name] += 1
allms = sum(tot.values())
for c in sys.argv[2:]: print("# " + c)
print("kernel,launches,total_ms,share")
for k in sorted(tot, key=lambda k: -tot[k]): print("%s,%d,%.3f,%.4f" % (k, cnt[k], tot[k], tot[k] / allms))
