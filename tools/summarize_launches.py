"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv "header line 1" ["header line 2" ...] > profiles/...txt
"""
import collections
import csv
import re
import sys

path, headers = sys.argv[1], sys.argv[2:]
rows = []
with open(path, newline="") as fh:
    lines = [ln for ln in fh if ln.startswith('"')]
for rec in csv.DictReader(lines):
    if rec.get("Metric Name") != "gpu__time_duration.sum":
        continue
    value = float(rec["Metric Value"].replace(",", ""))
    unit = rec["Metric Unit"]
    ms = value * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[unit]
    name = re.sub(r"\(.*", "", rec["Kernel Name"]).strip()
    rows.append((name, ms))
total = sum(ms for _, ms in rows)
agg = collections.OrderedDict()
for name, ms in rows:
    cnt, tot = agg.get(name, (0, 0.0))
    agg[name] = (cnt + 1, tot + ms)
for h in headers:
    print(h)
print("launches in the timed region: %d   sum of kernel durations: %.2f ms" % (len(rows), total))
print()
print("%-92s %8s %10s %7s" % ("kernel", "launches", "total_ms", "share"))
for name, (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print("%-92s %8d %10.3f %6.1f%%" % (name[:92], cnt, tot, 100.0 * tot / total))
