"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python tools/launch_summary.py X.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) != len(h):
        continue
    d = dict(zip(h, r))
    if d["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = d["Kernel Name"].split("(")[0]
    v = float(d["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1e-6)
    tot[name] += v
    cnt[name] += 1
allms = sum(tot.values())
print("total %.3f ms over %d launches" % (allms, sum(cnt.values())))
for name, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print("%9.3f ms %5.1f%% x%-5d %s" % (v, 100 * v / allms, cnt[name], name[:110]))
