"""Per-kernel device times of one chunk from an ncu launch list (gpu__time_duration.sum):

    python tools/launch_times.py launches.csv

Prints, per kernel name, the number of launches, total and mean time, and the per-launch series of the query kernels."""
import csv
import collections
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
per = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r and "Metric Value" in r:
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[hdr["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[hdr["Kernel Name"]].split("(")[0].replace("void r3d::", "")
    val = float(r[hdr["Metric Value"]].replace(",", ""))
    unit = r[hdr["Metric Unit"]]
    us = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1000.0)
    per.setdefault(name, []).append(us)
tot = sum(sum(v) for v in per.values())
for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    print("%-60s n=%4d total %9.1f us (%5.1f%%) mean %8.1f" % (k[:60], len(v), sum(v), 100 * sum(v) / tot, sum(v) / len(v)))
for k, v in per.items():
    if "nn_" in k:
        print(k[:60], " ".join("%.0f" % x for x in v))
