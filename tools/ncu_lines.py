"""Aggregate an ncu report's source page by CUDA source line: python tools/ncu_lines.py report.ncu-rep [launch index] [top N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
launch = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
if launch is not None:
    cmd += ["--launch-skip", launch, "--launch-count", "1"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
cur = None
hdr = None
rows = []
for ln in out.splitlines():
    if ln.startswith('"File Path"'):
        cur = list(csv.reader([ln]))[0][1].split("/")[-1]
        continue
    if ln.startswith('"Function Name"') or ln.startswith('"Kernel Name"'):
        continue
    r = list(csv.reader([ln]))[0]
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10 or r[2] != "-":
        continue
    try:
        s = int(r[6])
        i = int(r[7])
    except ValueError:
        continue
    rows.append((s, i, cur, r[0], r[1].strip()[:110]))
ts = sum(r[0] for r in rows) or 1
ti = sum(r[1] for r in rows) or 1
print("total samples", ts, "warp instructions", ti)
for s, i, f, l, t in sorted(rows, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%s  %s" % (100 * s / ts, 100 * i / ti, f, l, t))
