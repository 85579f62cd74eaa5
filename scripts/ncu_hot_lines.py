"""Summarise `ncu --page source --print-source cuda,sass --csv` output: hottest source lines.
usage: python scripts/ncu_hot_lines.py report.ncu-rep [top] [kernel-name regex]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sel = ["--kernel-name", "regex:" + sys.argv[3], "--launch-count", "1"] if len(sys.argv) > 3 else []
out = subprocess.run(["ncu", "-i", rep] + sel + ["--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = "?"
hdr = None
data = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        # the second "Source" column is the SASS text; the first is the CUDA source
        continue
    if hdr is None or len(r) < 12 or r[0] == "":
        continue
    try:
        samp = float(r[hdr["# Samples"]]); inst = float(r[hdr["Instructions Executed"]])
        tinst = float(r[hdr["Thread Instructions Executed"]])
        lsb = float(r[hdr["stall_long_sb"]]); 
    except Exception:
        continue
    data.append((samp, inst, tinst, lsb, cur_file, r[0], r[1].strip()[:95]))
ts = sum(d[0] for d in data) or 1
ti = sum(d[1] for d in data) or 1
tt = sum(d[2] for d in data)
print("total samples %d, warp instructions %.3g, threads/instruction %.2f" % (ts, ti, tt / ti))
data.sort(reverse=True)
print("%6s %6s %5s %6s  %s" % ("samp%", "inst%", "thr", "longsb%", "source"))
for d in data[:top]:
    print("%6.2f %6.2f %5.1f %6.1f  %s:%s  %s" % (100 * d[0] / ts, 100 * d[1] / ti, d[2] / max(d[1], 1), 100 * d[3] / max(d[0], 1), d[4], d[5], d[6]))
