"""Aggregate an `ncu --csv` launch list by kernel: launches, total time, mean of the other metrics."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows:
    if "Kernel Name" in r:
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if not hdr or len(r) < len(hdr):
        continue
    k = r[hdr["Kernel Name"]].split("(")[0].replace("void ", "")[-44:]
    try:
        agg[k][r[hdr["Metric Name"]]].append(float(r[hdr["Metric Value"]].replace(",", "")))
    except ValueError:
        pass
total = sum(sum(m.get("gpu__time_duration.sum", [0])) for m in agg.values())
for k, m in sorted(agg.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
    t = m.get("gpu__time_duration.sum", [0])
    extra = "  ".join("%s=%.2f" % (n.split(".")[0][-30:], sum(v) / max(len(v), 1)) for n, v in m.items() if n != "gpu__time_duration.sum")
    print("%-44s launches %4d  total %8.3f ms  share %5.1f%%  mean %7.1f us  %s" % (k, len(t), sum(t) / 1e6, 100 * sum(t) / max(total, 1), sum(t) / len(t) / 1e3, extra))
