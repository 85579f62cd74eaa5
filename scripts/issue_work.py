"""Warp instructions per closest-hit query of a whole frame, from an ncu launch list with instruction counters:

    ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none \
        --csv --log-file launches.csv python scripts/one_frame.py synth_profile > frame.json
    python scripts/issue_work.py launches.csv frame.json   ->  profiles/r2_issue_work.json entry + per-kernel shares
"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
frame = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
hdr = None
agg = collections.defaultdict(lambda: collections.defaultdict(float))
for r in rows:
    if "Kernel Name" in r:
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if not hdr or len(r) < len(hdr):
        continue
    k = r[hdr["Kernel Name"]].split("(")[0].replace("void ", "")
    try:
        agg[k][r[hdr["Metric Name"]]] += float(r[hdr["Metric Value"]].replace(",", ""))
        agg[k]["launches:" + r[hdr["Metric Name"]]] += 1
    except ValueError:
        pass
dram = sum(m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0) for m in agg.values())
inst = sum(m["smsp__inst_executed.sum"] for m in agg.values())
tinst = sum(m["smsp__thread_inst_executed.sum"] for m in agg.values())
time_ns = sum(m["gpu__time_duration.sum"] for m in agg.values())
out = {frame["scene"]: {"warp_instr_per_query": inst / frame["rays"], "lanes_per_instr": tinst / inst,
                         "dram_bytes_per_query": (dram / frame["rays"]) if dram else None,
                         "queries": frame["rays"], "samples": frame["samples"], "workload": frame["workload"],
                         "kernels": {k: {"launches": int(m["launches:gpu__time_duration.sum"]),
                                         "time_share": m["gpu__time_duration.sum"] / time_ns,
                                         "warp_instr_per_query": m["smsp__inst_executed.sum"] / frame["rays"],
                                         "lanes_per_instr": m["smsp__thread_inst_executed.sum"] / max(m["smsp__inst_executed.sum"], 1)}
                                     for k, m in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"])},
                         "source": "ncu launch list of scripts/one_frame.py %s (smsp__inst_executed.sum over all launches / queries)" % frame["workload"]}}
print(json.dumps(out, indent=1))
