#!/bin/sh
# One bench line per BASELINE.json config (C1..C5) with the reference timed on the host cores beside it.
# usage (GPU box): scripts/config_table.sh > gpurun_out/config_table.jsonl
for w in cornell rtiow synth ods tof media; do
  spp=4; [ "$w" = ods ] && spp=2
  python bench.py --workload $w --steps 2 --warmup 3 --cpu-samples-sqrt $spp 2>/dev/null | tail -1
done
