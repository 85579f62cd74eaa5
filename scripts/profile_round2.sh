#!/bin/sh
# The ncu passes whose summaries are committed under profiles/ (GPU box; each pass only after the plain run exited 0).
set -e
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2_plain.json
# 1. launch list of the bench command (cold-cache, serialised per-launch times: the kernels' SHARE of a step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2_launches_bench.log 2>&1
python scripts/ncu_kernel_split.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_kernel_split.txt
# 2. instructions and DRAM bytes of every launch of one (reduced) frame -> per-query figures
python scripts/one_frame.py synth_profile > gpurun_out/r2_one_frame_plain.json
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r2_one_frame.csv python scripts/one_frame.py synth_profile > gpurun_out/r2_one_frame.json 2> gpurun_out/r2_one_frame.err
python scripts/issue_work.py gpurun_out/r2_one_frame.csv gpurun_out/r2_one_frame.json > gpurun_out/r2_issue_work.json
# 3. full capture of one steady-state round (trace, sort, material, generate)
ncu --set full --import-source on --clock-control none -k regex:wave -s 300 -c 5 -o gpurun_out/r2_round \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs > gpurun_out/r2_round.log 2>&1
