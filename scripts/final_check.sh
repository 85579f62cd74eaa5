#!/bin/sh
# One GPU-box call at the end of a round: A/B of experiment builds, then the bench line and the whole GPU test suite with
# the build that is to ship (the best trace-kernel candidate if it beats the default by >= 1 % on synth, else the default).
# Everything goes to gpurun_out/final/.
#   usage: scripts/final_check.sh "trace-kernel candidates" "small-frame candidates (cornell only)" ["ENV=a|ENV=b ..."]
out=gpurun_out/final
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $out/gpu.txt 2>&1
sh scripts/ab.sh "default $1" "synth cornell" > $out/ab.txt 2>&1
[ -n "$2" ] && sh scripts/ab.sh "$2" "cornell" >> $out/ab.txt 2>&1
cat $out/ab.txt
if [ -n "$3" ]; then
  sh scripts/ab_env.sh "$3" "synth" > $out/ab_env.txt 2>&1
  cat $out/ab_env.txt
fi
best=$(python - $out/ab.txt $1 <<'P'
import sys
val = {}
for line in open(sys.argv[1]):
    w = line.split()
    if len(w) > 3 and w[1] == "synth" and w[2] == "Msamples/s":
        val[w[0]] = float(w[3])
best = "default"
for v in sys.argv[2:]:
    if v in val and "default" in val and val[v] >= 1.01 * val["default"] and val[v] > val.get(best, 0.0):
        best = v
print(best)
P
)
echo "shipping candidate: $best" | tee $out/best.txt
if [ "$best" != "default" ]; then export WPT_LIB=$PWD/wurblpt_b200/lib/libwurblpt_b200_$best.so; fi
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
tail -c 400 $out/bench_n1.json
python -m pytest tests -x -q -m gpu > $out/pytest_gpu.txt 2>&1
tail -5 $out/pytest_gpu.txt
