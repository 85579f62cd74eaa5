#!/bin/sh
# usage: scripts/ab_env.sh "ENV1=a ENV2=b|ENV1=c ..." [workloads]   (settings separated by |)
IFS='|'
for setting in $1; do
  IFS=' '
  for w in ${2:-synth cornell}; do
    env $setting python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$setting]', '$w', 'Msamples/s %.1f  Mrays/s %.1f  e2e %.1f  frac %.3f launches %d' % (d['value'], d['mrays_per_s'], d['e2e']['value'], (d['roofline'] or {}).get('frac', 0), d['gpu_launches']))"
  done
  IFS='|'
done
