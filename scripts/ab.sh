#!/bin/sh
# A/B helper for the GPU box: runs bench.py with each experiment build of the library.
# usage: scripts/ab.sh "variant1 variant2 ..." [workloads]
for v in $1; do
  for w in ${2:-synth cornell}; do
    if [ "$v" = "default" ]; then lib=""; else lib="$PWD/wurblpt_b200/lib/libwurblpt_b200_$v.so"; fi
    WPT_LIB=$lib python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '$w', 'Msamples/s %.1f  Mrays/s %.1f  e2e %.1f  frac %.3f' % (d['value'], d['mrays_per_s'], d['e2e']['value'], (d['roofline'] or {}).get('frac', 0)))"
  done
done
