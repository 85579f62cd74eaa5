#!/bin/sh
# Multi-GPU checks on an N-GPU box (usage: scripts/multi_gpu_check.sh N): bit-equality tests, then C3 through the three arms.
N=${1:-2}
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
for w in dynamic static; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --work $w 2> gpurun_out/multi_$w.err | tail -1 > gpurun_out/multi_nccl_${w}_n$N.json
  echo rc=$?
  python -c "
import json
d=json.loads(open('gpurun_out/multi_nccl_${w}_n$N.json').read())
print('$w', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['frame'])"
done
timeout 300 python bench.py --inprocess --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 > gpurun_out/multi_inprocess_n$N.json
python -c "
import json
d=json.loads(open('gpurun_out/multi_inprocess_n$N.json').read())
print('inprocess', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['frame'])"
