timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
for w in dynamic static; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --work $w 2> gpurun_out/r2p_$w.err | tail -1 > gpurun_out/r2p_nccl_${w}_n2.json
  echo rc=$?
  python -c "
import json
d=json.loads(open('gpurun_out/r2p_nccl_${w}_n2.json').read())
print('$w', d['value'], d['ms_per_step'], d['frame'], d['config']['parallelism'][:60])"
  tail -2 gpurun_out/r2p_$w.err
done
timeout 200 python bench.py --inprocess --steps 5 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | cut -c1-200
