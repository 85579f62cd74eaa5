set -x
python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
tail -c 600 gpurun_out/r2z_bench_n1.json
sh scripts/profile_round2.sh > gpurun_out/r2z_profile.log 2>&1
tail -5 gpurun_out/r2z_profile.log
ls -la gpurun_out | tail -15
