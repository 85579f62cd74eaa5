set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipeline_trace_kernels or secondary_ray" 2>&1 | tail -15 > gpurun_out/c7_tests.log
cat gpurun_out/c7_tests.log
scripts/ab_env.sh "WPT_TRACE_QUAD=0|WPT_TRACE_QUAD=1|WPT_TRACE_QUAD=1 WPT_LIB=$PWD/wurblpt_b200/lib/libwurblpt_b200_b9.so|WPT_TRACE_QUAD=0 WPT_LIB=$PWD/wurblpt_b200/lib/libwurblpt_b200_b9.so" "synth cornell rtiow" > gpurun_out/c7_ab.txt 2>&1
cat gpurun_out/c7_ab.txt
