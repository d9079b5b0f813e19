set -x
mkdir -p gpurun_out
T=r02_t13
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu > gpurun_out/${T}_convtests.log 2>&1
timeout 300 python scratch/bench_conv.py wgrad > gpurun_out/${T}_wgrad.log 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --infer-batch 0 > gpurun_out/${T}_bench.log 2>&1
echo done
