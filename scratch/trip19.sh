set -x
mkdir -p gpurun_out
T=r02_t19
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_modules.py -q -m gpu > gpurun_out/${T}_tests.log 2>&1
ONLY=D.block timeout 200 python scratch/bench_conv.py fwd dgrad > gpurun_out/${T}_s2.log 2>&1
ONLY=D.block timeout 200 python scratch/bench_conv.py fwd dgrad >> gpurun_out/${T}_s2.log 2>&1
TOP=60 timeout 300 python scratch/prof_step.py 16 > gpurun_out/${T}_prof_b16.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --infer-batch 0 > gpurun_out/${T}_bench.log 2>&1
echo done
