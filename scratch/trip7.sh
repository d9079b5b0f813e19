set -x
mkdir -p gpurun_out
T=r02_t7
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/${T}_tests.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --infer-batch 0 > gpurun_out/${T}_bench.log 2>&1
TOP=400 timeout 300 python scratch/prof_step.py 16 > gpurun_out/${T}_prof_b16.log 2>&1
echo done
