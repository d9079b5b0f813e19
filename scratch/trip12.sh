set -x
mkdir -p gpurun_out
T=r02_t12
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
SSG_WGRAD_WAVES=2 timeout 300 python scratch/bench_conv.py wgrad > gpurun_out/${T}_wgrad_2waves.log 2>&1
timeout 300 python scratch/bench_conv.py wgrad > gpurun_out/${T}_wgrad_model.log 2>&1
SSG_WGRAD_WAVES=2 timeout 300 python scratch/bench_conv.py wgrad > gpurun_out/${T}_wgrad_2waves_b.log 2>&1
timeout 300 python scratch/bench_conv.py wgrad > gpurun_out/${T}_wgrad_model_b.log 2>&1
SSG_WGRAD_WAVES=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --infer-batch 0 > gpurun_out/${T}_bench_2waves.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --infer-batch 0 > gpurun_out/${T}_bench_model.log 2>&1
echo done
