set -x
mkdir -p gpurun_out
T=r02_t9
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu -x > gpurun_out/${T}_convtests.log 2>&1
SSG_WGRAD_BN=64 timeout 300 python scratch/bench_conv.py wgrad > gpurun_out/${T}_wgrad_bn64.log 2>&1
timeout 300 python scratch/bench_conv.py wgrad > gpurun_out/${T}_wgrad_bn128.log 2>&1
SSG_WGRAD_BN=64 timeout 300 python scratch/bench_conv.py wgrad > gpurun_out/${T}_wgrad_bn64_b.log 2>&1
timeout 300 python scratch/bench_conv.py wgrad > gpurun_out/${T}_wgrad_bn128_b.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/${T}_tests.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --infer-batch 0 > gpurun_out/${T}_bench.log 2>&1
timeout 600 python profiles/extras_bench.py --out gpurun_out/${T}_extras.json > gpurun_out/${T}_extras.log 2>&1
echo done
