set -x
mkdir -p gpurun_out
T=r02_t3
SSG_WGRAD_ISSUERS=1 timeout 300 python scratch/bench_conv.py wgrad > gpurun_out/${T}_wgrad_issuers1.log 2>&1
SSG_WGRAD_ISSUERS=2 timeout 300 python scratch/bench_conv.py wgrad > gpurun_out/${T}_wgrad_issuers2.log 2>&1
timeout 900 python -m pytest tests/test_gpu_headline.py -q -s -m gpu > gpurun_out/${T}_headline.log 2>&1
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_headline.py > gpurun_out/${T}_tests.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.log 2>&1
timeout 300 python scratch/prof_step.py > gpurun_out/${T}_prof_step.log 2>&1
echo done
