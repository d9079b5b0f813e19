set -x
mkdir -p gpurun_out
T=r02_t2
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python -m pytest tests/test_gpu_headline.py -x -q -s -m gpu > gpurun_out/${T}_headline.log 2>&1
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_headline.py > gpurun_out/${T}_tests.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.log 2>&1
timeout 600 python bench.py --config sn7 --steps 5 --warmup 3 > gpurun_out/${T}_bench_sn7.log 2>&1
timeout 300 python scratch/prof_step.py > gpurun_out/${T}_prof_step.log 2>&1
timeout 300 python profiles/hbm_kernels.py --out gpurun_out/${T}_hbm.json > gpurun_out/${T}_hbm.log 2>&1
echo done
