cd /root/repo
mkdir -p gpurun_out
timeout 100 python -m pytest tests -q -m gpu -x > gpurun_out/r02_last_tests.log 2>&1
tail -2 gpurun_out/r02_last_tests.log
timeout 60 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --infer-batch 0 > gpurun_out/r02_last_bench.json 2> gpurun_out/r02_last_bench.err
cut -c1-240 gpurun_out/r02_last_bench.json
