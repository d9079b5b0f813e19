#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r02_t25_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_t25_bench.json 2> gpurun_out/r02_t25_bench.err
cat gpurun_out/r02_t25_tests.log; cut -c1-600 gpurun_out/r02_t25_bench.json; tail -3 gpurun_out/r02_t25_bench.err
