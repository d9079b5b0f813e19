#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r02_t29_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_t29_bench.json 2> gpurun_out/r02_t29_bench.err
cat gpurun_out/r02_t29_tests.log; cut -c1-300 gpurun_out/r02_t29_bench.json; tail -3 gpurun_out/r02_t29_bench.err
bash profiles/capture_r02e.sh
python scratch/prof_step.py > gpurun_out/r02_t29_breakdown.txt 2>&1; head -30 gpurun_out/r02_t29_breakdown.txt
