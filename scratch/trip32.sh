#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_t32.log
: > $O
timeout 600 python -m pytest tests/test_gpu_archs.py -x -q -m gpu 2>&1 | tail -4 >> $O
timeout 300 python scratch/bench_infer.py >> $O 2>&1
cat $O
