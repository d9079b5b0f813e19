#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_t27.log
: > $O
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k linear 2>&1 | tail -5 >> $O
timeout 120 python scratch/bench_linear.py >> $O 2>&1
cat $O
