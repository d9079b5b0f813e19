#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_t26.log
: > $O
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu 2>&1 | tail -15 >> $O
for s2 in 0 1; do
  echo "== SSG_S2_HALO=$s2" >> $O
  SSG_S2_HALO=$s2 ONLY="s2" timeout 120 python scratch/bench_conv.py dgrad >> $O 2>&1
done
cat $O
