#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_t28.log
: > $O
for e in 2; do
  echo "== SSG_HALO_EPI=$e" >> $O
  SSG_HALO_EPI=$e ONLY="short0_1" timeout 120 python scratch/bench_conv.py fwd dgrad >> $O 2>&1
done
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu 2>&1 | tail -2 >> $O
cat $O
