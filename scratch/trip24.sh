#!/bin/bash
# A/B: lean streaming issue path (non-resident halo instances)
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_t24.log
: > $O
python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu 2>&1 | tail -3 >> $O
for lean in 0 1; do
  echo "== SSG_HALO_LEAN_STREAM=$lean" >> $O
  SSG_HALO_LEAN_STREAM=$lean ONLY2=1 python scratch/bench_conv.py fwd dgrad 2>&1 | grep -v "s2" >> $O
done
cat $O
