#!/bin/bash
# A/B: lean (straight-line) issue path of the resident-weights halo instance x epilogue warp groups
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_t23.log
: > $O
python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu 2>&1 | tail -3 >> $O
for lean in 0 1; do for epi in 1 2 24; do
  echo "== SSG_HALO_LEAN=$lean SSG_HALO_EPI=$epi" >> $O
  export SSG_HALO_LEAN=$lean SSG_HALO_EPI=$epi
  ONLY=conv0_0.conv2 python scratch/bench_conv.py fwd dgrad >> $O 2>&1
  ONLY="spade gb" python scratch/bench_conv.py fwd >> $O 2>&1
  ONLY="D.conv0" python scratch/bench_conv.py fwd >> $O 2>&1
done; done
cat $O
