#!/bin/bash
# A/B: one vs two epilogue warp groups in the halo kernel (resident-weights instance and the wide instance)
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_t22.log
: > $O
for epi in 1 2 24; do
  echo "== SSG_HALO_EPI=$epi" >> $O
  SSG_HALO_EPI=$epi ONLY=conv0_0.conv2 python scratch/bench_conv.py fwd dgrad >> $O 2>&1
  SSG_HALO_EPI=$epi ONLY="spade gb" python scratch/bench_conv.py fwd >> $O 2>&1
  SSG_HALO_EPI=$epi ONLY="D.conv0" python scratch/bench_conv.py fwd >> $O 2>&1
done
for epi in 1 2; do
  echo "== SSG_HALO_EPI_WIDE=$epi" >> $O
  SSG_HALO_EPI_WIDE=$epi ONLY=conv1_ python scratch/bench_conv.py fwd dgrad >> $O 2>&1
  SSG_HALO_EPI_WIDE=$epi ONLY=conv2_1 python scratch/bench_conv.py fwd dgrad >> $O 2>&1
  SSG_HALO_EPI_WIDE=$epi ONLY=dgradsplit python scratch/bench_conv.py dgrad >> $O 2>&1
done
python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu 2>&1 | tail -3 >> $O
cat $O
