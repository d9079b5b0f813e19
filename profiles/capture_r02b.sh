#!/bin/bash
# Round-2 ncu captures of the kernels that changed this round (one B200, through gpurun from the repo root):
#   bash profiles/capture_r02b.sh [tag]
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --import-source on --clock-control none"
cap() {   # name, kernel regex, launch-skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  if "$@" > $OUT/${TAG}_plain_${name}.log 2>&1; then
    timeout 600 $NCU -k regex:$regex -s $skip -c 1 -f -o $OUT/${TAG}_${name} "$@" > $OUT/${TAG}_ncu_${name}.log 2>&1
  else
    echo "plain run of $name failed" >> $OUT/${TAG}_capture_errors.log
  fi
}
export B=16
ONLY=conv1_1 cap halo_wgrad128_l1c conv_tc_halo_wgrad_kernel 2 python scratch/bench_conv.py wgrad
ONLY=conv0_0 cap halo_wgrad64_l0   conv_tc_halo_wgrad_kernel 2 python scratch/bench_conv.py wgrad
ONLY="D.block1" cap s2_dgrad_merged_l0 conv_tc_fwd_kernel 2 python scratch/bench_conv.py dgrad
ONLY=conv0_0 cap halo_fwd_l0_2issuers conv_tc_halo_kernel 2 python scratch/bench_conv.py fwd
ONLY=x2map cap halo_thin_x2map_fwd conv_tc_halo_kernel 2 python scratch/bench_conv.py fwd
ls -la $OUT/${TAG}_*.ncu-rep
