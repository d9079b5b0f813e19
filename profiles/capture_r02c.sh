#!/bin/bash
# Final round-2 captures of the weight-gradient kernels (after the one-wave split-K choice); same recipe as capture_r02b.sh.
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --import-source on --clock-control none"
cap() {
  local name=$1 regex=$2 skip=$3; shift 3
  if "$@" > $OUT/${TAG}_plain_${name}.log 2>&1; then
    timeout 600 $NCU -k regex:$regex -s $skip -c 1 -f -o $OUT/${TAG}_${name} "$@" > $OUT/${TAG}_ncu_${name}.log 2>&1
  else
    echo "plain run of $name failed" >> $OUT/${TAG}_capture_errors.log
  fi
}
export B=16
ONLY=conv1_1 cap halo_wgrad128_l1c conv_tc_halo_wgrad_kernel 2 python scratch/bench_conv.py wgrad
ONLY=conv0_1 cap halo_wgrad64_l0c  conv_tc_halo_wgrad_kernel 2 python scratch/bench_conv.py wgrad
ls -la $OUT/${TAG}_*.ncu-rep
