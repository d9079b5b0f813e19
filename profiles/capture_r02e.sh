#!/bin/bash
# Captures after the lean issue paths / halo-form stride-2 data gradient (same recipe as capture_r02b.sh).
TAG=${1:-r02e}
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
ONLY="D.block1 s2" cap s2_dgrad_halo_l0 conv_tc_halo_s2dgrad_kernel 2 python scratch/bench_conv.py dgrad
ONLY="D.conv0" cap halo_thin_in_dconv0_fwd_lean conv_tc_halo_kernel 2 python scratch/bench_conv.py fwd
ONLY="conv0_0.conv2" cap halo_fwd_l0_lean conv_tc_halo_kernel 2 python scratch/bench_conv.py fwd
ls -la $OUT/${TAG}_*.ncu-rep
