#!/bin/bash
# Thin-input (Cin = 8) forward through the resident-weights halo instance: what paces it?  Same recipe as capture_r02b.sh.
TAG=${1:-r02d}
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
SSG_HALO_EPI=1 ONLY="spade gb L0" cap halo_thin_in_gb_fwd conv_tc_halo_kernel 2 python scratch/bench_conv.py fwd
SSG_HALO_EPI=1 ONLY="D.conv0" cap halo_thin_in_dconv0_fwd conv_tc_halo_kernel 2 python scratch/bench_conv.py fwd
ls -la $OUT/${TAG}_*.ncu-rep
