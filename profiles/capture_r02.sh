#!/bin/bash
# Round-2 ncu captures (one B200, run through gpurun from the repo root):
#   bash profiles/capture_r02.sh [tag]
# Each kernel is first run WITHOUT ncu (must exit 0), then captured once with `--set full`.  The .ncu-rep files land in
# gpurun_out/ (scratch); `python profiles/summarize_ncu.py gpurun_out/<rep>` writes the summaries committed under profiles/.
TAG=${1:-r02}
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
ONLY=conv0_0 cap halo_wgrad_l0   conv_tc_halo_wgrad_kernel 2 python scratch/bench_conv.py wgrad
ONLY=conv1_1 cap halo_wgrad_l1c  conv_tc_halo_wgrad_kernel 2 python scratch/bench_conv.py wgrad
ONLY="spade gb" cap thin_wgrad_gb conv_tc_thin_wgrad_kernel 2 python scratch/bench_conv.py wgrad
ONLY="D.block1" cap s2_dgrad_l0  conv_tc_fwd_kernel 2 python scratch/bench_conv.py dgrad
ONLY="D.block1" cap s2_fwd_l0    conv_tc_fwd_kernel 2 python scratch/bench_conv.py fwd
ONLY="D.block1" cap s2_wgrad_l0  conv_tc_wgrad_kernel 2 python scratch/bench_conv.py wgrad
ONLY=conv0_0 cap halo_fwd_l0     conv_tc_halo_kernel 2 python scratch/bench_conv.py fwd
ONLY=conv0_0 cap halo_dgrad_l0   conv_tc_halo_kernel 2 python scratch/bench_conv.py dgrad
cap bn_bwd_apply bn_bwd_apply_rows_kernel 2 python profiles/hbm_kernels.py --only bn --out $OUT/${TAG}_hbm_bn.json
cap bn_bwd_reduce channel_reduce_vec_kernel 27 python profiles/hbm_kernels.py --only bn --out $OUT/${TAG}_hbm_bn.json
ls -la $OUT/${TAG}_*.ncu-rep
