#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_t31.log
: > $O
python scratch/prof_step.py > gpurun_out/r02_t31_breakdown.txt 2>&1
grep -E "pack_conv|eager step" gpurun_out/r02_t31_breakdown.txt | head -20 >> $O
cat $O
