#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_t30.log
: > $O
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu 2>&1 | tail -2 >> $O
echo "== old paths (SSG_HALO_LEAN=0 SSG_HALO_EPI=1 SSG_S2_HALO=0)" >> $O
SSG_HALO_LEAN=0 SSG_HALO_EPI=1 SSG_S2_HALO=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'])" >> $O
echo "== default" >> $O
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'])" >> $O
python scratch/prof_step.py 2>&1 | grep -E "dgrad_tc_mask|eager step" | head -3 >> $O
cat $O
