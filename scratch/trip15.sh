set -x
mkdir -p gpurun_out
T=r02_t15
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.log 2>&1
timeout 600 python bench.py --config sn7 --steps 10 --warmup 3 > gpurun_out/${T}_bench_sn7.log 2>&1
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_ref.log 2>&1
timeout 300 python scratch/bench_conv.py > gpurun_out/${T}_bench_conv.log 2>&1
timeout 300 python profiles/hbm_kernels.py --out gpurun_out/${T}_hbm.json > gpurun_out/${T}_hbm.log 2>&1
timeout 600 python profiles/extras_bench.py --out gpurun_out/${T}_extras.json > gpurun_out/${T}_extras.log 2>&1
TOP=400 timeout 300 python scratch/prof_step.py 16 > gpurun_out/${T}_prof_b16.log 2>&1
TOP=400 timeout 300 python scratch/prof_step.py 8 > gpurun_out/${T}_prof_b8.log 2>&1
bash profiles/capture_r02c.sh r02c
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --infer-batch 0 > gpurun_out/${T}_ncu_launches.log 2>&1
echo done
