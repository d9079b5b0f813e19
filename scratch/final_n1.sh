set -x
cd /root/repo
mkdir -p gpurun_out
T=r02_final
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
TOP=400 timeout 300 python scratch/prof_step.py 16 > gpurun_out/${T}_breakdown.txt 2>&1
TOP=400 timeout 300 python scratch/prof_step.py 8 > gpurun_out/${T}_breakdown_b8.txt 2>&1
timeout 300 python scratch/bench_conv.py > gpurun_out/${T}_conv_layers.txt 2>&1
timeout 300 python profiles/hbm_kernels.py > gpurun_out/${T}_hbm.log 2>&1
tail -3 gpurun_out/${T}_tests.log; tail -1 gpurun_out/${T}_smoke.log; cut -c1-250 gpurun_out/${T}_bench_n1.json
echo done
