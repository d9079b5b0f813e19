set -x
cd /root/repo
mkdir -p gpurun_out
T=r02_t33
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
timeout 600 python bench.py --config sn7 --steps 10 --warmup 3 > gpurun_out/${T}_bench_sn7_n1.json 2> gpurun_out/${T}_bench_sn7_n1.err
tail -3 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; cut -c1-250 gpurun_out/${T}_bench_n1.json; cut -c1-250 gpurun_out/${T}_bench_sn7_n1.json
echo done
