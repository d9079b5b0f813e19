set -x
cd /root/repo
mkdir -p gpurun_out
T=r02_t38
for ov in 0 1 0 1; do
SSG_OVERLAP_GRAD_SYNC=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2959$ov bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_n2_ov$ov.log 2>&1
echo "overlap=$ov $(grep '^{"metric"' gpurun_out/${T}_bench_n2_ov$ov.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])")"
done
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/${T}_multi_tests_n2.log 2>&1
tail -2 gpurun_out/${T}_multi_tests_n2.log
echo done
