set -x
cd /root/repo
mkdir -p gpurun_out
T=r02_t39
for ov in 0 1; do
SSG_OVERLAP_GRAD_SYNC=$ov timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2960$ov bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_n8_ov$ov.log 2>&1
echo "overlap=$ov $(grep '^{"metric"' gpurun_out/${T}_bench_n8_ov$ov.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])")"
done
echo done
