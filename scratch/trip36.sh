set -x
cd /root/repo
mkdir -p gpurun_out
T=r02_t36
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${T}_bench_n8.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 8 --config sn7 --steps 10 --warmup 3 > gpurun_out/${T}_bench_sn7_n8.log 2>&1
tail -1 gpurun_out/${T}_bench_n8.log | cut -c1-300
tail -1 gpurun_out/${T}_bench_sn7_n8.log | cut -c1-300
echo done
