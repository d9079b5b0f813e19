set -x
cd /root/repo
mkdir -p gpurun_out
T=r02_t37
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/${T}_bench_n4.log 2>&1
tail -1 gpurun_out/${T}_bench_n4.log | cut -c1-300
echo done
