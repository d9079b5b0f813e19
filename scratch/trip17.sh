set -x
mkdir -p gpurun_out
T=r02_t17
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/${T}_bench_n4.log 2>&1
echo done
