set -x
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_final_bench_n4.log 2>&1
tail -1 gpurun_out/r02_final_bench_n4.log | cut -c1-260
echo done
