set -x
mkdir -p gpurun_out
T=r02_t16
timeout 600 python -m pytest tests/test_gpu_multi.py -v -s -m gpu -k "2" > gpurun_out/${T}_multi2.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_n2.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref_n2.log 2>&1
echo done
