set -x
mkdir -p gpurun_out
T=r02_t4
nvidia-smi -L > gpurun_out/${T}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -v -m gpu > gpurun_out/${T}_multi.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n2.log 2>&1
echo done
