set -x
mkdir -p gpurun_out
T=r02_t10
nvidia-smi -L > gpurun_out/${T}_gpus.txt
timeout 500 python -m pytest tests/test_gpu_multi.py -v -s -m gpu -k "8" > gpurun_out/${T}_multi8.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n8.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config sn7 --steps 5 --warmup 3 > gpurun_out/${T}_bench_sn7_n8.log 2>&1
echo done
