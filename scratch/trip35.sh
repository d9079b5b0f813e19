set -x
cd /root/repo
mkdir -p gpurun_out
T=r02_t35
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_n2.log 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/${T}_multi_tests_n2.log 2>&1
tail -2 gpurun_out/${T}_multi_tests_n2.log
tail -1 gpurun_out/${T}_bench_n2.log | cut -c1-300
echo done
