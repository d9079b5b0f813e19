set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_t1_gpu.txt
SSG_SPADE_FUSED_TEST=1 timeout 300 python -m pytest tests/test_gpu_spade_fused.py -q -x > gpurun_out/r02_t1_spade_fused_test.log 2>&1
SSG_SPADE_FUSED_TEST=1 timeout 300 python -m pytest tests/test_gpu_spade_fused.py -q > gpurun_out/r02_t1_spade_fused_test_all.log 2>&1
timeout 300 python scratch/bench_spade_fused.py > gpurun_out/r02_t1_spade_fused_bench.log 2>&1
timeout 300 python scratch/bench_conv.py > gpurun_out/r02_t1_bench_conv.log 2>&1
bash profiles/capture_r02.sh r02a
echo done
