set -x
mkdir -p gpurun_out
T=r02_t21
for i in 1 2 3 4 5 6; do timeout 200 python -m pytest tests/test_gpu_spade_fused.py -q -m gpu 2>&1 | tail -1 >> gpurun_out/${T}_spade_loop.log; done
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/${T}_tests.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -x -p no:randomly > gpurun_out/${T}_tests_b.log 2>&1
echo done
