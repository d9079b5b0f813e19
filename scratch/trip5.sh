set -x
mkdir -p gpurun_out
T=r02_t5
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_ops.py -q -m gpu > gpurun_out/${T}_convtests.log 2>&1
SSG_HALO_ISSUERS=1 SSG_S2_MERGE=0 timeout 300 python scratch/bench_conv.py fwd dgrad > gpurun_out/${T}_conv_iss1.log 2>&1
timeout 300 python scratch/bench_conv.py fwd dgrad > gpurun_out/${T}_conv_default.log 2>&1
SSG_HALO_ISSUERS=2 timeout 300 python scratch/bench_conv.py fwd dgrad > gpurun_out/${T}_conv_iss2.log 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --infer-batch 0 > gpurun_out/${T}_bench.log 2>&1
echo done
