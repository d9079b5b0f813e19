set -x
mkdir -p gpurun_out
T=r02_t14
for st in 0 2 3 0 2 3; do
  ONLY=D.block SSG_PLAIN_STAGES=$st timeout 200 python scratch/bench_conv.py fwd dgrad >> gpurun_out/${T}_plain_stages_$st.log 2>&1
done
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu > gpurun_out/${T}_convtests.log 2>&1
echo done
