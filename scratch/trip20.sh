set -x
mkdir -p gpurun_out
T=r02_t20
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
echo done
