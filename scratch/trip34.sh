cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02_t34_tests.log 2>&1
tail -4 gpurun_out/r02_t34_tests.log
