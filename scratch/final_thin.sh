cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_thin_epi.log
: > $O
timeout 200 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_headline.py -x -q -m gpu 2>&1 | tail -2 >> $O
for s in "x2map" "spade gb L0" "D.conv0" "final 1x1"; do ONLY="$s" timeout 60 python scratch/bench_conv.py fwd dgrad >> $O 2>&1; done
cat $O
