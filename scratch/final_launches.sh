set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --infer-batch 0 > gpurun_out/r02_final_launch_plain.log 2>&1 || exit 1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --infer-batch 0 > gpurun_out/r02_final_launch_ncu.log 2>&1
gzip -9 -f gpurun_out/r02_final_launches.csv
ls -la gpurun_out/r02_final_launches.csv.gz
echo done
