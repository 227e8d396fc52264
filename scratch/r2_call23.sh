#!/usr/bin/env bash
# final evidence of round 2: launch list of the bench command, full ncu captures of the dominant kernels
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hbm-bound --e2e-steps 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/r02_c23_list.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sample_persistent -s 6 -c 1 -f -o gpurun_out/r02_final_headline $B > gpurun_out/r02_c23_headline.log 2>&1; echo "headline rc=$?"
for sh in GDELT-16.7K GDELT-16.7M; do
GF_NCU_RANGE=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sample_persistent -f \
  -o gpurun_out/r02_final_hbm_$sh python bench_configs.py --config hbm_bound --shape $sh --scale 0.25 --steps 1 --warmup 3 > gpurun_out/r02_c23_hbm_$sh.log 2>&1; echo "hbm $sh rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cache_gather -s 8 -c 1 -f -o gpurun_out/r02_final_gather python scratch/bench_gather.py > gpurun_out/r02_c23_gather.log 2>&1; echo "gather rc=$?"
timeout 600 python scratch/bench_gather.py > gpurun_out/r02_gather_sweep.json 2>/dev/null; echo "gather sweep rc=$?"
ls -la gpurun_out/r02_final_* gpurun_out/r02_launches_bench.csv
