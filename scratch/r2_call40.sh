#!/usr/bin/env bash
# evidence after the record-layout change: launch list of the bench command, full ncu captures of the headline kernel and of
# the sampler on the GDELT shapes, launch list of one TGAT batch
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hbm-bound --e2e-steps 1 --no-per-batch-models"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_c40_launches_bench.csv $B > gpurun_out/r02_c40_list.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sample_persistent -s 6 -c 1 -f -o gpurun_out/r02_c40_headline $B > gpurun_out/r02_c40_headline.log 2>&1; echo "headline rc=$?"
for sh in GDELT-16.7K GDELT-16.7M; do
GF_NCU_RANGE=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sample_persistent -f \
  -o gpurun_out/r02_c40_hbm_$sh python bench_configs.py --config hbm_bound --shape $sh --scale 0.25 --steps 1 --warmup 3 > gpurun_out/r02_c40_hbm_$sh.log 2>&1; echo "hbm $sh rc=$?"
done
timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c40_tgat_batch_launches.csv python scratch/ncu_fetch.py uniform 10,10 > gpurun_out/r02_c40_tgat.log 2>&1; echo "tgat rc=$?"
ls -la gpurun_out/r02_c40_*
