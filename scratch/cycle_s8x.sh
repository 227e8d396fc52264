#!/usr/bin/env bash
# session 8, call I: ncu of the two launches of the REDDIT two-layer recent replay (where does layer 2's time go?)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sample_persistent -s 4 -c 2 -f -o gpurun_out/prof_s8x_reddit2l \
   python bench_configs.py --config two_layer_sat --dataset REDDIT --strategy uniform --steps 2 > gpurun_out/s8x_ncu.log 2>&1
tail -3 gpurun_out/s8x_ncu.log
python profiles/ncu_summary.py gpurun_out/prof_s8x_reddit2l.ncu-rep --kernel sample_persistent --source 40 > gpurun_out/s8x_reddit2l_summary.txt 2>&1
tail -5 gpurun_out/s8x_reddit2l_summary.txt
ls -la gpurun_out/*.ncu-rep
