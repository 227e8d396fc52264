#!/usr/bin/env bash
# Run on the GPU box (under gpurun): launch list of the bench command + full ncu capture of the sampling kernels.
# usage: profiles/run_profile.sh <tag> [variant]
set -u
TAG=${1:-r01}
VAR=${2:-0}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --variant $VAR > gpurun_out/ncu_bench_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'locate_|emit_kernel' -s 8 -c 2 -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --variant $VAR > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/
