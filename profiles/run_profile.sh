#!/usr/bin/env bash
# Run on the GPU box (under gpurun): launch list of the bench command + one full ncu capture of the dominant kernels.
# usage: profiles/run_profile.sh <tag> [variant] [kernel regex] [extra bench args...]
set -u
TAG=${1:-r01}
VAR=${2:-2}
KRE=${3:-sample_fused_kernel|scatter_kernel}
shift $(( $# < 3 ? $# : 3 ))
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --variant $VAR $*"
# every launch of the bench command with its device time (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_bench_${TAG}.log 2>&1
# full section set for the dominant kernels, skipping the warm-up launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 6 -c 3 -f \
    -o gpurun_out/prof_${TAG} $BENCH > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/
