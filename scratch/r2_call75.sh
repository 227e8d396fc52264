#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c75_pytest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r02_c75_pytest.log
timeout 900 python bench.py > gpurun_out/r02_c75_bench.json 2> gpurun_out/r02_c75_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r02_c75_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r02_c75_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-hbm-bound --no-cpu-baseline --e2e-steps 1 --no-per-batch-models > /dev/null 2>&1; echo "ncu rc $?"
