#!/usr/bin/env bash
# full verification cycle on one B200: parity tests, both bench arms, launch list, one full ncu capture
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_all.log; cat gpurun_out/pytest_all.log
timeout 600 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -3 gpurun_out/bench_ours.err; cut -c1-600 gpurun_out/bench_ours.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sample_persistent_kernel' -s 3 -c 1 -f -o gpurun_out/prof_persistent \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
