#!/usr/bin/env bash
# final verification of the round on one B200: every GPU test, smoke(), both bench arms, launch list, ncu of the headline kernel
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/final_pytest.log; cat gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -2 gpurun_out/final_bench.err; cut -c1-300 gpurun_out/final_bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; cut -c1-200 gpurun_out/final_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/final_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/final_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sample_persistent_kernel' -s 3 -c 1 -f -o gpurun_out/final_prof_persistent \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/final_ncu_full.log 2>&1
python profiles/ncu_summary.py gpurun_out/final_prof_persistent.ncu-rep --kernel sample_persistent --source 30 > gpurun_out/final_prof_persistent_summary.txt 2>&1
head -20 gpurun_out/final_prof_persistent_summary.txt | tail -8
