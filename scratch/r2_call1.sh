#!/usr/bin/env bash
# round 2, GPU call 1: full GPU test suite, reference-uniform diagnosis, headline bench, first HBM-bound numbers
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -rs > gpurun_out/r02_c1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_c1_pytest.log
tail -15 gpurun_out/r02_c1_pytest.log
python scratch/diag_ref_uniform.py /tmp/diag > gpurun_out/r02_ref_uniform_diag.json 2> gpurun_out/r02_ref_uniform_diag.err; echo "diag rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c1_bench.json 2> gpurun_out/r02_c1_bench.err; echo "bench rc=$?"
python bench_configs.py --config hbm_bound --shape GDELT-16.7K --scale 1.0 > gpurun_out/r02_c1_hbm_16k.json 2> gpurun_out/r02_c1_hbm_16k.err; echo "hbm16k rc=$?"
python bench_configs.py --config hbm_bound --shape GDELT-16.7M --scale 1.0 > gpurun_out/r02_c1_hbm_16m.json 2> gpurun_out/r02_c1_hbm_16m.err; echo "hbm16m rc=$?"
tail -c 1500 gpurun_out/r02_c1_hbm_16k.json; tail -c 600 gpurun_out/r02_c1_hbm_16k.err; tail -c 600 gpurun_out/r02_c1_hbm_16m.err
