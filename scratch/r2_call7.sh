#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r02_c7_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02_c7_pytest.log
timeout 900 python bench.py > gpurun_out/r02_c7_bench.json 2> gpurun_out/r02_c7_bench.err; echo "bench rc=$?"
cut -c1-3000 gpurun_out/r02_c7_bench.json
tail -5 gpurun_out/r02_c7_bench.err
