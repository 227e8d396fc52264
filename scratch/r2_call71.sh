#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c71_pytest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r02_c71_pytest.log
timeout 900 python bench.py > gpurun_out/r02_c71_bench.json 2> gpurun_out/r02_c71_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r02_c71_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c71_bench_reference.json 2> gpurun_out/r02_c71_bench_reference.err; echo "ref rc $?"; tail -c 400 gpurun_out/r02_c71_bench_reference.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
