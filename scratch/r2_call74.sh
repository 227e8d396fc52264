#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r02_c74_pytest_spin.log 2>&1; tail -16 gpurun_out/r02_c74_pytest_spin.log
GNNFLOW_B200_NO_SPIN=1 timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r02_c74_pytest_nospin.log 2>&1; tail -16 gpurun_out/r02_c74_pytest_nospin.log
