#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs --timeout 300 > gpurun_out/r02_c13_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r02_c13_pytest.log
