#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_arena.py tests/test_gpu_async_ingest.py -m gpu -q -rs > gpurun_out/r02_c8_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r02_c8_pytest.log
