#!/usr/bin/env bash
# round 2, GPU call 2: parity of the rewritten store (ingest v2 + reclaiming arena + 64-byte vertex entries)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_async_ingest.py -x -q -rs > gpurun_out/r02_c2_pytest_a.log 2>&1; echo "pytest a rc=$?"
tail -25 gpurun_out/r02_c2_pytest_a.log
timeout 900 python -m pytest tests -m gpu -q -rs --deselect tests/test_gpu_parity.py --deselect tests/test_gpu_async_ingest.py > gpurun_out/r02_c2_pytest_b.log 2>&1; echo "pytest b rc=$?"
tail -25 gpurun_out/r02_c2_pytest_b.log
