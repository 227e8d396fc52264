#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_async_ingest.py tests/test_gpu_baseline_shapes.py tests/test_gpu_arena.py tests/test_gpu_checkpoint.py -m gpu -x -q -k "store or ingest or config4 or async or arena or checkpoint" 2>&1 | tail -4
echo "== 16M"; timeout 300 python scratch/ingest_100k.py 16000000
echo "== 2M"; timeout 300 python scratch/ingest_100k.py 2000000
