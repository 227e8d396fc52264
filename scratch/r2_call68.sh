#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_async_ingest.py tests/test_gpu_baseline_shapes.py -m gpu -x -q -k "store or ingest or config4 or async" 2>&1 | tail -2
echo "== 16M"; timeout 300 python scratch/ingest_100k.py 16000000
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
GF_SHAPE=GDELT-16.7K GF_NCU_RANGE=1 timeout 300 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_c68_ingest16m_GDELT-16.7K_launches.csv python scratch/ingest_100k.py 16000000 >/dev/null 2>&1
python profiles/launch_summary.py gpurun_out/r02_c68_ingest16m_GDELT-16.7K_launches.csv
