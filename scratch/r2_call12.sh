#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs -x --timeout 300 > gpurun_out/r02_c12_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_c12_pytest.log
timeout 300 python scratch/ingest_100k.py | tee gpurun_out/r02_c12_ingest_100k_fused.json
GNNFLOW_B200_FUSED_MAX=0 timeout 300 python scratch/ingest_100k.py | tee gpurun_out/r02_c12_ingest_100k_unfused.json
timeout 300 python scratch/ingest_100k.py 10000 | tee gpurun_out/r02_c12_ingest_10k_fused.json
GNNFLOW_B200_FUSED_MAX=0 timeout 300 python scratch/ingest_100k.py 10000 | tee gpurun_out/r02_c12_ingest_10k_unfused.json
timeout 300 python scratch/ingest_100k.py 1000 | tee gpurun_out/r02_c12_ingest_1k_fused.json
