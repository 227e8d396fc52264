#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r02_c6_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_c6_pytest.log
timeout 300 python scratch/ingest_100k.py | tee gpurun_out/r02_c6_ingest_100k_pdl.json
GNNFLOW_B200_LIB=$PWD/scratch/variants/lib_nopdl.so timeout 300 python scratch/ingest_100k.py | tee gpurun_out/r02_c6_ingest_100k_nopdl.json
timeout 300 python scratch/ingest_100k.py 16000000 | tee gpurun_out/r02_c6_ingest_16m.json
timeout 300 python scratch/ingest_100k.py 1000 | tee gpurun_out/r02_c6_ingest_1k.json
