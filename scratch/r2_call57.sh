#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c57_pytest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r02_c57_pytest.log
timeout 300 python scratch/ingest_100k.py 16000000
timeout 300 python scratch/ingest_100k.py 100000
GF_SHAPE=GDELT-16.7K GF_NCU_RANGE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_c57_launches.csv python scratch/ingest_100k.py 16000000 >/dev/null 2>&1
grep -o '"ingest_[a-z_]*kernel[^"]*"\|"[0-9]*"$' gpurun_out/r02_c57_launches.csv | paste - - | cut -c1-40,100- | head -12
cut -d, -f5,15 gpurun_out/r02_c57_launches.csv | cut -c1-30,60- | tail -8
