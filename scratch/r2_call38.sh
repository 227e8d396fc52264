#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for v in default w16 w32; do
  if [ $v = default ]; then unset GNNFLOW_B200_LIB; else export GNNFLOW_B200_LIB=$PWD/scratch/variants/lib_$v.so; fi
  timeout 300 python scratch/ingest_100k.py 16000000 | tee gpurun_out/r02_c38_ingest16m_$v.json
  timeout 300 python scratch/ingest_100k.py 100000 | tee gpurun_out/r02_c38_ingest100k_$v.json
done
