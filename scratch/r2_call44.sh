#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for v in default ao6 ao8; do
  if [ $v = default ]; then unset GNNFLOW_B200_LIB; else export GNNFLOW_B200_LIB=$PWD/scratch/variants/lib_$v.so; fi
  timeout 300 python scratch/ingest_100k.py 16000000
  timeout 300 python scratch/ingest_100k.py 100000
done
