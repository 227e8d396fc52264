#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python scratch/ingest_100k.py 16000000 | tee gpurun_out/r02_c24_ingest16m_bigtiles.json
GNNFLOW_B200_SORT_SMALL_TILES=1 timeout 300 python scratch/ingest_100k.py 16000000 | tee gpurun_out/r02_c24_ingest16m_smalltiles.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cache_gather -s 410 -c 1 -f -o gpurun_out/r02_final_gather python scratch/bench_gather.py > gpurun_out/r02_c24_gather.log 2>&1; echo "gather rc=$?"
timeout 300 python scratch/d2h_ceiling.py | tee gpurun_out/r02_d2h_ceiling_1gpu.json
timeout 600 python scratch/prof_api.py > gpurun_out/r02_c24_prof_api.txt 2>&1; echo "prof rc=$?"
grep "==" gpurun_out/r02_c24_prof_api.txt
