#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python scratch/diag_ref_uniform.py /tmp/diag > gpurun_out/r02_c9_ref_uniform_diag.json 2> gpurun_out/r02_c9_ref_uniform_diag.err; echo "diag rc=$?"
# sampler, graph >> L2: full ncu of the timed launches (recent L0, L1, uniform L0, L1)
GF_NCU_RANGE=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sample_persistent -f \
  -o gpurun_out/r02_c9_sampler_hbm python bench_configs.py --config hbm_bound --shape GDELT-16.7K --scale 0.25 --steps 1 --warmup 3 \
  > gpurun_out/r02_c9_sampler_hbm.log 2>&1; echo "ncu sampler rc=$?"
# ingest: one 9.5 M-edge batch into the 16.7 M-vertex shape, and 100k-edge batches
GF_SHAPE=GDELT-16.7M GF_NCU_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ingest_ -c 8 -f \
  -o gpurun_out/r02_c9_ingest_16m python scratch/ingest_100k.py 16000000 > gpurun_out/r02_c9_ingest_16m.log 2>&1; echo "ncu ingest16m rc=$?"
GF_SHAPE=REDDIT GF_NCU_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ingest_ -c 5 -f \
  -o gpurun_out/r02_c9_ingest_100k python scratch/ingest_100k.py 100000 > gpurun_out/r02_c9_ingest_100k.log 2>&1; echo "ncu ingest100k rc=$?"
ls -la gpurun_out/*c9*
