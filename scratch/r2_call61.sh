#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_async_ingest.py tests/test_gpu_baseline_shapes.py -m gpu -x -q 2>&1 | tail -2
export GF_SHAPE=GDELT-16.7K
for v in default m0 m1 m3; do
  if [ $v = default ]; then unset GNNFLOW_B200_LIB; else export GNNFLOW_B200_LIB=$PWD/scratch/variants/lib_$v.so; fi
  GF_NCU_RANGE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_c61_launches_$v.csv python scratch/ingest_100k.py 16000000 >/dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_c61_launches_$v.csv')) if len(r)>10 and r[0].isdigit()]
print('$v', [(r[4][7:18], int(r[-1])//1000) for r in rows])
PY
  echo "== $v"; timeout 300 python scratch/ingest_100k.py 16000000
done
unset GNNFLOW_B200_LIB GF_SHAPE
echo "== default all"; timeout 300 python scratch/ingest_100k.py 16000000
