#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_async_ingest.py tests/test_gpu_baseline_shapes.py -m gpu -x -q -k "store or ingest or config4 or async" 2>&1 | tail -4
for sh in GDELT-16.7K GDELT-16.7M; do
GF_SHAPE=$sh GF_NCU_RANGE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_c62_launches_$sh.csv python scratch/ingest_100k.py 16000000 >/dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_c62_launches_$sh.csv')) if len(r)>10 and r[0].isdigit()]
print('$sh', [(r[4][:26], r[8], int(r[-1])//1000) for r in rows])
PY
done
echo "== 16M"; timeout 300 python scratch/ingest_100k.py 16000000
echo "== 16M no bitmap"; GNNFLOW_B200_NO_BOOKKEEP_BITMAP=1 timeout 300 python scratch/ingest_100k.py 16000000
echo "== 4M"; timeout 300 python scratch/ingest_100k.py 4000000
echo "== 100k"; timeout 300 python scratch/ingest_100k.py 100000
