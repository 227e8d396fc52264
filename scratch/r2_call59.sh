#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export GF_SHAPE=GDELT-16.7K
for v in default u1 u1atom u2; do
  if [ $v = default ]; then unset GNNFLOW_B200_LIB; else export GNNFLOW_B200_LIB=$PWD/scratch/variants/lib_$v.so; fi
  echo "== $v"; timeout 300 python scratch/ingest_100k.py 16000000
  GF_NCU_RANGE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_c59_launches_$v.csv python scratch/ingest_100k.py 16000000 >/dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_c59_launches_$v.csv')) if len(r)>10 and r[0].isdigit()]
print('$v', [(r[4][:22], r[8], int(r[-1])//1000) for r in rows])
PY
done
unset GNNFLOW_B200_LIB
GF_NCU_RANGE=1 timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:ingest_bookkeep -c 1 -o gpurun_out/r02_c59_bookkeep_u4 -f python scratch/ingest_100k.py 16000000 >/dev/null 2>&1
ls -la gpurun_out/
