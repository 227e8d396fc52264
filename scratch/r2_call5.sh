#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r02_c5_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02_c5_pytest.log
timeout 300 python scratch/ingest_100k.py | tee gpurun_out/r02_c5_ingest_100k.json
timeout 300 python scratch/ingest_100k.py 16000000 | tee gpurun_out/r02_c5_ingest_16m.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ingest_ -s 40 -c 60 --csv --log-file gpurun_out/r02_c5_ingest_launches_100k.csv python scratch/ingest_100k.py > /dev/null 2> gpurun_out/r02_c5_ncu_list.err; echo "ncu list rc=$?"
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_c5_ingest_launches_100k.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
for r in rows[1:26]: print(r[ki][:40], r[gi], r[vi])
P
