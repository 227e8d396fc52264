#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_async_ingest.py tests/test_gpu_arena.py tests/test_gpu_baseline_shapes.py tests/test_gpu_checkpoint.py tests/test_gpu_vs_reference.py -m gpu -x -q > gpurun_out/r02_c58_pytest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r02_c58_pytest.log
echo "== default"; timeout 300 python scratch/ingest_100k.py 16000000
echo "== plan small tiles"; GNNFLOW_B200_PLAN_SMALL_TILES=1 timeout 300 python scratch/ingest_100k.py 16000000
echo "== atom"; GNNFLOW_B200_LIB=$PWD/scratch/variants/lib_atom.so timeout 300 python scratch/ingest_100k.py 16000000
echo "== default 100k"; timeout 300 python scratch/ingest_100k.py 100000
echo "== nscg 100k"; GNNFLOW_B200_LIB=$PWD/scratch/variants/lib_nscg.so timeout 300 python scratch/ingest_100k.py 100000
for sh in GDELT-16.7K GDELT-16.7M; do
GF_SHAPE=$sh GF_NCU_RANGE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_c58_launches_$sh.csv python scratch/ingest_100k.py 16000000 >/dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_c58_launches_$sh.csv')) if len(r)>10 and r[0].isdigit()]
print('$sh', [(r[4][:22], r[8], int(r[-1])//1000) for r in rows])
PY
done
