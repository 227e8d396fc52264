#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_arena.py -x -q -rs > gpurun_out/r02_c3_pytest_arena.log 2>&1; echo "arena rc=$?"
tail -30 gpurun_out/r02_c3_pytest_arena.log
for s in "REDDIT 1" "GDELT-16.7K 0.2" "GDELT-16.7M 0.2"; do set -- $s
  timeout 600 python scratch/ingest_bench.py $1 $2 > gpurun_out/r02_c3_ingest_$1.json 2> gpurun_out/r02_c3_ingest_$1.err; echo "ingest $1 rc=$?"
  tail -12 gpurun_out/r02_c3_ingest_$1.err | cut -c1-420
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c3_bench.json 2> gpurun_out/r02_c3_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_c3_bench.err
python - <<'P'
import json
b=json.loads(open('gpurun_out/r02_c3_bench.json').read().strip().splitlines()[-1])
print('value',b['value'],'ms',b['ms_per_step'],'frac',b['roofline']['frac'],'e2e',b['e2e']['value'],'ingest',b['ingest'],'e2e_ing',b['e2e']['ingest'],'pb',b['e2e']['per_batch'])
P
