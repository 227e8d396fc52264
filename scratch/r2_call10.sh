#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vs_reference.py tests/test_abi.py -q -rs > gpurun_out/r02_c10_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_c10_pytest.log
for g in 128 64 32; do
  GNNFLOW_B200_L2_FETCH=$g timeout 600 python bench_configs.py --config hbm_bound --shape GDELT-16.7K --scale 0.25 --steps 5 --warmup 3 > gpurun_out/r02_c10_hbm16k_l2f$g.json 2> gpurun_out/r02_c10_hbm16k_l2f$g.err; echo "hbm $g rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/r02_c10_hbm16k_l2f$g.json'))
print($g, [(r['strategy'], r['layer'], round(r['ms_per_launch'],3), round(r['frac'],3)) for r in d['hbm_bound']['launches']])
P
  GNNFLOW_B200_L2_FETCH=$g timeout 300 python scratch/ingest_100k.py 16000000 | tee gpurun_out/r02_c10_ingest_16m_l2f$g.json
done
GNNFLOW_B200_L2_FETCH=128 timeout 300 python scratch/ingest_100k.py | tee gpurun_out/r02_c10_ingest_100k_l2f128.json
timeout 300 python scratch/ingest_100k.py | tee gpurun_out/r02_c10_ingest_100k_l2f32.json
for g in 128 32; do
GNNFLOW_B200_L2_FETCH=$g timeout 600 python bench.py --no-hbm-bound --no-cpu-baseline --e2e-steps 3 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($g, d['value'], d['ms_per_step'], d['roofline']['frac'], d['ingest']['value'], d['e2e']['value'], d['per_batch']['us_per_batch'])"
done
