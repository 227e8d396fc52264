#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rs --timeout 400 > gpurun_out/r02_c22_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_c22_pytest.log
timeout 900 python bench.py > gpurun_out/r02_c22_bench.json 2> gpurun_out/r02_c22_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_c22_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['ingest']['value'], d['e2e']['value'], d['per_batch']['us_per_batch'], d['e2e_device']['value'])
for k,v in d['hbm_bound'].items(): print(k, [(r['strategy'], r['layer'], round(r['ms_per_launch'],3), round(r['frac'],3)) for r in v['launches']])
"
