#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r02_c11_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r02_c11_pytest.log
timeout 600 python bench_configs.py --config hbm_bound --shape GDELT-16.7K --scale 0.25 --steps 5 --warmup 3 > gpurun_out/r02_c11_hbm16k.json 2> gpurun_out/r02_c11_hbm16k.err; echo "hbm rc=$?"
python - <<P
import json
d=json.load(open('gpurun_out/r02_c11_hbm16k.json'))
print([(r['strategy'], r['layer'], round(r['ms_per_launch'],3), round(r['frac'],3)) for r in d['hbm_bound']['launches']])
P
timeout 900 python bench.py > gpurun_out/r02_c11_bench.json 2> gpurun_out/r02_c11_bench.err; echo "bench rc=$?"
python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_c11_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['ingest']['value'], d['e2e']['value'], d['per_batch']['us_per_batch'])
for k,v in d['hbm_bound'].items(): print(k, [(r['strategy'], r['layer'], round(r['ms_per_launch'],3), round(r['frac'],3)) for r in v['launches']])
"
for s in recent uniform; do timeout 600 python bench_configs.py --config two_layer_sat --dataset REDDIT --strategy $s --scale 1 > gpurun_out/r02_c11_two_layer_$s.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_c11_two_layer_$s.json').read().strip().splitlines()[-1]); print('$s', json.dumps(d)[:600])"; done
