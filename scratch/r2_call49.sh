#!/usr/bin/env bash
set -u
for v in default pft; do
  if [ $v = default ]; then unset GNNFLOW_B200_LIB; else export GNNFLOW_B200_LIB=$PWD/scratch/variants/lib_$v.so; fi
  for sh in GDELT-16.7M GDELT-16.7K; do
  timeout 600 python bench_configs.py --config hbm_bound --shape $sh --scale 1.0 --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v', '$sh', [(l['strategy'], l['layer'], round(l['ms_per_launch'], 3), round(l.get('frac', 0), 3)) for l in d['hbm_bound']['launches'] if l['layer'] != 'chain'])
"
  done
  for rep in 1 2; do
  timeout 300 python bench.py --steps 30 --no-hbm-bound --no-cpu-baseline --e2e-steps 0 --no-per-batch-models 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'headline value %.2f G' % (d['value'] / 1e9), 'kernel %.4f ms' % d['roofline']['ms_per_launch'], 'frac %.3f' % d['roofline']['frac'])
"
  done
done
