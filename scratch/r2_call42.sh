#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for rep in 1 2; do
for v in default p192 p256 st3; do
  if [ $v = default ]; then unset GNNFLOW_B200_LIB; else export GNNFLOW_B200_LIB=$PWD/scratch/variants/lib_$v.so; fi
  timeout 300 python bench.py --steps 30 --no-hbm-bound --no-cpu-baseline --e2e-steps 0 --no-per-batch-models 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'rep$rep', 'value %.2f G' % (d['value'] / 1e9), 'kernel %.4f ms' % d['roofline']['ms_per_launch'], 'frac %.3f' % d['roofline']['frac'])
"
done
done
