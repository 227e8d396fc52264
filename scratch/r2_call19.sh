#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for rep in 1 2; do
( cd scratch/r1tree && timeout 300 python bench.py --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r1', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])" )
for occ in 4 3; do GNNFLOW_B200_OCC=$occ timeout 300 python bench.py --no-cpu-baseline --no-hbm-bound --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r2 occ $occ', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"; done
done
