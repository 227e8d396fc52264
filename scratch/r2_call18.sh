#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sample_persistent -s 6 -c 2 -f -o gpurun_out/r02_c18_headline python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hbm-bound --e2e-steps 0 > gpurun_out/r02_c18_ncu.log 2>&1; echo "ncu rc=$?"
for occ in 4 3; do GNNFLOW_B200_OCC=$occ timeout 300 python bench.py --no-cpu-baseline --no-hbm-bound --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('occ $occ', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"; done
