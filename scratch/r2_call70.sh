#!/usr/bin/env bash
set -u
for ch in 1 2 4 8 16; do
GNNFLOW_B200_HOST_CHUNKS=$ch timeout 300 python bench.py --no-hbm-bound --no-cpu-baseline --e2e-steps 5 --steps 5 --no-per-batch-models 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
e = d['e2e']
print('chunks $ch', 'e2e %.3f G' % (e['value'] / 1e9), round(e['ms_per_step'], 3), round(e['pcie_GBps'], 1), 'int64: %.3f ms' % e['other_format']['ms_per_step'])
"
done
