#!/usr/bin/env bash
set -u
for rep in 1 2; do
timeout 300 python bench.py --no-hbm-bound --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
t = d['tgat_per_batch']
print('tgat', t['us_per_batch'], t['sample_us_per_batch'], t['host_us_p50'], t['host_us_max'], t['host_us_spikes'], 'tgn', d['tgn_per_batch']['us_per_batch'], d['tgn_per_batch']['host_us_spikes'])
"
done
