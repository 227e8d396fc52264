#!/usr/bin/env bash
set -u
for rep in 1 2 3; do
timeout 300 python bench.py --no-hbm-bound --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.2f G' % (d['value'] / 1e9), 'tgat', d['tgat_per_batch']['us_per_batch'], d['tgat_per_batch']['sample_us_per_batch'], 'tgn', d['tgn_per_batch']['us_per_batch'], 'ingest', d['ingest']['value']/1e9)
"
done
