#!/usr/bin/env bash
set -u
timeout 300 python bench.py --no-hbm-bound --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
for k in ('tgat_per_batch', 'tgn_per_batch'):
    t = d[k]
    print(k, {q: t[q] for q in ('us_per_batch', 'sample_us_per_batch', 'launches_per_batch', 'device_us_per_batch', 'host_us_p50', 'host_us_max', 'host_us_spikes')})
"
