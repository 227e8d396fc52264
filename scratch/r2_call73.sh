#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c73_pytest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r02_c73_pytest.log
for v in spin nospin spin nospin; do
if [ $v = nospin ]; then export GNNFLOW_B200_NO_SPIN=1; else unset GNNFLOW_B200_NO_SPIN; fi
timeout 300 python bench.py --no-hbm-bound --no-cpu-baseline --e2e-steps 4 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'value %.2f G' % (d['value'] / 1e9), 'ingest %.3f G %.1f us' % (d['ingest']['value'] / 1e9, d['ingest']['us_per_batch']), 'per_batch %.1f' % d['per_batch']['us_per_batch'],
      'tgat %.1f (sample %.1f)' % (d['tgat_per_batch']['us_per_batch'], d['tgat_per_batch']['sample_us_per_batch']), 'tgn %.1f (sample %.1f)' % (d['tgn_per_batch']['us_per_batch'], d['tgn_per_batch']['sample_us_per_batch']))
"
done
unset GNNFLOW_B200_NO_SPIN
timeout 120 python scratch/ingest_100k.py 100000
timeout 120 python scratch/ingest_100k.py 1000
