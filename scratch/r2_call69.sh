#!/usr/bin/env bash
set -u
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_abi.py -m gpu -x -q -k "batched or abi" 2>&1 | tail -3
timeout 300 python bench.py --no-hbm-bound --no-cpu-baseline --e2e-steps 6 > gpurun_out/r02_c69_bench.json 2>gpurun_out/r02_c69_bench.err; tail -3 gpurun_out/r02_c69_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_c69_bench.json').read().strip().splitlines()[-1])
e = d['e2e']
print('value %.2f G' % (d['value'] / 1e9), 'e2e %.3f G' % (e['value'] / 1e9), e['ms_per_step'], e['pcie_GBps'], e['d2h_bytes_per_step'], e['other_format'])
print('tgat', d['tgat_per_batch']['us_per_batch'], d['tgat_per_batch']['passes_us_per_batch'], 'tgn', d['tgn_per_batch']['us_per_batch'])
PY
