#!/usr/bin/env bash
# session-6 first cycle: parity suite at HEAD, both bench arms, ingest sweep with phase split
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_all.log; cat gpurun_out/pytest_all.log
timeout 600 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -3 gpurun_out/bench_ours.err; cut -c1-400 gpurun_out/bench_ours.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
timeout 600 python bench_configs.py --config ingest_sweep > gpurun_out/ingest_sweep.json 2> gpurun_out/ingest_sweep.err; tail -3 gpurun_out/ingest_sweep.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/ingest_sweep.json"))
for r in d["sweep"]: print(r["batch_edges"], "%.1f M e/s" % (r["edges_per_s"]/1e6), "%.0f us" % r["us_per_batch"], {k: round(v,1) for k,v in r["phase_us_per_batch"].items()})
PY
