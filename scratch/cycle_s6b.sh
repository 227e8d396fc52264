#!/usr/bin/env bash
# session-6: sort / ingest changes: parity suite, ingest sweep with phase split, bench
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_all.log; cat gpurun_out/pytest_all.log
timeout 600 python bench_configs.py --config ingest_sweep > gpurun_out/ingest_sweep.json 2> gpurun_out/ingest_sweep.err; tail -3 gpurun_out/ingest_sweep.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/ingest_sweep.json"))
for r in d["sweep"]: print(r["batch_edges"], "%.1f M e/s" % (r["edges_per_s"]/1e6), "%.0f us" % r["us_per_batch"], {k: round(v,1) for k,v in r["phase_us_per_batch"].items()})
PY
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -3 gpurun_out/bench_ours.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_ours.json"))
print("value %.2f G  frac %.3f  ingest %.1f M e/s  e2e %.1f M  e2e ingest %.1f M" % (d["value"]/1e9, d["roofline"]["frac"], d["ingest"]["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ingest_edges_per_s"]/1e6))
print(d["ingest"]["phase_ms_per_batch"])
PY
