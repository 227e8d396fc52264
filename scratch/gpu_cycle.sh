#!/usr/bin/env bash
# one GPU iteration: parity tests (subset or all), bench line, optional ncu of the sampling kernel
# usage: scratch/gpu_cycle.sh <tag> <pytest-target> [ncu-kernel-regex] [bench args...]
TAG=$1; TESTS=$2; KRE=${3:-}; shift 3 2>/dev/null || shift $#
mkdir -p gpurun_out
(timeout 1200 python -m pytest $TESTS -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_$TAG.log; cat gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
r=d["roofline"]
e=d["e2e"]
print("value %.2f G nbr/s  step %.4f ms  kernel %s %.4f ms frac %.3f | e2e %.1f M nbr/s (%.2f ms/step, %.1f GB/s pcie) per-batch %.1f M (%.1f us/batch) | ingest %.1f M e/s (e2e %.1f)" % (
  d["value"]/1e9, d["ms_per_step"], r["kernel"], r["ms_per_launch"], r["frac"], (e["value"] or 0)/1e6, e["ms_per_step"] or 0, e["pcie_GBps"] or 0,
  (e["per_batch"]["value"] or 0)/1e6, (e["per_batch"]["ms_per_batch"] or 0)*1e3, d["ingest"]["value"]/1e6, (e["ingest_edges_per_s"] or 0)/1e6))
print(d["ingest"]["phase_ms_per_batch"])
PY
if [ -n "$KRE" ]; then
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
     python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 "$@" > gpurun_out/ncu_full_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_full_$TAG.log
fi
