#!/usr/bin/env bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "store or golden or errors" 2>&1 | tail -3)
for v in "" occ5 occ6; do
  if [ -n "$v" ]; then export GNNFLOW_B200_LIB=$PWD/scratch/variants/lib_$v.so; else unset GNNFLOW_B200_LIB; fi
  timeout 300 python bench.py --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_occ_$v.json 2> gpurun_out/bench_occ_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_occ_$v.json"))
print("variant '$v': value %.2f G  kernel %.4f ms frac %.3f  ingest %.1f M e/s" % (d["value"]/1e9, d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["ingest"]["value"]/1e6))
PY
done
unset GNNFLOW_B200_LIB
for a in "--dataset REDDIT --strategy uniform" "--dataset WIKI --strategy uniform" "--dataset WIKI --strategy recent" "--dataset REDDIT --strategy recent"; do
  timeout 300 python bench_configs.py --config two_layer_sat $a > gpurun_out/two_layer_$(echo $a | tr -d ' -').json 2> gpurun_out/two_layer.err || tail -5 gpurun_out/two_layer.err
  python - <<PY
import json
d=json.load(open("gpurun_out/two_layer_$(echo $a | tr -d ' -').json"))
print("$a: %.2f G nbr/s frac %.3f" % (d["value"]/1e9, d["roofline"]["frac"]), [(l["targets"], l["neighbors"], round(l["ms"],4), round(l["frac"],3), round(l["targets_with_edges_frac"],2)) for l in d["layers"]])
PY
done
timeout 600 python bench_configs.py --config ingest_sweep > gpurun_out/ingest_sweep.json 2> gpurun_out/ingest_sweep.err; tail -3 gpurun_out/ingest_sweep.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/ingest_sweep.json"))
for r in d["sweep"]: print(r["batch_edges"], "%.1f M e/s" % (r["edges_per_s"]/1e6), "%.0f us" % r["us_per_batch"], {k: round(v,1) for k,v in r["phase_us_per_batch"].items()})
PY
