#!/usr/bin/env bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cache.py -m gpu -x -q 2>&1 | tail -3)
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_runahead.json 2> gpurun_out/bench_runahead.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_runahead.json"))
print("run-ahead: value %.2f G  kernel %.4f ms frac %.3f  ingest %.1f M e/s" % (d["value"]/1e9, d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["ingest"]["value"]/1e6))
PY
for a in "--dataset REDDIT --strategy uniform" "--dataset WIKI --strategy recent" "--dataset REDDIT --strategy recent"; do
  timeout 300 python bench_configs.py --config two_layer_sat $a > gpurun_out/two_layer_$(echo $a | tr -d ' -').json 2> gpurun_out/two_layer.err || tail -5 gpurun_out/two_layer.err
  python - <<PY
import json
d=json.load(open("gpurun_out/two_layer_$(echo $a | tr -d ' -').json"))
print("$a: %.2f G nbr/s frac %.3f" % (d["value"]/1e9, d["roofline"]["frac"]), [(l["targets"], l["neighbors"], round(l["ms"],4), round(l["frac"],3)) for l in d["layers"]])
PY
done
timeout 600 python scratch/bench_gather.py > gpurun_out/gather_sweep.json 2> gpurun_out/gather_sweep.err; grep -v Warn gpurun_out/gather_sweep.err | tail -20
