#!/usr/bin/env bash
# session 8, last experiment: multi-batch launches run over the list of targets whose vertex has out-edges
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batched" 2>&1 | tail -12 | cut -c1-300) | tee gpurun_out/s9b_pytest.log
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 > gpurun_out/s9b_bench.json 2> gpurun_out/s9b_bench.err || tail -5 gpurun_out/s9b_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/s9b_bench.json"))
print("headline: value %.2f G  step %.4f ms kernel %.4f ms frac %.3f launches %d e2e %.2f G" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["gpu_launches"], d["e2e"]["value"]/1e9))
PY
for a in "--dataset REDDIT --strategy uniform" "--dataset REDDIT --strategy recent" "--dataset WIKI --strategy recent" "--dataset WIKI --strategy uniform"; do
  tag=$(echo $a | tr -d ' -')
  timeout 300 python bench_configs.py --config two_layer_sat $a > gpurun_out/s9b_two_layer_$tag.json 2> gpurun_out/s9b_two_layer_$tag.err || tail -5 gpurun_out/s9b_two_layer_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s9b_two_layer_$tag.json"))
    print("$tag: %.2f G nbr/s frac %.3f" % (d["value"]/1e9, d["roofline"]["frac"]), [(l["targets"], l["neighbors"], round(l["ms"],4), round(l["frac"],3)) for l in d["layers"]], "chain %.4f" % d["chain"]["ms"])
except Exception as e: print("$tag failed", e)
PY
done
