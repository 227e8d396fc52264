#!/usr/bin/env bash
# session 8: device-side chaining of multi-batch launches -- parity test, then the two-layer saturation configs with the
# chaining pass inside the timed region
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batched" 2>&1 | tail -5) | tee gpurun_out/s8s_pytest.log
for a in "--dataset REDDIT --strategy uniform" "--dataset REDDIT --strategy recent" "--dataset WIKI --strategy recent" "--dataset WIKI --strategy uniform"; do
  tag=$(echo $a | tr -d ' -')
  timeout 300 python bench_configs.py --config two_layer_sat $a > gpurun_out/s8s_two_layer_$tag.json 2> gpurun_out/s8s_two_layer_$tag.err || tail -5 gpurun_out/s8s_two_layer_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8s_two_layer_$tag.json"))
    print("$tag: %.2f G nbr/s frac %.3f" % (d["value"]/1e9, d["roofline"]["frac"]), [(l["targets"], l["neighbors"], round(l["ms"],4), round(l["frac"],3)) for l in d["layers"]], "chain %.4f ms frac %.3f" % (d["chain"]["ms"], d["chain"]["frac"]))
except Exception as e: print("$tag failed", e)
PY
done
