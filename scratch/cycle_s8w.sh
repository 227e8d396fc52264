#!/usr/bin/env bash
# session 8: two binary-search steps per round trip in the directory searches (locate: end_ts; uniform emit: cum_before)
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sampler or sample" 2>&1 | tail -4) | tee gpurun_out/s8w_pytest.log
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 0 > gpurun_out/s8w_bench.json 2> gpurun_out/s8w_bench.err || tail -5 gpurun_out/s8w_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/s8w_bench.json"))
print("headline: value %.2f G  kernel %.4f ms frac %.3f" % (d["value"]/1e9, d["roofline"]["ms_per_launch"], d["roofline"]["frac"]))
PY
for a in "--dataset REDDIT --strategy uniform" "--dataset REDDIT --strategy recent" "--dataset WIKI --strategy recent" "--dataset WIKI --strategy uniform"; do
  tag=$(echo $a | tr -d ' -')
  timeout 300 python bench_configs.py --config two_layer_sat $a > gpurun_out/s8w_two_layer_$tag.json 2> gpurun_out/s8w_two_layer_$tag.err || tail -5 gpurun_out/s8w_two_layer_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8w_two_layer_$tag.json"))
    print("$tag: %.2f G nbr/s frac %.3f" % (d["value"]/1e9, d["roofline"]["frac"]), [(l["targets"], l["neighbors"], round(l["ms"],4), round(l["frac"],3)) for l in d["layers"]], "chain %.4f ms" % d["chain"]["ms"])
except Exception as e: print("$tag failed", e)
PY
done
