#!/usr/bin/env bash
# session 8: ticket drawn by the leader worker while the previous tile is emitted (claim when the CTA is about to be free)
mkdir -p gpurun_out
LT=$PWD/scratch/variants/lib_lt.so
(GNNFLOW_B200_LIB=$LT timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sampler or sample" 2>&1 | tail -4) | tee gpurun_out/s8y_pytest.log
for tag in def lt; do
  lib=$PWD/gnnflow_b200/lib/libgnnflow_b200.so; [ $tag = lt ] && lib=$LT
  GNNFLOW_B200_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --e2e-steps 0 > gpurun_out/s8y_bench_$tag.json 2> gpurun_out/s8y_bench_$tag.err || tail -5 gpurun_out/s8y_bench_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/s8y_bench_$tag.json"))
print("$tag headline: value %.2f G  kernel %.4f ms frac %.3f" % (d["value"]/1e9, d["roofline"]["ms_per_launch"], d["roofline"]["frac"]))
PY
  for a in "--dataset REDDIT --strategy uniform" "--dataset REDDIT --strategy recent" "--dataset WIKI --strategy recent" "--dataset WIKI --strategy uniform"; do
    t=${tag}_$(echo $a | tr -d ' -')
    GNNFLOW_B200_LIB=$lib timeout 300 python bench_configs.py --config two_layer_sat $a > gpurun_out/s8y_two_layer_$t.json 2> gpurun_out/s8y_two_layer_$t.err || tail -5 gpurun_out/s8y_two_layer_$t.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8y_two_layer_$t.json"))
    print("$t: %.2f G nbr/s frac %.3f" % (d["value"]/1e9, d["roofline"]["frac"]), [(l["targets"], l["neighbors"], round(l["ms"],4), round(l["frac"],3)) for l in d["layers"]])
except Exception as e: print("$t failed", e)
PY
  done
done
