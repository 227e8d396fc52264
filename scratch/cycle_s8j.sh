#!/usr/bin/env bash
# session 8, call J: control warp resolves the current tile before the batch lookup of the next (late) vs before (early); pipeline depth 2 / 3, poll back-off, window 64
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sampler or sample" 2>&1 | tail -4) | tee gpurun_out/s8j_pytest.log
run_bench() {  # tag lib variant
  local tag=$1 lib=$2 var=$3
  GNNFLOW_B200_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --e2e-steps 0 --variant $var > gpurun_out/s8j_bench_$tag.json 2> gpurun_out/s8j_bench_$tag.err || tail -5 gpurun_out/s8j_bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8j_bench_$tag.json"))
    print("$tag: value %.2f G  kernel %s %.4f ms frac %.3f" % (d["value"]/1e9, d["roofline"]["kernel"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"]))
except Exception as e: print("$tag failed", e)
PY
}
two_layer() {  # tag lib variant
  local tag0=$1 lib=$2 var=$3
  for a in "--dataset REDDIT --strategy uniform" "--dataset REDDIT --strategy recent" "--dataset WIKI --strategy recent" "--dataset WIKI --strategy uniform"; do
    tag=${tag0}_$(echo $a | tr -d ' -')
    GNNFLOW_B200_LIB=$lib timeout 300 python bench_configs.py --config two_layer_sat --variant $var $a > gpurun_out/s8j_two_layer_$tag.json 2> gpurun_out/s8j_two_layer.err || tail -5 gpurun_out/s8j_two_layer.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8j_two_layer_$tag.json"))
    print("$tag: %.2f G nbr/s frac %.3f" % (d["value"]/1e9, d["roofline"]["frac"]), [(l["targets"], l["neighbors"], round(l["ms"],4), round(l["frac"],3)) for l in d["layers"]])
except Exception as e: print("$tag failed", e)
PY
  done
}
DEF=$PWD/gnnflow_b200/lib/libgnnflow_b200.so
for v in late early late_st3; do
  lib=$PWD/scratch/variants/lib_$v.so; [ $v = late ] && lib=$DEF
  run_bench ${v}_v3 $lib 3
  two_layer ${v}_v3 $lib 3
done
