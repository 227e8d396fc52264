#!/usr/bin/env bash
# session 7, call A: parity of the warp-autonomous kernels (variants 4/5) + A/B against the persistent kernel
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sampler or sample" 2>&1 | tail -4) | tee gpurun_out/s7a_pytest.log
run_bench() {  # tag lib variant
  local tag=$1 lib=$2 var=$3
  GNNFLOW_B200_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --e2e-steps 0 --variant $var > gpurun_out/s7a_bench_$tag.json 2> gpurun_out/s7a_bench_$tag.err || tail -5 gpurun_out/s7a_bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s7a_bench_$tag.json"))
    print("$tag: value %.2f G  kernel %s %.4f ms frac %.3f" % (d["value"]/1e9, d["roofline"]["kernel"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"]))
except Exception as e: print("$tag failed", e)
PY
}
DEF=$PWD/gnnflow_b200/lib/libgnnflow_b200.so
run_bench v3 $DEF 3
run_bench v4_occ5 $DEF 4
run_bench v5_occ5 $DEF 5
run_bench v4_occ4 $PWD/scratch/variants/lib_wocc4.so 4
run_bench v5_occ4 $PWD/scratch/variants/lib_wocc4.so 5
run_bench v4_occ6 $PWD/scratch/variants/lib_wocc6.so 4
run_bench v5_occ6 $PWD/scratch/variants/lib_wocc6.so 5
for var in 3 4 5; do
for a in "--dataset REDDIT --strategy uniform" "--dataset REDDIT --strategy recent" "--dataset WIKI --strategy recent"; do
  tag=v${var}_$(echo $a | tr -d ' -')
  timeout 300 python bench_configs.py --config two_layer_sat --variant $var $a > gpurun_out/s7a_two_layer_$tag.json 2> gpurun_out/s7a_two_layer.err || tail -5 gpurun_out/s7a_two_layer.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s7a_two_layer_$tag.json"))
    print("$tag: %.2f G nbr/s frac %.3f" % (d["value"]/1e9, d["roofline"]["frac"]), [(l["targets"], l["neighbors"], round(l["ms"],4), round(l["frac"],3)) for l in d["layers"]])
except Exception as e: print("$tag failed", e)
PY
done
done
