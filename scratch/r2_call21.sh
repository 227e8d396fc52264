#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
run() { # tag, lib, occ
  if [ "$2" = "default" ]; then unset GNNFLOW_B200_LIB; else export GNNFLOW_B200_LIB=$PWD/scratch/variants/lib_$2.so; fi
  GNNFLOW_B200_OCC=$3 timeout 300 python bench.py --no-cpu-baseline --no-hbm-bound --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['ms_per_launch'],4), round(d['roofline']['frac'],3))"
}
( cd scratch/r1tree && timeout 300 python bench.py --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r1', round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['ms_per_launch'],4), round(d['roofline']['frac'],3))" )
for v in default p224 p192; do for occ in 4 3; do run "$v occ$occ" $v $occ; done; done
for v in default p224; do
  if [ "$v" = "default" ]; then unset GNNFLOW_B200_LIB; else export GNNFLOW_B200_LIB=$PWD/scratch/variants/lib_$v.so; fi
  timeout 600 python scratch/sampler_variants.py 0.25 2>/dev/null | tee gpurun_out/r02_c21_var_$v.json | cut -c1-900
done
