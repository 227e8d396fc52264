#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r02_c4_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02_c4_pytest.log
for s in "REDDIT 1" "GDELT-16.7K 0.2" "GDELT-16.7M 0.2"; do set -- $s
  timeout 600 python scratch/ingest_bench.py $1 $2 > gpurun_out/r02_c4_ingest_$1.json 2> gpurun_out/r02_c4_ingest_$1.err; echo "ingest $1 rc=$?"
  grep -E '"batch_edges": (100000|16000000|4000000)' gpurun_out/r02_c4_ingest_$1.err | cut -c1-400
  tail -2 gpurun_out/r02_c4_ingest_$1.err | cut -c1-300
done
for v in "default 4" "default 3" "cond 4" "cond 3"; do set -- $v
  if [ "$1" = "default" ]; then unset GNNFLOW_B200_LIB; else export GNNFLOW_B200_LIB=$PWD/scratch/variants/lib_$1.so; fi
  GNNFLOW_B200_OCC=$2 timeout 600 python scratch/sampler_variants.py 0.25 2> gpurun_out/r02_c4_var_$1_$2.err | tee gpurun_out/r02_c4_var_$1_$2.json | cut -c1-900
done
unset GNNFLOW_B200_LIB
# ncu launch list of the ingest kernels: 100k-edge and 16M-edge batches
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ingest_ -c 120 --csv --log-file gpurun_out/r02_c4_ingest_launches.csv python scratch/ingest_bench.py GDELT-16.7K 0.2 > /dev/null 2> gpurun_out/r02_c4_ncu_list.err; echo "ncu list rc=$?"
