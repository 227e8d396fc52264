#!/usr/bin/env bash
# session 8: GDELT shapes at FULL scale (191 M edges) on one B200
mkdir -p gpurun_out
run() { local tag=$1; shift; timeout 400 python bench_configs.py "$@" > gpurun_out/s8n_$tag.json 2> gpurun_out/s8n_$tag.err; echo "== $tag rc=$? $(cut -c1-1200 gpurun_out/s8n_$tag.json)"; tail -2 gpurun_out/s8n_$tag.err; }
run dysat_16.7M --config dysat --scale 1.0 --shape GDELT-16.7M --max-batches 200 --steps 2
run dysat_16.7K --config dysat --scale 1.0 --shape GDELT-16.7K --max-batches 200 --steps 2
run online_16.7M --config online --scale 1.0 --shape GDELT-16.7M --steps 3
run online_16.7K --config online --scale 1.0 --shape GDELT-16.7K --steps 3
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
