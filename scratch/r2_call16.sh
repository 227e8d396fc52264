#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
NG=${NG:-8}
for cfg in "GDELT-16.7K 1.0 64" "GDELT-16.7K 1.0 256" "GDELT-16.7M 1.0 64"; do set -- $cfg
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench_configs.py --config partitioned --shape $1 --scale $2 --steps 5 --warmup 2 --max-batches $3 > gpurun_out/r02_c16_part${NG}_$1_$3.json 2> gpurun_out/r02_c16_part${NG}_$1_$3.err; echo "part $cfg rc=$?"
tail -2 gpurun_out/r02_c16_part${NG}_$1_$3.err | cut -c1-400
python - <<P
import json
d=json.load(open('gpurun_out/r02_c16_part${NG}_$1_$3.json'))['partitioned']
print({k:d[k] for k in ('value','ms_per_step','equals_unpartitioned_sampler','x_one_gpu','phase_ms_per_layer_snapshot_step_rank0')}, d['one_gpu_unpartitioned_same_call'], d['exchange'], d['features'], d['ingest'])
P
done
