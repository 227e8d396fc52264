#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q -rs --timeout 500 > gpurun_out/r02_c15_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_c15_pytest.log
for sb in 64 256; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench_configs.py --config partitioned --shape GDELT-16.7K --scale 0.25 --steps 3 --warmup 2 --max-batches $sb > gpurun_out/r02_c15_part_$sb.json 2> gpurun_out/r02_c15_part_$sb.err; echo "part $sb rc=$?"
tail -2 gpurun_out/r02_c15_part_$sb.err | cut -c1-400
python - <<P
import json
d=json.load(open('gpurun_out/r02_c15_part_$sb.json'))['partitioned']
print({k:d[k] for k in ('value','ms_per_step','equals_unpartitioned_sampler','x_one_gpu','phase_ms_per_layer_snapshot_step_rank0')}, d['one_gpu_unpartitioned_same_call'], d['exchange']['GBps_per_gpu'])
P
done
