#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
NG=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 scratch/d2h_ceiling.py > gpurun_out/r02_d2h_ceiling_8gpu.json 2> gpurun_out/r02_c25_d2h.err; echo "d2h rc=$?"; cat gpurun_out/r02_d2h_ceiling_8gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_c25_bench8.err; echo "bench8 rc=$?"
tail -2 gpurun_out/r02_c25_bench8.err | cut -c1-300
python - <<P
import json
d=json.loads(open('gpurun_out/r02_bench_8gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['n_gpus'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('pcie_GBps'), d['e2e'].get('host_affinity'))
p=d.get('partitioned',{}); print({k:p.get(k) for k in ('value','ms_per_step','equals_unpartitioned_sampler','x_one_gpu')}, p.get('exchange'), p.get('features'))
P
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus $NG --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_8gpu.json 2> gpurun_out/r02_c25_ref8.err; echo "ref8 rc=$?"
cut -c1-400 gpurun_out/r02_bench_reference_8gpu.json
