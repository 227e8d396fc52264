#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --part-scale 0.25 > gpurun_out/r02_c77_bench_8gpu.json 2> gpurun_out/r02_c77_bench_8gpu.err; echo "rc $?"; tail -2 gpurun_out/r02_c77_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c77_bench_8gpu.json').read().strip().splitlines()[-1])
print('value %.2f G'%(d['value']/1e9), d['n_gpus'], 'e2e %.3f G'%(d['e2e']['value']/1e9), d['e2e']['ms_per_step'], d['e2e']['pcie_GBps'])
p=d['partitioned']; print('part', p['value']/1e9, p['x_one_gpu'], p['equals_unpartitioned_sampler'], p['exchange']['GBps_per_gpu'])
PY
