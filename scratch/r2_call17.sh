#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_dispatch.py -m gpu -q -rs --timeout 500 > gpurun_out/r02_c17_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_c17_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_c17_bench2.json 2> gpurun_out/r02_c17_bench2.err; echo "bench2 rc=$?"
tail -3 gpurun_out/r02_c17_bench2.err | cut -c1-300
python - <<P
import json
d=json.loads(open('gpurun_out/r02_c17_bench2.json').read().strip().splitlines()[-1])
print(d['value'], d['n_gpus'], d['e2e']['value'], json.dumps(d.get('partitioned'))[:1500])
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_c17_ref2.json 2> gpurun_out/r02_c17_ref2.err; echo "ref2 rc=$?"
cut -c1-600 gpurun_out/r02_c17_ref2.json
