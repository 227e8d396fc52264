#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_dispatch.py -m gpu -q -rs --timeout 500 > gpurun_out/r02_c14_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r02_c14_pytest.log
for sh in GDELT-16.7K GDELT-16.7M; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench_configs.py --config partitioned --shape $sh --scale 0.25 --steps 3 --warmup 2 > gpurun_out/r02_c14_part_$sh.json 2> gpurun_out/r02_c14_part_$sh.err; echo "part $sh rc=$?"
tail -3 gpurun_out/r02_c14_part_$sh.err | cut -c1-600
cut -c1-2500 gpurun_out/r02_c14_part_$sh.json
done
