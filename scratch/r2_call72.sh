#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --part-scale 0.1 > gpurun_out/r02_c72_bench_2gpu.json 2> gpurun_out/r02_c72_bench_2gpu.err; echo "rc $?"; tail -3 gpurun_out/r02_c72_bench_2gpu.err; tail -c 1500 gpurun_out/r02_c72_bench_2gpu.json
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -2
