#!/usr/bin/env bash
# session 8, multi-GPU check (N ranks on one box): the driver's bench command for both arms + the multi-GPU parity tests
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/s8m_bench_${N}gpu.json 2> gpurun_out/s8m_bench_${N}gpu.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/s8m_bench_${N}gpu.json | cut -c1-700
timeout 300 $TR --master-port 29542 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/s8m_ref_${N}gpu.json 2> gpurun_out/s8m_ref_${N}gpu.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/s8m_ref_${N}gpu.json
timeout 300 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -3
timeout 200 $TR --master-port 29543 bench_configs.py --config two_layer_sat --dataset WIKI --strategy recent > gpurun_out/s8m_wiki2l_${N}gpu.json 2> gpurun_out/s8m_wiki2l.err; echo "wiki2l rc=$?"; cut -c1-200 gpurun_out/s8m_wiki2l_${N}gpu.json
