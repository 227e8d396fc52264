#!/usr/bin/env bash
# multi-GPU evidence run: headline bench + the multi-GPU configs at N ranks (one box).  usage: scratch/multi_gpu.sh N tag
N=${1:-8}; TAG=${2:-s5}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # name port cmd...
  local name=$1 port=$2; shift 2
  timeout 240 $TR --master-port $port "$@" > gpurun_out/${name}_${N}gpu_${TAG}.json 2> gpurun_out/${name}_${N}gpu_${TAG}.err
  echo "== $name rc=$? $(tail -c 600 gpurun_out/${name}_${N}gpu_${TAG}.json | cut -c1-600)"
}
nvidia-smi topo -m 2>/dev/null | head -12
run bench 29511 bench.py --gpus $N --steps 10 --warmup 3
run dysat_peer 29512 bench_configs.py --config dysat --gpus $N --steps 3 --exchange peer
run dysat_nccl 29513 bench_configs.py --config dysat --gpus $N --steps 2 --exchange nccl --max-batches 50
run tgat 29514 bench_configs.py --config tgat --gpus $N --steps 3
run online 29515 bench_configs.py --config online --gpus $N --steps 3
timeout 200 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -3
