#!/usr/bin/env bash
# session 8: 8-GPU evidence (one box): driver's bench command, DySAT hash-partitioned over peer memory, TGAT data-parallel
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # name port timeout cmd...
  local name=$1 port=$2 to=$3; shift 3
  timeout $to $TR --master-port $port "$@" > gpurun_out/s8p_${name}_${N}gpu.json 2> gpurun_out/s8p_${name}_${N}gpu.err
  echo "== $name rc=$? $(tail -c 2500 gpurun_out/s8p_${name}_${N}gpu.json | cut -c1-900)"
}
run bench 29551 100 bench.py --gpus $N --steps 20 --warmup 3
run dysat_peer 29552 80 bench_configs.py --config dysat --gpus $N --steps 2 --max-batches 100 --exchange peer
run tgat 29553 80 bench_configs.py --config tgat --gpus $N --steps 2 --max-batches 100
