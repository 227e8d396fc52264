#!/usr/bin/env bash
# session 8: 4-GPU e2e with every rank bound to the host cores next to its GPU
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14
python - <<PY
import os; print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print(open("/sys/devices/system/node/online").read().strip(), "numa nodes online")
except Exception as e: print(e)
PY
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29561 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s8q_bench_${N}gpu.json 2> gpurun_out/s8q_bench_${N}gpu.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/s8q_bench_${N}gpu.json").read().strip().splitlines()[-1])
e=d["e2e"]; print("value %.1f G  e2e %.2f G  %.2f ms/step  %.1f GB/s per gpu  | %s" % (d["value"]/1e9, e["value"]/1e9, e["ms_per_step"], e["pcie_GBps"], e.get("host_affinity")))
PY
tail -3 gpurun_out/s8q_bench_${N}gpu.err
