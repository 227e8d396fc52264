#!/usr/bin/env bash
# session 8: asynchronous ingest -- parity tests, then the bench with queued vs per-batch-synchronised ingest
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_async_ingest.py tests/test_gpu_mfg_ops.py -m gpu -x -q 2>&1 | tail -15) | tee gpurun_out/s8r_pytest.log
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cache.py -m gpu -x -q 2>&1 | tail -3) | tee -a gpurun_out/s8r_pytest.log
for mode in "" "--ingest-sync"; do
  tag=async; [ -n "$mode" ] && tag=sync
  timeout 300 python bench.py --no-cpu-baseline --e2e-steps 0 $mode > gpurun_out/s8r_bench_$tag.json 2> gpurun_out/s8r_bench_$tag.err || tail -5 gpurun_out/s8r_bench_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/s8r_bench_$tag.json"))
i=d["ingest"]; print("$tag: value %.2f G frac %.3f | ingest %.1f M edges/s %.3f ms/step  %s" % (d["value"]/1e9, d["roofline"]["frac"], i["value"]/1e6, i["ms_per_step"], {k: round(v*1e3,1) for k,v in i["phase_ms_per_batch"].items()}))
PY
done
timeout 200 python bench_configs.py --config online --scale 0.1 > gpurun_out/s8r_online.json 2>gpurun_out/s8r_online.err; cut -c1-300 gpurun_out/s8r_online.json
