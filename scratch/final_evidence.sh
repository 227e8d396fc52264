#!/usr/bin/env bash
# The evidence pass of a session, run on the GPU box through `gpurun -- bash scratch/final_evidence.sh`: the GPU suite,
# the bench line, the launch list of the bench command, the ingest launch lists (profiles/launch_summary.py turns the
# CSVs into profiles/rNN_*_launches.txt) and one full ncu capture of the large-batch ingest kernels
# (profiles/ncu_summary.py).  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
T=${TAG:-final}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc $?"; tail -3 gpurun_out/${T}_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-hbm-bound --no-cpu-baseline --e2e-steps 1 --no-per-batch-models > /dev/null 2>&1
for sh in GDELT-16.7K GDELT-16.7M REDDIT; do
  GF_SHAPE=$sh GF_NCU_RANGE=1 timeout 300 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_ingest16m_${sh}_launches.csv python scratch/ingest_100k.py 16000000 > /dev/null 2>&1
done
for sh in GDELT-16.7K REDDIT; do
  GF_SHAPE=$sh GF_NCU_RANGE=1 timeout 300 ncu --metrics $M --clock-control none --profile-from-start off -c 24 --csv --log-file gpurun_out/${T}_ingest100k_${sh}_launches.csv python scratch/ingest_100k.py 100000 > /dev/null 2>&1
done
GF_SHAPE=GDELT-16.7K GF_NCU_RANGE=1 timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"ingest_(bookkeep|plan|flags_merge|apply)" -c 4 -o gpurun_out/${T}_ingest16m_16k -f python scratch/ingest_100k.py 16000000 > /dev/null 2>&1
if [ -n "${SAMPLER_NCU:-}" ]; then
  # full captures of the headline kernel and of the sampler on the GDELT shapes (what profiles/make_ncu_traffic.py reads:
  # launches in the order recent L0, recent L1, uniform L0, uniform L1), launch list of one TGAT batch
  B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hbm-bound --e2e-steps 1 --no-per-batch-models"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:sample_persistent -s 6 -c 1 -f -o gpurun_out/${T}_headline $B > /dev/null 2>&1
  for sh in GDELT-16.7K GDELT-16.7M; do
    GF_NCU_RANGE=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sample_persistent -f \
      -o gpurun_out/${T}_hbm_$sh python bench_configs.py --config hbm_bound --shape $sh --scale 0.25 --steps 1 --warmup 3 > /dev/null 2>&1
  done
  timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_tgat_batch_launches.csv python scratch/ncu_fetch.py uniform 10,10 > /dev/null 2>&1
fi
ls -la gpurun_out | tail -14
