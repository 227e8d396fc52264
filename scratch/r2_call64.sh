#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c64_pytest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r02_c64_pytest.log
timeout 900 python bench.py > gpurun_out/r02_c64_bench.json 2> gpurun_out/r02_c64_bench.err; echo "bench rc $?"; tail -c 600 gpurun_out/r02_c64_bench.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for sh in GDELT-16.7K GDELT-16.7M REDDIT; do
GF_SHAPE=$sh GF_NCU_RANGE=1 timeout 300 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_c64_ingest16m_${sh}_launches.csv python scratch/ingest_100k.py 16000000 >/dev/null 2>&1
done
for sh in GDELT-16.7K REDDIT; do
GF_SHAPE=$sh GF_NCU_RANGE=1 timeout 300 ncu --metrics $M --clock-control none --profile-from-start off -c 24 --csv --log-file gpurun_out/r02_c64_ingest100k_${sh}_launches.csv python scratch/ingest_100k.py 100000 >/dev/null 2>&1
done
GF_SHAPE=GDELT-16.7K GF_NCU_RANGE=1 timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"ingest_(bookkeep|plan|flags_merge|apply)" -c 4 -o gpurun_out/r02_c64_ingest16m_16k -f python scratch/ingest_100k.py 16000000 >/dev/null 2>&1
ls -la gpurun_out | tail -12
