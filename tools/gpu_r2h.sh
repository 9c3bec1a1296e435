#!/bin/bash
set -u
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
timeout 300 python tools/gpu_gemm_sustained.py > $O/sustained.log 2>&1; cat $O/sustained.log
timeout 200 python tools/gpu_mlp_bench.py --only fwd > $O/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/mlp_fwd_launches.csv python tools/gpu_mlp_bench.py --only fwd > $O/ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2h/mlp_fwd_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4][:60], r[8], r[-1])
PY
