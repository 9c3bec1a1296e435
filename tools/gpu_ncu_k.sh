#!/bin/bash
# ncu --set full of kernels matching $2 in the default bench (8 views): tools/gpu_ncu_k.sh <outdir> <regex> <skip> <count>
set -u
O=gpurun_out/${1:-ncu_k}
mkdir -p $O
CMD="python bench.py --steps 1 --warmup 3 --streams 1 --no-cpu-baseline --no-mlp"
timeout 300 $CMD > $O/plain.json 2> $O/plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s ${3:-2} -c ${4:-1} -o $O/rep $CMD > $O/ncu.log 2>&1
tail -3 $O/ncu.log
