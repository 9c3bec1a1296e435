#!/bin/bash
set -u
O=gpurun_out/${1:-ncu_ds}
mkdir -p $O
CMD="python bench.py --steps 1 --warmup 3 --views 1 --streams 1 --no-cpu-baseline --no-mlp --sync-free 0"
timeout 300 $CMD > $O/plain.json 2> $O/plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${2:-ds_}" -s ${3:-9} -c ${4:-3} -o $O/rep $CMD > $O/ncu.log 2>&1
tail -3 $O/ncu.log; ls -la $O
