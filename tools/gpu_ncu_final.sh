#!/bin/bash
set -u
O=gpurun_out/ncu_final; mkdir -p $O
CMD="python bench.py --steps 1 --warmup 3 --streams 1 --no-cpu-baseline --no-mlp --sync-free 0"
timeout 300 $CMD > $O/plain.json 2> $O/plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd_v2|tile_scatter|tile_count|tile_column_scan" -s 12 -c 4 -o $O/rep $CMD > $O/ncu.log 2>&1
tail -2 $O/ncu.log
