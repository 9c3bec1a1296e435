#!/bin/bash
set -u
O=gpurun_out/${1:-ncu_b2}
mkdir -p $O
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; tail -4 $O/smoke.log
CMD="python bench.py --steps 1 --warmup 3 --streams 1 --no-cpu-baseline --no-mlp"
timeout 300 $CMD > $O/plain.json 2> $O/plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"preprocess_fwd_batched|preprocess_bwd_batched" -s 6 -c 2 -o $O/rep $CMD > $O/ncu.log 2>&1
tail -3 $O/ncu.log
