#!/bin/bash
# round-3 (second session of round 2) evidence: bench lines of both arms at C2 / C3 / C4, the per-view variants, ncu launch list
# of one step and ncu --set full of the pipeline's kernels
set -u
O=gpurun_out/r3p
mkdir -p $O
timeout 600 python bench.py > $O/bench_ours_c2.json 2> $O/bench_ours_c2.err
timeout 600 python bench.py --impl reference > $O/bench_ref_c2.json 2> $O/bench_ref_c2.err
timeout 300 python bench.py --batched-backward 0 --no-cpu-baseline --no-mlp > $O/bench_ours_c2_per_view_backward.json 2> /dev/null
timeout 300 python bench.py --views 1 --streams 1 --no-cpu-baseline --no-mlp > $O/bench_ours_c2_one_view.json 2> /dev/null
timeout 300 python bench.py --config C3 --no-cpu-baseline --no-mlp > $O/bench_ours_c3.json 2> /dev/null
timeout 300 python bench.py --config C4 --no-cpu-baseline --no-mlp > $O/bench_ours_c4.json 2> /dev/null
timeout 300 python bench.py --train --no-cpu-baseline --no-mlp > $O/bench_ours_c2_train.json 2> /dev/null
CMD="python bench.py --steps 1 --warmup 3 --streams 1 --no-cpu-baseline --no-mlp --sync-free 0"
timeout 300 $CMD > $O/plain.json 2> $O/plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 130 --csv --log-file $O/ncu_launches.csv $CMD > $O/ncu1.log 2>&1
timeout 300 $CMD > /dev/null 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"blend_bwd_v2|blend_fwd_v2|preprocess_fwd|preprocess_bwd_batched|tile_scatter|tile_count|tile_column_scan|ds_hist|ds_scan|ds_scatter|ds_local" -s 33 -c 18 -o $O/ours_full $CMD > $O/ncu2.log 2>&1
tail -2 $O/ncu1.log $O/ncu2.log; ls -la $O
