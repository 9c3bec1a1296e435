#!/bin/bash
# round 2 profile captures: launch list of one view (ours), ncu --set full of the dominant kernels, GEMM kernels
set -u
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
CMD="python bench.py --steps 1 --warmup 3 --views 1 --streams 1 --no-cpu-baseline --no-mlp --sync-free 0"
timeout 300 $CMD > $O/plain_v1.json 2> $O/plain_v1.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file $O/ours_launches.csv $CMD > $O/ncu1.log 2>&1
timeout 300 $CMD > /dev/null 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"blend_bwd_v2|blend_fwd_v2|preprocess_fwd|preprocess_bwd|tile_scatter|onesweep" -s 60 -c 9 -o $O/ours_full $CMD > $O/ncu2.log 2>&1
timeout 200 python tools/gpu_mlp_bench.py --only fwd --P 1000000 > $O/plain_mlp.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_gemm_kernel -s 3 -c 2 -o $O/gemm_v2_full python tools/gpu_mlp_bench.py --only fwd --P 1000000 > $O/ncu3.log 2>&1
ls -la $O; tail -2 $O/ncu1.log $O/ncu2.log $O/ncu3.log
