#!/bin/bash
set -u
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
timeout 600 python tools/gpu_mlp_bench.py > $O/mlp_bench.json 2> $O/mlp_bench.err
cat $O/mlp_bench.json | cut -c1-1500
timeout 200 python tools/gpu_mlp_bench.py --only fwd --P 300000 > $O/plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_gemm_kernel -s 2 -c 2 -o $O/gemm_prof python tools/gpu_mlp_bench.py --only fwd --P 300000 > $O/ncu.log 2>&1
tail -3 $O/ncu.log
