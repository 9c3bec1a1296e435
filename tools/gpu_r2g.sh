#!/bin/bash
set -u
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
rm -f gpurun_out/mlp_report.jsonl
timeout 300 python tools/gpu_gemm_check.py > $O/gemm_check.log 2>&1; echo "gemm rc=$?" >> $O/gemm_check.log
grep -v "^$" $O/gemm_check.log | cut -c1-230
timeout 900 python -m pytest tests/test_gpu_mlp.py -m gpu -q -rf -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
cp gpurun_out/mlp_report.jsonl $O/ 2>/dev/null
grep -n "^E  " $O/pytest.log | head -5; tail -4 $O/pytest.log
timeout 600 python tools/gpu_mlp_bench.py > $O/mlp_bench.json 2> $O/mlp_bench.err; cut -c1-1500 $O/mlp_bench.json; tail -3 $O/mlp_bench.err
