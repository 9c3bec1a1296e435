#!/bin/bash
set -u
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
rm -f gpurun_out/mlp_report.jsonl
timeout 300 python tools/gpu_gemm_check.py > $O/gemm_check.log 2>&1; echo "gemm rc=$?" >> $O/gemm_check.log
timeout 900 python -m pytest tests/test_gpu_mlp.py -m gpu -q -rf > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
cp gpurun_out/mlp_report.jsonl $O/ 2>/dev/null
grep -v "^$" $O/gemm_check.log | cut -c1-220; grep -n "^E  " $O/pytest.log | head; tail -8 $O/pytest.log
