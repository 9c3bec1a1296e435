#!/bin/bash
set -u
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
timeout 300 python tools/gpu_gemm_check.py > $O/gemm_check.log 2>&1; echo "gemm rc=$?" >> $O/gemm_check.log
cp gpurun_out/gemm_check.json $O/ 2>/dev/null
timeout 600 python -m pytest tests -m gpu -q -x -k "render or fused_se3" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -30 $O/gemm_check.log; tail -5 $O/pytest.log
