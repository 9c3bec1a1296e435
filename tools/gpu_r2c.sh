#!/bin/bash
set -u
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
rm -f gpurun_out/mlp_report.jsonl
timeout 900 python -m pytest tests/test_gpu_mlp.py -m gpu -q -rf > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
cp gpurun_out/mlp_report.jsonl $O/ 2>/dev/null
tail -60 $O/pytest.log
